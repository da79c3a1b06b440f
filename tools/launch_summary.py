"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum[,...]` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches_per_kernel]
"""
import collections
import csv
import re
import sys


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void |esr::<unnamed>::|<unnamed>::", "", name)[:64]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(name, collections.OrderedDict()).setdefault(r[mi], []).append(v)
    tot = sum(sum(m.get("gpu__time_duration.sum", [0])[skip:]) for m in agg.values())
    print("%-66s %5s %10s %7s  other metrics (mean)" % ("kernel", "n", "avg_us", "share"))
    for name, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0])[skip:])):
        t = m.get("gpu__time_duration.sum", [0])[skip:]
        if not t:
            continue
        extra = "  ".join("%s=%.3g" % (k.split(".")[0][-28:], sum(v[skip:]) / max(1, len(v[skip:]))) for k, v in m.items()
                          if k != "gpu__time_duration.sum")
        print("%-66s %5d %10.2f %6.1f%%  %s" % (name, len(t), sum(t) / len(t) / 1e3, 100 * sum(t) / max(tot, 1), extra))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
