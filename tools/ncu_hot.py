"""Top stall sites of an .ncu-rep source page (SASS level).   python tools/ncu_hot.py rep [N]"""
import csv
import io
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    h = rows[0]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stalls = [(k, h.index(k)) for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data = []
    for r in rows[1:]:
        try:
            data.append((int(r[isamp] or 0), r))
        except ValueError:
            pass
    tot = sum(d[0] for d in data)
    print("total samples", tot)
    for s, r in sorted(data, key=lambda x: -x[0])[:n]:
        top = sorted(((int(r[i] or 0), k) for k, i in stalls), reverse=True)[:3]
        print("%5.1f%% ex=%-9s %-70s %s" % (100.0 * s / max(tot, 1), r[iex], r[isrc][:70], " ".join("%s=%d" % (k[6:], v) for v, k in top if v)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
