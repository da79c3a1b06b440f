"""One table out of gpurun_out/r2_*: bench lines (value / e2e / row-pass fraction on the Zipf and uniform streams), the
row-pass A/B probe (time per variant, checksum equality), the table sweep, and the tails of the test logs.

    python tools/r2_digest.py [gpurun_out]
"""
import glob
import json
import os
import sys


def last_json_line(path):
    try:
        for line in reversed(open(path).read().splitlines()):
            line = line.strip()
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        pass
    return None


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    print("== bench lines ==")
    print("%-34s %5s %10s %10s %9s %9s %9s" % ("file", "gpus", "value G/s", "e2e G/s", "us/step", "frac zipf", "frac unif"))
    for f in sorted(glob.glob(os.path.join(d, "r2_bench_*.json"))):
        j = last_json_line(f)
        if not j or "value" not in j:
            err = f[:-5] + ".err"
            tail = open(err).read()[-300:].replace("\n", " | ") if os.path.exists(err) else ""
            print("%-34s  no line  %s" % (os.path.basename(f), tail))
            continue
        print("%-34s %5d %10.3f %10.3f %9.1f %9s %9s" % (
            os.path.basename(f), j.get("n_gpus", 1), j["value"] / 1e9, j.get("e2e", {}).get("value", 0) / 1e9,
            j["ms_per_step"] * 1e3, "%.3f" % j["roofline"]["frac"] if "roofline" in j else "-",
            "%.3f" % j["roofline_uniform"]["frac"] if "roofline_uniform" in j else "-"))
    for f in sorted(glob.glob(os.path.join(d, "r2_probe_*.json"))):
        print("== row-pass probe %s ==" % os.path.basename(f))
        try:
            res = json.load(open(f))
        except Exception as e:
            print("  unreadable:", e)
            continue
        rows = res if isinstance(res, list) else res.get("results", [])
        base = next((r for r in rows if r.get("variant") == 0 and "zipf" in r), None)
        for r in rows:
            if "zipf" not in r:
                print("  variant %s FAILED: %s" % (r.get("variant"), " ".join(str(r.get("failed")).split())[-160:]))
                continue
            same = base is not None and all(r[s]["checksum"] == base[s]["checksum"] for s in ("zipf", "uniform"))
            print("  variant %s: zipf %.1f us (%.0f GB/s), uniform %.1f us (%.0f GB/s), checksums %s" % (
                r["variant"], r["zipf"]["rows_ms"] * 1e3, r["zipf"]["alg_GBs"], r["uniform"]["rows_ms"] * 1e3,
                r["uniform"]["alg_GBs"], "== variant 0" if same else "DIFFER"))
    for f in sorted(glob.glob(os.path.join(d, "r2_table_sweep_*.json"))):
        j = last_json_line(f)
        print("== %s ==" % os.path.basename(f))
        if not j:
            print("  no line")
            continue
        for k in ("lookup_zipf", "lookup_uniform"):
            if k in j:
                print("  %-15s %.1f us, %.0f GB/s per direction per GPU (%.2f of NVLink), %d rows checked" % (
                    k, j[k]["us_per_lookup"], j[k]["gbs_per_dir_per_gpu"], j[k]["frac_of_nvlink"], j[k]["rows_checked"]))
        if "step" in j:
            print("  step            %.1f us, %.3f G pairs/s" % (j["step"]["us_per_step"], j["step"]["pairs_per_s"] / 1e9))
    print("== test logs ==")
    for f in sorted(glob.glob(os.path.join(d, "r2_*.log"))):
        lines = [x for x in open(f, errors="replace").read().splitlines() if x.strip()]
        print("%-34s %s" % (os.path.basename(f), lines[-1][:160] if lines else "(empty)"))


if __name__ == "__main__":
    main()
