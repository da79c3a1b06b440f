"""configs[3] two-tower trainer step, eager, for an ncu launch list: which kernels a step launches and what share the
MLP towers (cuBLAS GEMMs + elementwise through torch) take next to libesr's gather / score / Adagrad kernels.
   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file out.csv python tools/prof_twotower.py
   python tools/prof_twotower.py --summarize out.csv"""
import argparse
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(steps):
    import torch
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import TwoTowerInBatch
    V, D, B = 1_000_000, 256, 4096
    ts = engine.EmbeddingTable(V, D, sparse=False, adagrad=True)
    tp = engine.EmbeddingTable(V, D, sparse=False, adagrad=True)
    ts.rows0.normal_(0.0, 1.0 / D ** 0.5)
    tp.rows0.normal_(0.0, 1.0 / D ** 0.5)
    qs, ks = synth.pair_batches(V, V, B, 4, 6)
    sid = [torch.from_numpy(qs[k]).cuda() for k in range(4)]
    pid = [torch.from_numpy(ks[k]).cuda() for k in range(4)]
    tr = TwoTowerInBatch(ts, tp, B, loss="softmax")
    for k in range(steps):
        tr.step(sid[k % 4], pid[k % 4])
    torch.cuda.synchronize()


def summarize(path, steps, skip_steps):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    launches = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:]]
    per = len(launches) // steps
    tail = launches[per * skip_steps:per * steps]          # drop the first steps (lazy init, cuBLAS heuristics)
    n = steps - skip_steps
    agg = {}
    for name, ns in tail:
        short = name.split("(")[0].replace("void ", "")[:70]
        own = "esr::" in name or "k_inbatch" in name or "k_topk" in name
        kind = "libesr" if own else ("cuBLAS / cutlass GEMM" if any(t in name for t in ("gemm", "cutlass", "sm90", "sm100", "nvjet", "cublas")) else "torch elementwise / reduce")
        a = agg.setdefault((kind, short), [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    print("launches per step: %d, kernel time per step %.1f us (ncu: cold caches, serialised)" % (len(tail) // n, tot / n / 1e3))
    by_kind = {}
    for (kind, short), (c, ns) in agg.items():
        by_kind[kind] = by_kind.get(kind, 0.0) + ns
    for kind, ns in sorted(by_kind.items(), key=lambda x: -x[1]):
        print("  %-28s %7.1f us/step  %5.1f %%" % (kind, ns / n / 1e3, 100 * ns / tot))
    print("top kernels:")
    for (kind, short), (c, ns) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
        print("  %7.1f us/step  x%-3d %-22s %s" % (ns / n / 1e3, c // n, kind, short))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--summarize", default=None)
    a = ap.parse_args()
    if a.summarize:
        summarize(a.summarize, a.steps, 3)
    else:
        run(a.steps)
