"""Developer probe: per-phase device time of the GloVe step (CUDA events), for kernel tuning.

    python tools/prof_glove.py --V 1000000 --D 128 --B 65536 [--uniform] [--chunk 32] [--steps 20]

Not a benchmark of record (bench.py is); prints one JSON line per configuration.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=int, default=1000000)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--uniform", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--nbatch", type=int, default=8)
    ap.add_argument("--bias_mode", default="reference_broadcast")
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    a = ap.parse_args()
    ids, counts = synth.glove_batches(a.V, a.B, a.nbatch, 0, a.uniform)
    U = np.mean([np.unique(ids[k]).size for k in range(a.nbatch)])
    t = engine.EmbeddingTable(a.V, a.D)
    t.rows0.normal_(0, 1.0 / np.sqrt(a.D))
    step = engine.GloveStep(t, a.B, bias_mode=a.bias_mode, chunk=a.chunk, impl=a.impl, variant=a.variant)
    plan = engine.IndexPlan(2 * a.B, a.V)
    d_ids = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(a.nbatch)]
    d_cnt = [torch.from_numpy(counts[k]).cuda() for k in range(a.nbatch)]
    names = ["plan", "prep", "rows", "finish"]
    acc = {n: 0.0 for n in names}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(a.steps + 3):
        k = it % a.nbatch
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        plan.build(d_ids[k])
        ev[1].record()
        step.prep(plan, d_cnt[k])
        ev[2].record()
        step.rows(plan)
        ev[3].record()
        step.finish(plan)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1])
    ms = {n: acc[n] / a.steps for n in names}
    R = a.D * 4
    alg = U * R * 4 + U * 16 + a.B * 12
    out = dict(impl=a.impl, variant=a.variant, V=a.V, D=a.D, B=a.B, uniform=a.uniform, chunk=a.chunk, U=U, ms=ms, total_ms=sum(ms.values()),
               rows_GBs=alg / (ms["rows"] * 1e-3) / 1e9, step_GBs=alg / (sum(ms.values()) * 1e-3) / 1e9,
               pairs_per_s=a.B / (sum(ms.values()) * 1e-3), loss=float(step.scalars[5].item()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
