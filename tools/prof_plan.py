"""Index plan alone (esr_plan_build_i32): hand-written slot sort vs the cub control, CUDA-event time per build.
   python tools/prof_plan.py [--batch 262144]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import engine as eng, synth  # noqa: E402


def time_build(plan, keys, iters=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(20):
            plan.build(keys, stream=s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            plan.build(keys, stream=s)
        for _ in range(10):
            g.replay()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        for _ in range(iters):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=262144)
    ap.add_argument("--eager", type=int, default=0, help="N eager builds of the V = 1M Zipf plan only (for ncu)")
    ap.add_argument("--trace", action="store_true", help="per-block phase timestamps of the sort kernels (ESR_PLAN_TRACE)")
    a = ap.parse_args()
    out = {}
    if a.trace:
        os.environ["ESR_PLAN_TRACE"] = "1"
        ids, _ = synth.glove_batches(10 ** 6, a.batch, 1, 7, False)
        keys = torch.from_numpy(ids[0].reshape(-1)).cuda()
        plan = eng.IndexPlan(2 * a.batch, 10 ** 6, sort="wide")
        for _ in range(5):
            plan.build(keys)
        torch.cuda.synchronize()
        tb = 6 * 512 * 8 * 8
        off = (plan.ws_bytes - tb) // 256 * 256
        tr = plan.ws[off:off + tb].cpu().numpy().view(np.uint64).reshape(6, 512, 8).astype(np.int64)
        tiles = (2 * a.batch + 2047) // 2048
        t00 = tr[0, :tiles, 0].min()
        names = ["hist", "pass0", "pass1", "pass2", "pass3", "heads"]
        for k in range(6):
            x = tr[k, :tiles]
            if x[:, 0].max() == 0:
                continue
            t0 = x[:, 0].min()
            polls = x[:, 7].copy()
            x = x.copy()
            x[:, 7] = 0
            nph = int((x.max(axis=0) > 0).sum())
            if polls.max() > 0:
                print("   polls until ready (tile words / group words), every 16th tile:",
                      [(int(v >> 32), int(v & 0xffffffff)) for v in polls[0:tiles:16]])
            print("%s: first block starts at %.2f us after hist start; block starts spread %.2f us" %
                  (names[k], (t0 - t00) / 1e3, (x[:, 0].max() - t0) / 1e3))
            for ph in range(nph):
                col = x[:, ph]
                col = col[col > 0] - t0
                print("   phase %d: min %.2f  median %.2f  p90 %.2f  max %.2f us" %
                      (ph, col.min() / 1e3, np.median(col) / 1e3, np.percentile(col, 90) / 1e3, col.max() / 1e3))
            d = (x[:, 1:nph] - x[:, 0:nph - 1])
            print("   median phase durations (us):", [round(float(np.median(d[:, q])) / 1e3, 2) for q in range(nph - 1)])
            order = np.argsort(x[:, 0])
            print("   start time by tile id (every 32nd):", [round(float(x[t, 0] - t0) / 1e3, 2) for t in range(0, tiles, 32)])
        return
    if a.eager:
        ids, _ = synth.glove_batches(10 ** 6, a.batch, 1, 7, False)
        keys = torch.from_numpy(ids[0].reshape(-1)).cuda()
        plan = eng.IndexPlan(2 * a.batch, 10 ** 6, sort="wide")
        for _ in range(a.eager):
            plan.build(keys)
        torch.cuda.synchronize()
        return
    for V in (10 ** 6, 10 ** 8):
        for uniform in (False, True):
            ids, _ = synth.glove_batches(V, a.batch, 1, 7, uniform)
            keys = torch.from_numpy(ids[0].reshape(-1)).cuda()
            for sort in ("own", "cub"):
                os.environ["ESR_PLAN_SORT"] = sort
                plan = eng.IndexPlan(2 * a.batch, V)
                us = time_build(plan, keys)
                out["V%g_%s_%s" % (V, "uniform" if uniform else "zipf", sort)] = round(us, 2)
    print(json.dumps({"plan_build_us": out, "slots": 2 * a.batch}))


if __name__ == "__main__":
    main()
