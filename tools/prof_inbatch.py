"""Developer probe: device time of the in-batch score path (CUDA events) and its tensor-pipe rate.

    python tools/prof_inbatch.py --B 8192 --D 128 --loss hinge [--steps 50]

flops counted: forward 2*Bq*Bk*D per score pass (softmax runs two) + 4*Bq*Bk*D for dQ and dK.
Not a benchmark of record (bench.py is); prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200.engine import InBatchScorer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8192)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--loss", default="hinge")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--splits", type=int, default=0)
    ap.add_argument("--chunk_rows", type=int, default=0)
    a = ap.parse_args()
    torch.manual_seed(0)
    Q = torch.randn(a.B, a.D, device="cuda") / a.D ** 0.25
    K = torch.randn(a.B, a.D, device="cuda") / a.D ** 0.25
    sc = InBatchScorer(a.B, a.D, loss=a.loss, splits=a.splits, chunk_rows=a.chunk_rows)
    for _ in range(5):
        sc.run(Q, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        sc.run(Q, K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    passes = 2 if a.loss == "softmax" else 1
    useful = 6.0 * a.B * a.B * a.D
    issued = (2.0 * passes + 4.0) * a.B * a.B * a.D
    print(json.dumps(dict(B=a.B, D=a.D, loss=a.loss, ms=ms, pairs_per_s=a.B / (ms * 1e-3), useful_tflops=useful / (ms * 1e-3) / 1e12,
                          issued_tflops=issued / (ms * 1e-3) / 1e12, loss_value=float(sc.loss.item()))))


if __name__ == "__main__":
    main()
