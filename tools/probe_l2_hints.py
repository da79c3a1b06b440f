"""Developer probe: A/B of row-pass staging variants (GloveStep(variant=...)) against the default (variant 0).

    python tools/probe_l2_hints.py --variants 0,3 [--steps 20] [--B 262144]   # parent: one child process per variant
    python tools/probe_l2_hints.py --variant 3                                # child

Variant 3 = `accreg` (accumulator rows through ld.global.cs registers, evict-first in L2; in the library, experimental).
HISTORY (round 1): the first use of this probe tested L2::cache_hint descriptors on the cp.async copies -- at that
time numbered variants 3 / 4 -- and both faulted with "an illegal instruction was encountered" on B200
(profiles/r1_l2_hints_probe.json).  That kernel change is NOT in the library; it is kept as
profiles/r1_l2_hint_probe.patch (a diff against commit 1644865's csrc/glove_step.cu).

Each child trains the same seeded table on the same Zipf and uniform batches, prints the row-pass time (CUDA
events, 256 MB L2 flush between steps) and a bit-pattern checksum of the final table / accumulator / bias.  The
variants do not touch the arithmetic, so the checksums must equal variant 0's -- that is the parity check.  One process
per variant so that a faulting variant cannot take the others' numbers with it.  Not a benchmark of record.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import numpy as np
    import torch
    from esrecsys_b200 import engine, synth
    out = {"variant": a.variant}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, uniform in (("zipf", False), ("uniform", True)):
        ids, counts = synth.glove_batches(a.V, a.B, a.nbatch, 0, uniform)
        U = float(np.mean([np.unique(ids[k]).size for k in range(a.nbatch)]))
        t = engine.EmbeddingTable(a.V, a.D)
        t.rows0.copy_(torch.randn(a.V, a.D, generator=torch.Generator().manual_seed(1)) / np.sqrt(a.D))
        step = engine.GloveStep(t, a.B, variant=a.variant)
        plan = engine.IndexPlan(2 * a.B, a.V)
        d_ids = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(a.nbatch)]
        d_cnt = [torch.from_numpy(counts[k]).cuda() for k in range(a.nbatch)]
        ms = 0.0
        for it in range(a.steps + 3):
            k = it % a.nbatch
            flush.zero_()
            plan.build(d_ids[k])
            step.prep(plan, d_cnt[k])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step.rows_main(plan)
            e1.record()
            step.rows_combine(plan)
            step.finish(plan)
            torch.cuda.synchronize()
            if it >= 3:
                ms += e0.elapsed_time(e1)
        ms /= a.steps
        alg = U * a.D * 4 * 4 + U * 16 + a.B * 12
        chk = [int(x.contiguous().view(torch.int32).sum(dtype=torch.int64).item()) for x in (t.dense(), t.acc, t.bias)]
        out[name] = {"rows_ms": ms, "alg_GBs": alg / (ms * 1e-3) / 1e9, "U": U, "checksum": chk,
                     "loss": float(step.scalars[5].item())}
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=int, default=1000000)
    ap.add_argument("--D", type=int, default=128)
    ap.add_argument("--B", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--nbatch", type=int, default=4)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--variants", default="0,3")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "l2_hints.json"))
    a = ap.parse_args()
    if a.variant >= 0:
        return child(a)
    res = []
    for v in [int(x) for x in a.variants.split(",")]:
        cmd = [sys.executable, os.path.abspath(__file__), "--variant", str(v), "--V", str(a.V), "--D", str(a.D), "--B", str(a.B),
               "--steps", str(a.steps), "--nbatch", str(a.nbatch)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=60)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
            res.append(json.loads(line[0][7:]) if line else {"variant": v, "failed": (r.stderr or r.stdout)[-600:]})
        except subprocess.TimeoutExpired:
            res.append({"variant": v, "failed": "timeout"})
    base = next((r for r in res if r.get("variant") == 0 and "zipf" in r), None)
    for r in res:
        if base and "zipf" in r:
            r["bit_identical_to_default"] = all(r[s]["checksum"] == base[s]["checksum"] for s in ("zipf", "uniform"))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
