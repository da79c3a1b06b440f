"""BASELINE.json configs[4]: table-scaling sweep -- V rows x D floats row-sharded (cyclic) over the ranks of one box,
lookup GB/s over NVLink against the 900 GB/s per-direction peak, plus the full sharded GloVe step on the same table.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29621 \
        tools/table_sweep.py [--vocab 100000000 --dim 128 --batch 262144 --steps 30]

Per rank the shard is V/N rows (+ the Adagrad slot): 100M x 128 at N = 2 is 25.6 + 25.6 GB per GPU, at N = 8
6.4 + 6.4 GB.  Prints ONE JSON line on rank 0.

What is measured (device time, CUDA events, max over ranks; SURVEY.md 8(d) K7 byte accounting):
  lookup      esr_peer_gather_f32 of the unique rows of a (2,B) id batch: per rank and step
              n_remote_unique * (4 + R) bytes arrive over NVLink (R = 4 D) -> GB/s per direction per GPU;
              Zipf(1) stream and the uniform no-reuse control.
  step        PeerShardedGloveTrainer.step (fetch + fused step + gradient scatter to the owners + merge/Adagrad).
What is CHECKED at full size (size-independent properties, bit-exact):
  * every shard row is initialised to a closed-form function of its GLOBAL row id; every looked-up row must equal that
    function of the requested id (lookup == identity on the function), rows and biases;
  * after one training step a sample of rows OUTSIDE the (global) batch is bit-identical, no Adagrad accumulator of a
    row INSIDE the batch shrank, and most of those rows moved.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NVLINK_GBS_PER_DIR = 900.0


def row_function(global_rows, D, device):
    """f32 (len, D): value of row g, column c = ((131 g + 7919 c) mod 2^20) / 2^20 - 0.5 (exact in f32)."""
    g = global_rows.to(device=device, dtype=torch.int64).reshape(-1, 1)
    c = torch.arange(D, device=device, dtype=torch.int64).reshape(1, -1)
    return (((g * 131 + c * 7919) & 0xFFFFF).to(torch.float32) * (1.0 / (1 << 20)) - 0.5).contiguous()


def bias_function(global_rows, device):
    g = global_rows.to(device=device, dtype=torch.int64)
    return ((g * 40503) & 0xFFFF).to(torch.float32) * (1.0 / (1 << 20))


def remote_unique(uniq, rank, n):
    """Number of unique rows of this rank's batch owned by another rank (cyclic ownership)."""
    return int((uniq % n != rank).sum())


def lookup_bytes(n_remote, D):
    """ids out (4 B) + rows back (R B) per remote unique row, per direction the larger is R."""
    return n_remote * (4 + 4 * D)


def fill_shard(tr, chunk=1 << 20):
    V_loc = tr.shard.V
    for s in range(0, V_loc, chunk):
        e = min(V_loc, s + chunk)
        g = torch.arange(s, e, device=tr.dev, dtype=torch.int64) * tr.n + tr.rank
        tr.shard.rows0[s:e] = row_function(g, tr.D, tr.dev)
        tr.shard.bias[s:e] = bias_function(g, tr.dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocab", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=262144, help="pairs per GPU per step")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nbatch", type=int, default=4)
    ap.add_argument("--no-step", action="store_true", help="lookup only")
    a = ap.parse_args()

    from esrecsys_b200 import _lib as L, synth
    from esrecsys_b200.engine import IndexPlan
    from esrecsys_b200.sharded import PeerShardedGloveTrainer

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    L.require_cuda()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, D, B = a.vocab, a.dim, a.batch
    tr = PeerShardedGloveTrainer(V, D, B)
    fill_shard(tr)
    torch.cuda.synchronize()
    tr.barrier()
    lib = L.lib()
    out = {"workload": "table sweep: %d rows x %d, cyclic row-sharding over %d GPUs, B = %d pairs per GPU" % (V, D, world, B),
           "n_gpus": world, "shard_gb": tr.shard.V * D * 4 / 1e9, "nvlink_peak_gbs_per_dir": NVLINK_GBS_PER_DIR}

    plan = IndexPlan(2 * B, V, tr.dev)
    rows_out = torch.empty(2 * B, D, device=tr.dev)
    bias_out = torch.empty(2 * B, device=tr.dev)

    def lookup_leg(uniform):
        ids, _ = synth.glove_batches(V, B, a.nbatch, 17 * rank + (99 if uniform else 0), uniform=uniform)
        d_ids = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(a.nbatch)]
        tot_ms, tot_bytes, checked = 0.0, 0, 0
        for it in range(a.steps + a.warmup):
            plan.build(d_ids[it % a.nbatch])
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.esr_peer_gather_f32(tr.p_rows, tr.p_bias, world, L.ptr(plan.uniq), L.ptr(plan.n_uniq), plan.capacity, D,
                                            L.ptr(rows_out), L.ptr(bias_out), L.stream_ptr()), "esr_peer_gather_f32")
            e1.record()
            torch.cuda.synchronize()
            U = int(plan.n_uniq.item())
            uniq = plan.uniq[:U]
            if it < a.warmup or it == a.steps + a.warmup - 1:      # full-size property: lookup == function of the id
                assert torch.equal(rows_out[:U], row_function(uniq, D, tr.dev)), "looked-up rows differ from f(id)"
                assert torch.equal(bias_out[:U], bias_function(uniq, tr.dev)), "looked-up biases differ from f(id)"
                checked += U
            if it >= a.warmup:
                tot_ms += e0.elapsed_time(e1)
                tot_bytes += lookup_bytes(remote_unique(uniq, rank, world), D)
        t = torch.tensor([tot_ms, float(tot_bytes), float(checked)], device="cuda", dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(mx[0]) / a.steps
        per_gpu = float(t[1]) / world / a.steps
        return {"us_per_lookup": ms * 1e3, "remote_bytes_per_gpu": per_gpu, "gbs_per_dir_per_gpu": per_gpu / (ms * 1e-3) / 1e9,
                "frac_of_nvlink": per_gpu / (ms * 1e-3) / 1e9 / NVLINK_GBS_PER_DIR, "rows_checked": int(t[2])}

    out["lookup_zipf"] = lookup_leg(False)
    out["lookup_uniform"] = lookup_leg(True)

    if not a.no_step:
        ids, counts = synth.glove_batches(V, B, a.nbatch, 17 * rank)
        dev_b = [(torch.from_numpy(ids[k].reshape(-1)).cuda(), torch.from_numpy(counts[k]).cuda()) for k in range(a.nbatch)]
        # property at full size: one step leaves rows outside the (global) batch untouched, accumulators only grow
        all_ids = [torch.empty(2 * B, dtype=torch.int32, device=tr.dev) for _ in range(world)]
        dist.all_gather(all_ids, dev_b[0][0])
        touched = torch.unique(torch.cat(all_ids).to(torch.int64))
        mine = touched[touched % world == rank] // world
        probe = torch.randint(0, tr.shard.V, (1 << 16,), device=tr.dev, generator=torch.Generator(tr.dev).manual_seed(5 + rank))
        keep = torch.ones(tr.shard.V, dtype=torch.bool, device=tr.dev)
        keep[mine] = False
        probe = probe[keep[probe]]
        del keep
        before = tr.shard.rows0[probe].clone()
        tr.step(*dev_b[0])
        torch.cuda.synchronize()
        assert torch.equal(tr.shard.rows0[probe], before), "a row outside the batch changed"
        assert bool((tr.shard.acc[mine] >= 0.1).all()), "an accumulator shrank"
        changed = (tr.shard.rows0[mine] != row_function(mine * world + rank, D, tr.dev)).any(dim=1).float().mean()
        assert float(changed) > 0.5, "most rows inside the batch did not move"
        assert int(tr.err.item()) == 0, "gradient inbox overflow"
        for k in range(a.warmup):
            tr.step(*dev_b[k % a.nbatch])
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(a.steps):
            tr.step(*dev_b[k % a.nbatch])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / a.steps
        out["step"] = {"us_per_step": ms * 1e3, "pairs_per_s": world * B / (ms * 1e-3), "final_loss": float(tr.loss.item()),
                       "rows_outside_batch_checked": int(probe.numel())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
