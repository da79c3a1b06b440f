#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config[1].

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Workload (N = 1): wikipedia GloVe, 1M-vocab x 128-dim table, synthetic Zipf(1) (i, j, count)
triples (esrecsys_b200/synth.py), fused gather -> dot -> loss -> sparse Adagrad scatter, one
"step" = one batch of --batch pairs through the whole hot path (index plan + prep + rows +
combine + finish).  Prints ONE JSON line (rank 0).

* value      : pairs/s, batches resident in HBM before the timed region.
* e2e        : pairs/s through GloveTrainer.submit() from PINNED HOST batches (H2D every step)
               plus a device->host read of the step's loss.
* roofline   : the row-pass kernel's ALGORITHMIC bytes (U*R*4 + U*16 + B*12)
               / its CUDA-event duration, on the timed (Zipf) stream; roofline_uniform is the same
               kernel on the uniform no-reuse control stream, where the >=70 % target is judged.
* cpu_baseline: oracle/glove_torch.py (a port of the reference's dense-Adam step) on the host;
               cpu_baseline_same_algorithm: the sparse-Adagrad rule the GPU path runs, on the host.
* --impl reference: only that CPU leg, as its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairs/sec"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=262144)
    ap.add_argument("--nbatch", type=int, default=8, help="distinct pre-generated batches cycled through")
    ap.add_argument("--kernel", default="auto", choices=["auto", "ldg", "tma", "fifo", "endfirst"])
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=16, help="timed steps of the CPU baseline leg (10-30 s of host work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-uniform", action="store_true")
    ap.add_argument("--no-inbatch", action="store_true")
    ap.add_argument("--depth", type=int, default=3, help="plan/staging buffers in flight (GloveTrainer)")
    ap.add_argument("--row-blocks", type=int, default=-1, help="persistent row-pass grid (-1 = trainer default, 0 = 2 CTAs per SM)")
    ap.add_argument("--stream-priority", action="store_true",
                    help="N=1, experimental: the step's stream at high priority over the plan stream")
    ap.add_argument("--exchange", default="routed", choices=["routed", "peer", "nccl"],
                    help="N>1: owner-computes pair routing over NVLink peer memory (default), the round-1 peer-memory path "
                         "(every rank keeps the pairs it was handed), or NCCL all-to-alls")
    ap.add_argument("--no-table-100m", action="store_true", help="skip the BASELINE configs[4] sub-record (100M-row table)")
    ap.add_argument("--table-rows", type=int, default=100_000_000, help="rows of the configs[4] table")
    ap.add_argument("--fast-sync", action="store_true",
                    help="N>1 peer path, experimental: libesr peer all-reduce / barrier kernels instead of NCCL + symm-mem barrier")
    ap.add_argument("--step-graphs", action="store_true", help="N>1 peer path, experimental: CUDA-graph the sharded step")
    ap.add_argument("--overlap-ids", action="store_true",
                    help="N>1 peer path, experimental: id pull / resolve / emit plan on a side stream next to gather + prep")
    return ap.parse_args()


def workload_name(a):
    return "wikipedia GloVe: %dk-vocab x %d-dim, synthetic Zipf triples, B=%d, fused gather->dot->scatter (BASELINE configs[1])" % (
        a.vocab // 1000, a.dim, a.batch)


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's own step (dense gradient + dense Adam), ported in oracle/glove_torch.py
# ------------------------------------------------------------------------------------------------
def cpu_reference(a, steps, warmup):
    import torch
    from esrecsys_b200 import synth
    from oracle import glove_torch as ogt
    torch.manual_seed(a.seed)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    V, D, B = a.vocab, a.dim, a.batch
    n = max(1, min(a.nbatch, steps + warmup))
    ids, counts = synth.glove_batches(V, B, n, a.seed)
    E = torch.randn(V, D) / np.sqrt(D)
    b = torch.zeros(V)
    st = dict(count=0, muE=torch.zeros_like(E), nuE=torch.zeros_like(E), mub=torch.zeros_like(b), nub=torch.zeros_like(b))
    ti = [torch.from_numpy(ids[k, 0].astype(np.int64)) for k in range(n)]
    tj = [torch.from_numpy(ids[k, 1].astype(np.int64)) for k in range(n)]
    tx = [torch.from_numpy(counts[k]) for k in range(n)]
    for k in range(warmup):
        ogt.step_adam_dense(E, b, st, ti[k % n], tj[k % n], tx[k % n], 1e-3)
    t0 = time.perf_counter()
    for k in range(steps):
        ogt.step_adam_dense(E, b, st, ti[k % n], tj[k % n], tx[k % n], 1e-3)
    dt = time.perf_counter() - t0
    return dict(value=B * steps / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d steps of B=%d on the %dx%d table: oracle/glove_torch.step_adam_dense (dense grad + optax.adam "
                       "over all rows, as wikipedia/train_cooccurence.py:71-101), torch CPU ops, %d threads"
                       % (steps, B, V, D, cores)), dt / steps * 1e3


def cpu_same_algorithm(a, steps, warmup):
    """SURVEY.md 8(d) baseline (S): the north-star rule itself (sparse Adagrad on the touched rows only) on the host,
    oracle/glove_torch.step_adagrad_sparse, multi-threaded torch CPU ops."""
    import torch
    from esrecsys_b200 import synth
    from oracle import glove_torch as ogt
    torch.manual_seed(a.seed)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    V, D, B = a.vocab, a.dim, a.batch
    n = max(1, min(a.nbatch, steps + warmup))
    ids, counts = synth.glove_batches(V, B, n, a.seed)
    E = torch.randn(V, D) / np.sqrt(D)
    b = torch.zeros(V)
    accE, accb = torch.full_like(E, 0.1), torch.full_like(b, 0.1)
    ti = [torch.from_numpy(ids[k, 0].astype(np.int64)) for k in range(n)]
    tj = [torch.from_numpy(ids[k, 1].astype(np.int64)) for k in range(n)]
    tx = [torch.from_numpy(counts[k]) for k in range(n)]
    for k in range(warmup):
        ogt.step_adagrad_sparse(E, b, accE, accb, ti[k % n], tj[k % n], tx[k % n], a.lr)
    t0 = time.perf_counter()
    for k in range(steps):
        ogt.step_adagrad_sparse(E, b, accE, accb, ti[k % n], tj[k % n], tx[k % n], a.lr)
    dt = time.perf_counter() - t0
    return dict(value=B * steps / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d steps of B=%d on the %dx%d table: oracle/glove_torch.step_adagrad_sparse (unique rows, index_add "
                       "segment sums, Adagrad on the touched rows only), torch CPU ops, %d threads" % (steps, B, V, D, cores))


def config_single(a, table_gb=None):
    """``config`` of the N = 1 workload (shared by both arms so that the driver's same-config check holds)."""
    V, D, B = a.vocab, a.dim, a.batch
    return {"workload": workload_name(a), "vocab": V, "dim": D, "batch": B, "optimizer": "sparse adagrad (north star)",
            "bias_mode": "reference_broadcast", "kernel": a.kernel, "stream": "zipf(1)",
            **({"stream_priority": True} if a.stream_priority else {}),
            "l2": "no flush: table + state = %.2f GB per GPU >> 126 MB L2, fresh random rows every step"
                  % (table_gb if table_gb is not None else (3 * V * D * 4 + V * 9) / 1e9),
            "parallelism": "single"}


PARALLELISM = {
    "routed": "row-sharded table (cyclic), dp%d over pairs, OWNER-COMPUTES: every pair is routed to the rank owning row i "
              "(16 B per pair), only the unique partner rows and their gradients cross NVLink (libesr peer-memory "
              "kernels, no NCCL inside the step, CUDA graphs)",
    "peer": "row-sharded table (cyclic), dp%d over pairs; rows fetched and gradients merged by libesr kernels over "
            "NVLink peer memory",
    "nccl": "row-sharded table (cyclic), dp%d over pairs; NCCL all-to-all of ids / rows / gradients"}


def config_sharded(a, world):
    V, D, B = a.vocab, a.dim, a.batch
    return {"workload": workload_name(a) + "; table row-sharded (cyclic) over %d GPUs, B per GPU" % world, "vocab": V,
            "dim": D, "batch_per_gpu": B, "global_batch": B * world, "optimizer": "sparse adagrad (north star)",
            "bias_mode": "reference_broadcast", "stream": "zipf(1)", "exchange": a.exchange,
            "l2": "no flush: per-step working set (fetched rows + shard rows + state) >> 126 MB L2",
            "parallelism": PARALLELISM[a.exchange] % world}


def run_reference(a):
    """Reference arm: the reference's own step (dense gradient + dense optax.adam over every row,
    wikipedia/train_cooccurence.py:71-101) on the host cores -- the oracle port, since jax / flax cannot be installed here --
    on OUR arm's config, EXACTLY the requested steps (bounded at 60: ~0.3 s per step) after the requested warm-up (bounded
    at 5).  N > 1: rank 0 alone runs, one B-pair batch per step (the per-GPU share of the global batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    steps = max(1, min(a.steps, 60))
    warm = max(0, min(a.warmup, 5))
    cb, ms = cpu_reference(a, steps, warm)
    cfg = config_single(a) if world <= 1 else config_sharded(a, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "note": "CPU-only reference (jax / flax absent): oracle/glove_torch.step_adam_dense on all host threads; "
                "steps bounded at 60, warm-up at 5",
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks: sampled with NVML during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # pragma: no cover
            self.nv = None
        self.active = False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self.stop:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                        nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # pragma: no cover
                    pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self.th.start()

    def summary(self):
        self.stop = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------
def alg_bytes(ids_np, D, B):
    """bytes_step = U*R*(2+2S) + U_b*4*(2+2S) + B*12 with S = 1 (Adagrad)."""
    U = float(np.mean([np.unique(ids_np[k]).size for k in range(ids_np.shape[0])]))
    R = D * 4
    return U * R * 4 + U * 16 + B * 12, U


def timed_region(tr, batches, steps, warmup, read_loss, world):
    import torch
    import torch.distributed as dist
    n = len(batches)
    for k in range(warmup):
        tr.submit(*batches[k % n])
    tr.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.s_side.wait_stream(tr.s_main)
    e0.record(tr.s_main)
    tr.s_side.wait_event(e0)
    tr.s_copy.wait_event(e0)                  # the first upload of the timed region starts inside it
    for k in range(steps):
        tr.submit(*batches[(warmup + k) % n], read_loss=read_loss)
    e1.record(tr.s_main)
    tr.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def timed_region_sharded(tr, batches, steps, warmup, pinned_loss, world):
    """N > 1: one sharded trainer step per batch.  The routed trainer runs on its own streams: the timing events on the
    current stream bracket them explicitly."""
    import torch
    import torch.distributed as dist
    n = len(batches)
    own = [getattr(tr, s) for s in ("s_main", "s_side") if hasattr(tr, s)] if hasattr(tr, "s_main") else []
    for k in range(warmup):
        tr.step(*batches[k % n])
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in own:
        s.wait_event(e0)
    for k in range(steps):
        loss = tr.step(*batches[(warmup + k) % n])
        if pinned_loss is not None and not hasattr(tr, "loss_host"):
            # (the owner-routed trainer's finish kernel writes every step's loss into pinned host memory itself:
            # tr.loss_host, 4 bytes per step over PCIe -- checked after the timed region)
            slot = pinned_loss[k % pinned_loss.numel(): k % pinned_loss.numel() + 1]
            if hasattr(tr, "read_loss_to"):
                tr.read_loss_to(slot)
            else:
                slot.copy_(loss.reshape(1), non_blocking=True)
    for s in own:
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    if pinned_loss is not None and hasattr(tr, "loss_host"):
        last = tr.host_loss(tr.t - 1)           # the host really holds the losses of the timed steps
        if not (last == float(tr.loss.item()) and np.isfinite(last)):
            raise RuntimeError("host loss mirror %r != device loss %r" % (last, float(tr.loss.item())))
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t.item())


def make_sharded_trainer(a, V, D, B):
    from esrecsys_b200.sharded import OwnerRoutedGloveTrainer, PeerShardedGloveTrainer, ShardedGloveTrainer
    if a.exchange == "routed":
        kw = {"impl": a.kernel} if a.kernel != "auto" else {}
        if a.row_blocks >= 0:
            kw["row_blocks"] = a.row_blocks
        return OwnerRoutedGloveTrainer(V, D, B, lr=a.lr, **kw)
    if a.exchange == "peer":
        kw = {"fast_sync": a.fast_sync, "graphs": a.step_graphs, "overlap_ids": a.overlap_ids}
        if a.kernel != "auto":
            kw["impl"] = a.kernel
        return PeerShardedGloveTrainer(V, D, B, lr=a.lr, **kw)
    return ShardedGloveTrainer(V, D, B, lr=a.lr)


def sharded_parity_check(a, rank, world):
    """Before the timed region of every N > 1 run: the SAME trainer class on a small table -- 6 global steps (the last ones
    CUDA-graph replays) -- against the single-table engine on the concatenated batch (itself held to the oracle by
    tests/): losses and the re-assembled table.  Semantics: wikipedia/train_cooccurence.py:71-101."""
    import torch
    from esrecsys_b200 import engine, synth
    V, D, B, steps = 50_000, a.dim, 4096, 6
    E, b = synth.init_glove_tables(V, D, 3)
    b = (np.random.default_rng(4).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = synth.glove_batches(V, B * world, steps, 123)
    tr = make_sharded_trainer(a, V, D, B)
    tr.load_dense(E, b)
    table = engine.EmbeddingTable.from_dense(E, b, sparse=True)
    single = engine.GloveStep(table, B * world, lr=a.lr)
    plan = engine.IndexPlan(2 * B * world, V)
    lo, hi = rank * B, (rank + 1) * B
    dl = 0.0
    for k in range(steps):
        loss = tr.step(torch.from_numpy(np.ascontiguousarray(ids[k][:, lo:hi])).cuda(), torch.from_numpy(counts[k][lo:hi]).cuda())
        plan.build(torch.from_numpy(ids[k].reshape(-1)).cuda())
        ref = single.run(plan, torch.from_numpy(counts[k]).cuda())[engine.L.SC_LOSS]
        got = tr.loss_value() if hasattr(tr, "loss_value") else float(loss.item())
        dl = max(dl, abs(got - float(ref.item())) / max(1e-12, abs(float(ref.item()))))
    Eg, bg = tr.gather_dense()
    Es, bs = table.dense(), table.bias
    dE = float((Eg - Es).abs().max().item())
    db = float((bg - bs).abs().max().item())
    tol = float(((Eg - Es).abs() - 1e-5 * Es.abs()).max().item())
    ok = bool(dl <= 2e-5 and tol <= 1e-5 and db <= 1e-5)
    del tr, table, single, plan
    torch.cuda.empty_cache()
    return {"ok": ok, "vs": "single-table libesr engine on the concatenated global batch", "steps": steps, "vocab": V,
            "batch_per_gpu": B, "max_rel_loss_diff": dl, "max_abs_row_diff": dE, "max_abs_bias_diff": db,
            "tolerance": "1e-5 abs + 1e-5 rel (rows, bias), 2e-5 rel (loss)"}


def sharded_timed(a, tr, V, B, rank, world, steps, warmup, seed):
    """value / e2e legs of one sharded trainer; returns (ms, ms_e2e, U_local_mean)."""
    import torch
    from esrecsys_b200 import synth
    warmup = max(warmup, 5)                   # the routed trainer captures its CUDA graphs at its 5th step
    ids, counts = synth.glove_batches(V, B, a.nbatch, seed + 17 * rank)
    dev_b = [(torch.from_numpy(ids[k]).cuda().reshape(-1), torch.from_numpy(counts[k]).cuda()) for k in range(a.nbatch)]
    pin_b = [(torch.from_numpy(ids[k]).pin_memory(), torch.from_numpy(counts[k]).pin_memory()) for k in range(a.nbatch)]
    ms = timed_region_sharded(tr, dev_b, steps, warmup, None, world)
    pinned_loss = torch.zeros(64).pin_memory()
    ms_e2e = timed_region_sharded(tr, pin_b, steps, max(3, warmup // 4), pinned_loss, world)
    if hasattr(tr, "check"):
        tr.check()
    return ms, ms_e2e


def nvlink_accounting(tr, world):
    """Rows that crossed NVLink in the LAST step on this rank (per direction: fetched rows in == gradient rows out), from
    the route plan the step published (device counters)."""
    import torch
    try:
        k = (tr.t - 1) % tr.DEPTH
        counts = tr.pub[k]["counts"][:world].cpu().numpy()
        remote = int(counts.sum() - counts[tr.rank])
        return remote, int(counts.sum())
    except Exception:
        return None, None


def run_sharded(a, rank, world, local):
    """N > 1: the table row-shards cyclically over the ranks (weak scaling: --batch pairs per GPU)."""
    import torch
    import torch.distributed as dist
    V, D, B = a.vocab, a.dim, a.batch
    torch.manual_seed(a.seed)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    parity = sharded_parity_check(a, rank, world)
    tr = make_sharded_trainer(a, V, D, B)
    tr.shard.rows0.normal_(0.0, 1.0 / np.sqrt(D))
    clocks = ClockSampler(local)
    clocks.start()
    clocks.active = True
    ms, ms_e2e = sharded_timed(a, tr, V, B, rank, world, a.steps, a.warmup, a.seed)
    clocks.active = False
    remote_rows, uniq_rows = nvlink_accounting(tr, world)
    R = 4 * D
    par = {"routed": "row-sharded table (cyclic), dp%d over pairs, OWNER-COMPUTES: every pair is routed to the rank owning row i "
                     "(12 B per pair), only the unique partner rows and their gradients cross NVLink (libesr peer-memory "
                     "kernels, no NCCL inside the step, CUDA graphs)" % world,
           "peer": "row-sharded table (cyclic), dp%d over pairs; rows fetched and gradients merged by libesr kernels over "
                   "NVLink peer memory" % world,
           "nccl": "NCCL all-to-all of ids / rows / gradients"}[a.exchange]
    line = {
        "metric": METRIC, "value": world * B * a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_sharded(a, world),
        "e2e": {"value": world * B * a.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 12 * B,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": a.steps * tr.LAUNCHES_PER_STEP, "final_loss": float(tr.loss.item()),
        "parity_check": parity,
        "clocks": clocks.summary(),
    }
    if uniq_rows is not None:
        # whole-step roofline of one rank: the algorithmic HBM bytes of SURVEY.md 8(d) for the rows this rank's pairs
        # touch, and the NVLink bytes that had to cross per direction, both over the step time
        alg = uniq_rows * R * 4 + uniq_rows * 16 + B * 12
        nv = remote_rows * (R + 8)
        t_s = ms / a.steps * 1e-3
        line["roofline"] = {"bound": "hbm", "achieved": alg / t_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg / t_s / 1e9 / hbm_peak, "traffic": None, "kernel": "whole sharded step (rank 0)",
                            "alg_bytes_per_launch": alg, "unique_rows_per_step": uniq_rows,
                            "nvlink": {"remote_rows_per_step": remote_rows, "bytes_per_dir_per_step": nv,
                                       "achieved_gbs_per_dir": nv / t_s / 1e9, "peak_gbs_per_dir": 900.0,
                                       "frac": nv / t_s / 1e9 / 900.0,
                                       "note": "not NVLink-bound by design: owner-computes routing leaves ~1/4 of the rows "
                                               "on the wire"}}
    del tr
    torch.cuda.empty_cache()
    wd = Watchdog(line, rank, 420.0)
    if not a.no_table_100m:
        wd.pending = "table_100m"
        line["table_100m"] = guarded(lambda: table_100m_leg(a, rank, world), "table_100m")
    if not a.no_inbatch:
        wd.pending = "inbatch_sharded"
        line["inbatch_sharded"] = guarded(lambda: inbatch_sharded_leg(a, rank, world), "inbatch_sharded")
    wd.pending = None
    wd.cancel()
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


class Watchdog:
    """Sub-records (100M-row table, sharded in-batch trainers) run AFTER the headline numbers are final.  If one of them
    hangs (a collective that never completes), every rank leaves on its own after ``deadline_s``: rank 0 prints the line it
    has -- the sub-record marked as timed out -- so the headline measurement is never lost with it."""

    def __init__(self, line, rank, deadline_s):
        self.line, self.rank, self.pending = line, rank, None
        self.t = threading.Timer(deadline_s, self._fire)
        self.t.daemon = True
        self.t.start()

    def _fire(self):  # pragma: no cover
        if self.rank == 0:
            if self.pending:
                self.line[self.pending] = {"error": "timed out (watchdog); headline numbers above are unaffected"}
            print(json.dumps(self.line), flush=True)
        os._exit(0)

    def cancel(self):
        self.t.cancel()


def inbatch_sharded_leg(a, rank, world):
    """BASELINE configs[2] across the ranks: 2M-row x 128 table row-sharded (cyclic), GLOBAL in-batch batch 8192 (every
    query scored against the items of all ranks: all-gather of K, tcgen05 scores, reduce-scatter of dK), hinge loss,
    sparse Adagrad at the owners.  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    from esrecsys_b200 import synth
    from esrecsys_b200.inbatch import ShardedSharedTableInBatch
    V, D, Bg = 2_000_000, 128, 8192
    B = Bg // world
    tr = ShardedSharedTableInBatch(V, D, B, lr=0.05, loss="hinge")
    tr.shard.rows0.normal_(0.0, 1.0 / D ** 0.25)
    qs, ks = synth.pair_batches(V, V, Bg, 4, 5)
    lo, hi = rank * B, (rank + 1) * B
    batches = [torch.from_numpy(np.stack([qs[k][lo:hi], ks[k][lo:hi]])).cuda() for k in range(4)]
    steps, warm = 20, 5
    for k in range(warm):
        tr.step(batches[k % 4])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        loss = tr.step(batches[k % 4])
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"workload": "BASELINE configs[2]: 2M x 128 table row-sharded over %d GPUs, global in-batch B=8192, hinge" % world,
            "ms_per_step": ms, "pairs_per_s": Bg / (ms * 1e-3), "final_loss": float(loss.item()),
            "useful_tflops_all_ranks": 6.0 * Bg * Bg * D / (ms * 1e-3) / 1e12,
            "exchange": "NCCL all-to-all of ids / rows / gradients + all-gather K + reduce-scatter dK"}


def guarded(fn, name):
    """Sub-records never take the headline line down with them."""
    try:
        return fn()
    except Exception as e:  # pragma: no cover
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def table_100m_leg(a, rank, world):
    """BASELINE configs[4]: the same step on a 100M-row x 128 table (51.2 GB + Adagrad state), Zipf(1) ids over all 100M
    rows, B pairs per GPU.  N = 1: the single-table engine (3 x 51.2 GB resident); N > 1: row-sharded, owner-computes.
    Reports pairs/s and, for N > 1, the looked-up rows per step and the NVLink rate of the lookup kernel timed alone."""
    import torch
    import torch.distributed as dist
    from esrecsys_b200 import _lib as L, engine, synth
    import ctypes as C
    V, D, B = a.table_rows, a.dim, a.batch
    steps, warm = 20, 5
    out = {"workload": "BASELINE configs[4]: %dM rows x %d, Zipf(1) over all rows, B=%d per GPU" % (V // 1_000_000, D, B),
           "n_gpus": world, "vocab": V}
    if world == 1:
        from esrecsys_b200.trainer import GloveTrainer
        table = engine.EmbeddingTable(V, D)
        table.rows0.normal_(0.0, 1.0 / np.sqrt(D))
        ids, counts = synth.glove_batches(V, B, 4, a.seed + 5)
        dev_b = [(torch.from_numpy(ids[k].reshape(-1)).cuda(), torch.from_numpy(counts[k]).cuda()) for k in range(4)]
        tr = GloveTrainer(table, B, lr=a.lr)
        ms = timed_region(tr, dev_b, steps, warm, False, 1)
        out.update({"ms_per_step": ms / steps, "pairs_per_s": B * steps / (ms * 1e-3), "table_gb": table.nbytes() / 1e9,
                    "unique_rows_per_step": float(np.mean([np.unique(ids[k]).size for k in range(4)])),
                    "final_loss": float(tr.losses(tr.t - 1, tr.t)[0])})
        return out
    tr = make_sharded_trainer(a, V, D, B)
    tr.shard.rows0.normal_(0.0, 1.0 / np.sqrt(D))
    ms, ms_e2e = sharded_timed(a, tr, V, B, rank, world, steps, warm, a.seed + 5)
    remote_rows, uniq_rows = nvlink_accounting(tr, world)
    out.update({"ms_per_step": ms / steps, "pairs_per_s": world * B * steps / (ms * 1e-3),
                "e2e_pairs_per_s": world * B * steps / (ms_e2e * 1e-3), "shard_gb": tr.shard.V * D * 4 / 1e9,
                "final_loss": float(tr.loss.item()), "exchange": a.exchange})
    if uniq_rows is not None and hasattr(tr, "fetch_rows"):
        # the lookup alone (esr_peer_gather_f32 of the last step's unique rows), all ranks at once: NVLink GB/s per direction
        k = (tr.t - 1) % tr.DEPTH
        plan = tr.plans[k]
        torch.cuda.synchronize()
        dist.barrier()
        tot = 0.0
        for it in range(6):
            tr.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pub = tr.pub[k]
            L.check(L.lib().esr_peer_gather_remote_f32(tr.p_rows, tr.p_bias, world, tr.rank, L.ptr(plan.uniq), L.ptr(pub["order"]),
                                                       L.ptr(pub["counts"]), plan.capacity, D, L.ptr(tr.fetch_rows),
                                                       L.ptr(tr.fetch_bias), 3, L.stream_ptr()), "esr_peer_gather_remote_f32")
            e1.record()
            torch.cuda.synchronize()
            if it >= 1:
                tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / 5], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item()) * 1e3
        nv = remote_rows * (4 * D + 8)
        out["lookup"] = {"unique_rows": uniq_rows, "remote_rows": remote_rows, "us": us, "bytes_per_dir": nv,
                         "gbs_per_dir": nv / (us * 1e-6) / 1e9, "nvlink_peak_gbs_per_dir": 900.0,
                         "frac_of_nvlink": nv / (us * 1e-6) / 1e9 / 900.0,
                         "note": "rank 0's numbers; time = max over ranks, all ranks looking up at once"}
    return out


def rows_kernel_time(table, a, ids_dev, cnt_dev, steps):
    """Average CUDA-event duration of the row-pass kernel alone (eager launches, same batches)."""
    import torch
    from esrecsys_b200 import engine
    step = engine.GloveStep(table, a.batch, lr=a.lr, impl=a.kernel)
    plan = engine.IndexPlan(2 * a.batch, a.vocab)
    n = len(ids_dev)
    tot = 0.0
    for k in range(steps + 3):
        plan.build(ids_dev[k % n])
        step.prep(plan, cnt_dev[k % n])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step.rows_main(plan)
        e1.record()
        step.rows_combine(plan)
        step.finish(plan)
        torch.cuda.synchronize()
        if k >= 3:
            tot += e0.elapsed_time(e1)
    return tot / steps


def ncu_traffic(stream, a):
    """DRAM bytes per row-pass launch from the committed ncu capture of the same kernel / batch (profiles/r2_traffic.json,
    captured with tools/r2_evidence.sh on the round-2 kernel; None for any other shape)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[stream]
        if t["batch"] == a.batch and a.vocab == 1_000_000 and a.dim == 128:
            return t["bytes"]
    except Exception:
        pass
    return None


def inbatch_leg(peaks):
    """BASELINE configs[2] / configs[3] scoring step (B x B in-batch negatives on tcgen05): device-timed, inputs
    resident.  flops = 6*B*B*D useful (one score pass + dQ + dK); softmax issues a second score pass."""
    import torch
    from esrecsys_b200.engine import InBatchScorer
    peak = float(peaks.get("bf16_tflops", 1590.0))
    out = {}
    for name, B, D, loss in (("spotify_inbatch_hinge_B8192_D128", 8192, 128, "hinge"),
                             ("spotify_inbatch_softmax_B8192_D128", 8192, 128, "softmax"),
                             ("two_tower_inbatch_softmax_B4096_D256", 4096, 256, "softmax")):
        g = torch.Generator(device="cuda").manual_seed(1)
        Q = torch.randn(B, D, device="cuda", generator=g) / D ** 0.25
        K = torch.randn(B, D, device="cuda", generator=g) / D ** 0.25
        sc = InBatchScorer(B, D, loss=loss)
        for _ in range(5):
            sc.run(Q, K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            sc.run(Q, K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 6.0 * B * B * D / (ms * 1e-3) / 1e12
        out[name] = {"ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3), "useful_tflops": tf, "bound": "tensor",
                     "peak": peak, "frac": tf / peak, "dtype": "bf16 operands, f32 accumulate (tcgen05)",
                     "kernels_per_step": 4 if loss == "hinge" else 6}
        del sc
    out.update(guarded(inbatch_trainer_steps, "inbatch_trainers"))
    return out


def inbatch_trainer_steps():
    """The FULL trainer step of configs[2] / configs[3] at one GPU (not just the scorer): id gather -> [MLP towers] ->
    tcgen05 B x B scores + loss + dQ / dK -> per-row gradient sums -> sparse Adagrad on the table(s) [+ Adam on the towers]."""
    import torch
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import SharedTableInBatch, TwoTowerInBatch
    out = {}

    def timeit(fn, n=30):
        for k in range(5):
            fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            fn(k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    V, D, B = 2_000_000, 128, 8192
    table = engine.EmbeddingTable(V, D, sparse=False, adagrad=True)
    table.rows0.normal_(0.0, 1.0 / D ** 0.25)
    qs, ks = synth.pair_batches(V, V, B, 4, 5)
    ids = [torch.from_numpy(np.stack([qs[k], ks[k]])).cuda() for k in range(4)]
    for loss in ("hinge", "softmax"):
        tr = SharedTableInBatch(table, B, lr=0.05, loss=loss)
        ms_eager = timeit(lambda k: tr.step(ids[k % 4]))
        g = tr.graphed(ids[0])
        ms = timeit(lambda k: g(ids[k % 4]))
        out["spotify_inbatch_%s_TRAINER_step_V2M_D128_B8192" % loss] = {
            "ms_per_step": ms, "ms_per_step_eager_launches": ms_eager, "pairs_per_s": B / (ms * 1e-3),
            "useful_tflops": 6.0 * B * B * D / (ms * 1e-3) / 1e12, "cuda_graph": True,
            "includes": "gather 2B rows, scores + loss + dQ/dK (tcgen05), segment sum of 2B gradient rows, sparse Adagrad"}
        del tr
    del table
    torch.cuda.empty_cache()
    V, D, B = 1_000_000, 256, 4096
    ts = engine.EmbeddingTable(V, D, sparse=False, adagrad=True)
    tp = engine.EmbeddingTable(V, D, sparse=False, adagrad=True)
    ts.rows0.normal_(0.0, 1.0 / D ** 0.5)
    tp.rows0.normal_(0.0, 1.0 / D ** 0.5)
    qs, ks = synth.pair_batches(V, V, B, 4, 6)
    sid = [torch.from_numpy(qs[k]).cuda() for k in range(4)]
    pid = [torch.from_numpy(ks[k]).cuda() for k in range(4)]
    for mm in ("fp32", "tf32"):
        tr = TwoTowerInBatch(ts, tp, B, loss="softmax", tower_matmul=mm)
        ms_eager = timeit(lambda k: tr.step(sid[k % 4], pid[k % 4]))
        g = tr.graphed(sid[0], pid[0])
        ms = timeit(lambda k: g(sid[k % 4], pid[k % 4]))
        out["two_tower_softmax_TRAINER_step_V1M_D256_B4096" + ("" if mm == "fp32" else "_tf32_towers")] = {
            "ms_per_step": ms, "ms_per_step_eager_launches": ms_eager, "pairs_per_s": B / (ms * 1e-3), "cuda_graph": True,
            "tower_matmul": mm,
            "includes": "2 id gathers, 2 x (Linear-ReLU-Linear) towers fwd + bwd + Adam (cuBLAS GEMMs through torch: 48 % of "
                        "the step's kernel time at fp32, profiles/r2_twotower_launches.txt), scores + loss + dQ/dK "
                        "(tcgen05), sparse Adagrad on both tables"}
        del tr, g
    return out


def retrieval_leg(peaks, table, a):
    """the fused score + top-k scan (esr_topk_scan_f32).  (1) dump_knn's shape on the bench table
    (wikipedia/train_cooccurence.py:114-126: T = 8 queries, k = 10): HBM-bound, V*D*4 bytes per call; (2) eval_step's
    shape (spotify/train_spotify.py:113-131: 2 262 292 tracks gathered from the 100000 x 32 album and 295861 x 32 artist
    tables, 5 context rows, k = 500): the tables sit in L2, the stream is the 8 bytes of ids per track."""
    import torch
    from esrecsys_b200 import engine
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {}

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    V, D = table.V, table.D
    q = table.gather(torch.tensor([7, 19, 4000, 1, 100, 33, 2, 5], dtype=torch.int32, device="cuda"))
    ms = timeit(lambda: engine.table_topk(table, q, 10, ties_high_index_first=True))
    gbs = V * D * 4 / (ms * 1e-3) / 1e9
    out["dump_knn_topk_V%dk_D%d_T8_k10" % (V // 1000, D)] = {
        "ms": ms, "bound": "hbm", "alg_bytes": V * D * 4, "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
        "note": "one table pass, running top-10 per query in shared memory; round 1: (V,8) score matrix + 8 full radix sorts"}
    N, F = 2_262_292, 32
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(100_000, F, device="cuda", generator=g) / F ** 0.5
    R = torch.randn(295_861, F, device="cuda", generator=g) / F ** 0.5
    alb = torch.randint(0, 734_684, (N,), device="cuda", generator=g, dtype=torch.int32)
    art = torch.randint(0, 295_861, (N,), device="cuda", generator=g, dtype=torch.int32)
    ctx_alb, ctx_art = alb[:5].clone(), art[:5].clone()
    ctx = torch.cat([A[(ctx_alb % 100_000).long()], R[ctx_art.long()]], dim=1).contiguous()
    ms = timeit(lambda: engine.topk_scan(A, ctx, 500, rows_b=R, idx_a=alb, idx_b=art, mod_a=100_000, max_over_queries=True,
                                         ctx_a=ctx_alb, ctx_b=ctx_art, boost=0.1))
    out["spotify_eval_top500_N2262292_F32"] = {
        "ms": ms, "tracks_per_s": N / (ms * 1e-3), "bound": "l2 / issue (embedding tables are L2-resident)",
        "id_bytes": 8 * N, "gathered_row_bytes": N * 2 * F * 4,
        "note": "fused gather + 5 dots + max + isin boosts + running top-500; round 1: (N,64) candidate matrix + (N,5) "
                "scores + full radix sort of N keys"}
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    from esrecsys_b200 import _lib, engine, synth
    from esrecsys_b200.trainer import GloveTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    _lib.lib()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return run_sharded(a, rank, world, local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    V, D, B = a.vocab, a.dim, a.batch
    torch.manual_seed(a.seed + rank)
    table = engine.EmbeddingTable(V, D)
    table.rows0.normal_(0.0, 1.0 / np.sqrt(D))        # flax nn.Embed default init (wikipedia/models.py:16-17)
    ids, counts = synth.glove_batches(V, B, a.nbatch, a.seed + 17 * rank)
    ids_dev = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(a.nbatch)]
    cnt_dev = [torch.from_numpy(counts[k]).cuda() for k in range(a.nbatch)]
    tr = GloveTrainer(table, B, lr=a.lr, impl=a.kernel, depth=a.depth, row_blocks=None if a.row_blocks < 0 else a.row_blocks,
                      priorities=a.stream_priority)
    ids_pin, cnt_pin = [], []
    for k in range(a.nbatch):        # host batches as a loader would leave them: (2,B) int32 + (B,) f32 in one pinned block
        hi, hc = tr.pinned_batch()
        hi.copy_(torch.from_numpy(ids[k]))
        hc.copy_(torch.from_numpy(counts[k]))
        ids_pin.append(hi)
        cnt_pin.append(hc)

    clocks = ClockSampler(local)
    clocks.start()
    clocks.active = True
    # --- value: device-resident batches ---
    ms = timed_region(tr, list(zip(ids_dev, cnt_dev)), a.steps, a.warmup, False, world)
    # --- e2e: pinned host batches, H2D + loss D2H every step ---
    ms_e2e = timed_region(tr, list(zip(ids_pin, cnt_pin)), a.steps, max(3, a.warmup // 4), True, world)
    loss = float(tr.losses(tr.t - 1, tr.t)[0])
    clocks.active = False
    value = world * B * a.steps / (ms * 1e-3)
    e2e = world * B * a.steps / (ms_e2e * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_single(a, table.nbytes() / 1e9),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": a.steps * tr.launches_per_step,
        "final_loss": loss,
    }
    if rank == 0:
        # --- roofline of the dominant kernel, Zipf (timed stream) and uniform control ---
        clocks.active = True
        ksteps = max(5, min(a.steps, 50))
        t_rows = rows_kernel_time(table, a, ids_dev, cnt_dev, ksteps)
        ab, U = alg_bytes(ids, D, B)
        ach = ab / (t_rows * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                            "traffic": ncu_traffic("zipf", a), "kernel": "k_glove_rows_grp_async", "kernel_ms": t_rows,
                            "alg_bytes_per_launch": ab, "unique_rows_per_step": U, "peak_source": peak_src,
                            "stream": "zipf(1)"}
        if not a.no_uniform:
            uids, ucnt = synth.glove_batches(V, B, 4, a.seed + 99, uniform=True)
            u_dev = [torch.from_numpy(uids[k].reshape(-1)).cuda() for k in range(4)]
            uc_dev = [torch.from_numpy(ucnt[k]).cuda() for k in range(4)]
            t_u = rows_kernel_time(table, a, u_dev, uc_dev, ksteps)
            abu, Uu = alg_bytes(uids, D, B)
            achu = abu / (t_u * 1e-3) / 1e9
            line["roofline_uniform"] = {"bound": "hbm", "achieved": achu, "peak": hbm_peak, "unit": "GB/s",
                                        "frac": achu / hbm_peak, "traffic": ncu_traffic("uniform", a),
                                        "kernel": "k_glove_rows_grp_async", "kernel_ms": t_u,
                                        "alg_bytes_per_launch": abu, "unique_rows_per_step": Uu, "stream": "uniform"}
        if not a.no_inbatch:
            line["other_workloads"] = guarded(lambda: inbatch_leg(peaks), "inbatch")
            line["retrieval"] = guarded(lambda: retrieval_leg(peaks, table, a), "retrieval")
        clocks.active = False
    line["clocks"] = clocks.summary()
    if rank == 0:
        del tr, table, ids_dev, cnt_dev
        torch.cuda.empty_cache()
        if world == 1 and not a.no_table_100m:
            line["table_100m"] = guarded(lambda: table_100m_leg(a, 0, 1), "table_100m")
            torch.cuda.empty_cache()
        if world == 1 and not a.no_cpu:
            cb, _ = cpu_reference(a, a.cpu_steps, 1)
            line["cpu_baseline"] = cb
            line["cpu_baseline_same_algorithm"] = cpu_same_algorithm(a, 2 * a.cpu_steps, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
