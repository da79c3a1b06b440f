"""Developer probe: per-phase device time of the peer-memory sharded step (run under torchrun)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import _lib as L, synth  # noqa: E402
from esrecsys_b200.sharded import PeerShardedGloveTrainer  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, D, B = 1000000, 128, int(os.environ.get("B", 262144))
    tr = PeerShardedGloveTrainer(V, D, B)
    tr.shard.rows0.normal_(0, 1 / np.sqrt(D))
    ids, counts = synth.glove_batches(V, B, 4, 17 * rank)
    d_ids = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(4)]
    d_cnt = [torch.from_numpy(counts[k]).cuda() for k in range(4)]
    names = ["plan", "route", "gather", "compact", "prep", "ar1", "emitplan", "rows", "ar2", "finish", "bar1", "pull", "merge", "bar2"]
    acc = {n: 0.0 for n in names}
    lib = L.lib()
    steps = 20
    for it in range(steps + 3):
        k = it % 4
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        sp = L.stream_ptr()
        plan, cplan, n, st, pub = tr.plans[0], tr.cplan, tr.n, tr.step_fn, tr.pub[0]
        i = 0
        ev[i].record(); i += 1
        plan.build(d_ids[k]); ev[i].record(); i += 1
        tr.ops.route_plan(plan.uniq, plan.n_uniq, n, out=(pub["order"], pub["send_local"], pub["counts"], pub["inv_order"])); ev[i].record(); i += 1
        L.check(lib.esr_peer_gather_f32(tr.p_rows, tr.p_bias, n, L.ptr(plan.uniq), L.ptr(plan.n_uniq), plan.capacity, tr.D,
                                        L.ptr(tr.compact.rows0), L.ptr(tr.compact.bias), sp)); ev[i].record(); i += 1
        L.check(lib.esr_plan_compact_i32(C.byref(plan.s), L.ptr(cplan.sorted_keys), L.ptr(cplan.partner), L.ptr(cplan.uniq),
                                         L.ptr(tr.scratch), sp))
        cs = cplan.s; cs.n_slots = plan.n_slots
        cs.perm, cs.useg, cs.seg_off, cs.n_uniq = plan.s.perm, plan.s.useg, plan.s.seg_off, plan.s.n_uniq
        ev[i].record(); i += 1
        st.prep(cplan, d_cnt[k]); ev[i].record(); i += 1
        dist.all_reduce(st.scalars[0:3]); ev[i].record(); i += 1
        L.check(lib.esr_peer_emit_plan_i32(pub["p_counts"], n, tr.rank, L.ptr(plan.uniq), L.ptr(plan.n_uniq), plan.capacity,
                                           L.ptr(pub["inv_order"]), tr.inbox_cap, L.ptr(tr.emit_map), L.ptr(tr.err), sp)); ev[i].record(); i += 1
        st.rows(cplan); ev[i].record(); i += 1
        dist.all_reduce(st.scalars[3:5]); ev[i].record(); i += 1
        st.finish(cplan); ev[i].record(); i += 1
        tr.barrier(); ev[i].record(); i += 1
        L.check(lib.esr_peer_pull_ids_i32(pub["p_counts"], pub["p_send_local"], n, tr.rank, tr.recv_cap, L.ptr(tr.recv_ids),
                                          L.ptr(tr.src_meta), L.ptr(tr.slot_map), tr.map_stride, sp)); ev[i].record(); i += 1
        L.check(lib.esr_peer_merge_adagrad_f32(C.byref(tr.shard.struct()), L.ptr(tr.inbox_dE), L.ptr(tr.inbox_db), n, L.ptr(tr.recv_ids),
                                               L.ptr(tr.src_meta), L.ptr(tr.slot_map), tr.map_stride, L.ptr(tr.desc), tr.recv_cap, tr.lr, 1e-7, sp)); ev[i].record(); i += 1
        tr.barrier(); ev[i].record(); i += 1
        torch.cuda.synchronize()
        if it >= 3:
            for j, nme in enumerate(names):
                acc[nme] += ev[j].elapsed_time(ev[j + 1])
    if rank == 0:
        us = {k: round(v / steps * 1e3, 1) for k, v in acc.items()}
        print(json.dumps(dict(world=world, B=B, us=us, total_us=round(sum(us.values()), 1), U=int(tr.plans[0].n_uniq.item()),
                              recv=int(tr.src_meta[3 * world].item()))))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
