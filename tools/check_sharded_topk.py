"""2+ GPUs: sharded_table_topk with the real fused scan kernel on every shard == engine.table_topk over the whole table,
both tie conventions (planted exact ties across shards).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_sharded_topk.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import engine  # noqa: E402
from esrecsys_b200.sharded import sharded_query_rows, sharded_table_topk  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
V, D, k = 50001, 128, 10
rng = np.random.default_rng(0)
E = rng.standard_normal((V, D)).astype(np.float32)
E[V - 1] = E[3]
E[V // 2] = E[3]                                            # rows 3, V//2, V-1 tie for every query
tokens = np.array([3, 17, 40000, 5, 123, 9999, 31337, 42], np.int32)
shard = engine.EmbeddingTable.from_dense(np.ascontiguousarray(E[rank::world]), sparse=False)
full = engine.EmbeddingTable.from_dense(E, sparse=False)
q = sharded_query_rows(shard.rows0, tokens, rank, world)
assert torch.equal(q.cpu(), torch.from_numpy(E[tokens])), "query rows"
ok = True
for high_first in (True, False):
    val, rows = sharded_table_topk(shard, q, k, rank, world, ties_high_index_first=high_first)
    wval, widx = engine.table_topk(full, q, k, ties_high_index_first=high_first)
    same = torch.equal(rows.cpu(), widx.long().cpu()) and torch.equal(val.cpu(), wval.cpu())
    ok &= same
    if rank == 0:
        print("ties_high_index_first=%s: %s  (query 0 -> %s)" % (high_first, "identical" if same else "MISMATCH", rows[0, :4].tolist()))
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("check_sharded_topk:", "ok" if int(t.item()) == 1 else "FAILED", "on", world, "GPUs")
dist.destroy_process_group()
