"""Developer probe: one dump_knn-shaped fused top-k scan (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import engine
V, D = 1_000_000, 128
table = engine.EmbeddingTable(V, D, sparse=False, adagrad=False)
table.rows0.normal_(0.0, 1.0 / D ** 0.5)
q = table.gather(torch.tensor([7, 19, 4000, 1, 100, 33, 2, 5], dtype=torch.int32, device="cuda"))
for _ in range(3):
    engine.table_topk(table, q, 10, ties_high_index_first=True)
torch.cuda.synchronize()
