"""compute-sanitizer target for the kernels added late in round 2: the wide plan sort (k_sort_hist / k_sort_pass / k_heads,
look-back through polled status words) and the block-cooperative segment sum.
   compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py
   compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import _lib as L, engine  # noqa: E402
from oracle import index as oidx  # noqa: E402

os.environ["ESR_PLAN_SORT"] = "own"
rng = np.random.default_rng(0)
for V, n in ((300, 2050), (10 ** 6, 40000), (10 ** 8, 20002)):
    keys = np.minimum(rng.zipf(1.2, size=n) - 1, V - 1).astype(np.int32)
    plan = engine.IndexPlan(n, V).build(torch.from_numpy(keys).cuda())
    sk, perm, uniq, off = plan.host_view()
    osk, operm = oidx.sort_slots(keys)
    assert np.array_equal(sk, osk) and np.array_equal(perm, operm)
    ou, oo = oidx.segments(osk)
    assert np.array_equal(uniq, ou) and np.array_equal(off, oo)
n, V, D = 6000, 4000, 128
keys = rng.integers(0, V, size=n).astype(np.int32)
keys[:2000] = 77
plan = engine.IndexPlan(n, V, with_partner=False).build(torch.from_numpy(keys).cuda())
g = torch.randn(n, D, device="cuda")
out = torch.zeros(n, D, device="cuda")
L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), D, L.ptr(g), None, L.ptr(out), None, L.stream_ptr()), "segment sum")
torch.cuda.synchronize()
U = int(plan.n_uniq.item())
want = torch.zeros(U, D, device="cuda").index_add_(0, plan.useg[:n].long(), g[plan.perm[:n].long()])
assert torch.allclose(out[:U], want, rtol=1e-4, atol=1e-3)
print("sanitize_new_kernels: ok")
