"""Developer probe: random-row gather throughput from ordinary vs symmetric memory (1 process)."""
import ctypes as C, os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import _lib as L, engine

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29514")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
    import torch.distributed._symmetric_memory as symm_mem
    V, D, U, cap = 1000000, 128, 137024, 524288
    lib = L.lib()
    rng = np.random.default_rng(0)
    uniq = np.sort(rng.choice(V, U, replace=False)).astype(np.int32)
    u_t = torch.zeros(cap, dtype=torch.int32, device="cuda"); u_t[:U] = torch.from_numpy(uniq).cuda()
    nu = torch.tensor([U], dtype=torch.int32, device="cuda")
    normal = torch.randn(V, D, device="cuda"); nbias = torch.randn(V, device="cuda")
    st = symm_mem.empty((V, D), dtype=torch.float32, device=torch.device("cuda", 0)); h = symm_mem.rendezvous(st, dist.group.WORLD.group_name)
    sb = symm_mem.empty((V,), dtype=torch.float32, device=torch.device("cuda", 0)); hb = symm_mem.rendezvous(sb, dist.group.WORLD.group_name)
    st.copy_(normal); sb.copy_(nbias)
    out = torch.empty(cap, D, device="cuda"); ob = torch.empty(cap, device="cuda")
    res = {}
    t = engine.EmbeddingTable.wrap(normal)
    res["table_gather_normal_us"] = timeit(lambda: t.gather(u_t[:U], out=out[:U]))
    for name, rows, bias in (("normal", normal, nbias), ("symm", st, sb)):
        pr = (C.c_void_p * 8)(rows.data_ptr()); pb = (C.c_void_p * 8)(bias.data_ptr())
        for capx in (cap, U):
            res["peer_gather_%s_cap%d_us" % (name, capx)] = timeit(lambda: L.check(lib.esr_peer_gather_f32(
                pr, pb, 1, L.ptr(u_t), L.ptr(nu), capx, D, L.ptr(out), L.ptr(ob), L.stream_ptr())))
    res["copy_70MB_us"] = timeit(lambda: out[:U].copy_(normal[:U]))
    # in-place sparse adagrad on U rows: table rows in normal vs symmetric memory, grads normal vs symmetric
    g_n = torch.randn(cap, D, device="cuda")
    sg = symm_mem.empty((cap, D), dtype=torch.float32, device=torch.device("cuda", 0)); hg = symm_mem.rendezvous(sg, dist.group.WORLD.group_name)
    sg.copy_(g_n)
    acc = torch.full((V, D), 0.1, device="cuda")
    for name, rows in (("normal", normal), ("symm", st)):
        for gname, g in (("gnormal", g_n), ("gsymm", sg)):
            tb = engine.EmbeddingTable.wrap(rows, acc=acc)
            res["sparse_adagrad_rows_%s_%s_us" % (name, gname)] = timeit(
                lambda: engine.sparse_adagrad(tb, u_t, nu, g, None, 0.05))
    print(json.dumps(res))
    dist.destroy_process_group()
main()
