"""CPU-side checks: the C-ABI library loads and exports every symbol include/esr.h declares, the
ctypes structs match the header's layout, the host never computes without CUDA, and the torch-CPU
baseline port agrees with the NumPy oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "esr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(esr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from esrecsys_b200 import _lib, build
    build.build()
    h = _lib.lib()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), "libesr.so does not export %s" % n
    # the ctypes binding covers the header exactly
    assert sorted(_lib.declared_symbols()) == names
    assert h.esr_version() == 100
    assert b"workspace" in h.esr_strerror(-2)


def test_struct_layouts_match_header():
    from esrecsys_b200 import _lib
    # EsrTable: u32, i32, i64, 2 ptr, 4 ptr ; EsrPlan: u32, i32, i64, 9 ptr, sort_impl + reserved ; EsrGloveCfg: 4x4, 2x8, 4 floats, 2 ints
    assert C.sizeof(_lib.EsrTable) == 16 + 6 * 8
    assert C.sizeof(_lib.EsrPlan) == 16 + 9 * 8 + 8 and _lib.EsrPlan.n_valid.offset == 16 + 8 * 8
    assert _lib.EsrPlan.sort_impl.offset == 16 + 9 * 8
    assert C.sizeof(_lib.EsrGloveCfg) == 16 + 16 + 16 + 8 + 8 + 24 + 24 + 8      # + loss_log, loss_step, loss_log_len, reserved2, loss_host
    assert _lib.EsrGloveCfg.B.offset == 16 and _lib.EsrGloveCfg.lr.offset == 32


def test_header_enums_match_the_binding():
    """The ESR_SORT_* values of include/esr.h are the ones esrecsys_b200/_lib.py passes in EsrPlan.sort_impl."""
    import re
    from esrecsys_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "esr.h")).read()
    m = re.search(r"enum \{ ESR_SORT_AUTO = (\d+), ESR_SORT_WIDE = (\d+), ESR_SORT_LIBRARY = (\d+) \}", hdr)
    assert m, "ESR_SORT_* enum not found in include/esr.h"
    assert tuple(int(x) for x in m.groups()) == (_lib.ESR_SORT_AUTO, _lib.ESR_SORT_WIDE, _lib.ESR_SORT_LIBRARY)


def test_pure_queries_work_without_gpu():
    from esrecsys_b200 import _lib
    h = _lib.lib()
    assert h.esr_plan_workspace_bytes(0) > 0
    a, b = h.esr_plan_workspace_bytes(2048), h.esr_plan_workspace_bytes(1 << 21)
    assert 0 < a < b
    assert h.esr_glove_workspace_bytes(65536, 128, 0) > 2 * 65536 * 20
    assert h.esr_glove_workspace_bytes(-1, 128, 0) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from esrecsys_b200 import _lib, engine
    with pytest.raises(_lib.EsrError):
        engine.EmbeddingTable(16, 64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "esrecsys_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
def test_torch_port_matches_numpy_oracle(bias_mode):
    import torch
    from esrecsys_b200 import synth
    from oracle import glove as og
    from oracle import glove_torch as ogt
    from oracle import optim as oopt
    V, D, B = 500, 32, 256
    E, b = synth.init_glove_tables(V, D, 0)
    b = (np.random.default_rng(3).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = synth.glove_batches(V, B, 2, 1)
    # dense Adam (the reference's rule)
    En, bn = E.copy(), b.copy()
    st = dict(count=0, muE=np.zeros_like(E), nuE=np.zeros_like(E), mub=np.zeros_like(b), nub=np.zeros_like(b))
    Et, bt = torch.from_numpy(E.copy()), torch.from_numpy(b.copy())
    stt = dict(count=0, muE=torch.zeros_like(Et), nuE=torch.zeros_like(Et), mub=torch.zeros_like(bt), nub=torch.zeros_like(bt))
    for k in range(2):
        ln = og.step_adam(En, bn, st, ids[k, 0], ids[k, 1], counts[k], 1e-3, bias_mode)
        lt = ogt.step_adam_dense(Et, bt, stt, torch.from_numpy(ids[k, 0].astype(np.int64)),
                                 torch.from_numpy(ids[k, 1].astype(np.int64)), torch.from_numpy(counts[k]), 1e-3, bias_mode)
        np.testing.assert_allclose(lt, ln, rtol=1e-5)
    np.testing.assert_allclose(Et.numpy(), En, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(bt.numpy(), bn, rtol=1e-5, atol=2e-6)
    # sparse Adagrad (north star)
    En, bn = E.copy(), b.copy()
    aE, ab = np.full_like(E, 0.1), np.full_like(b, 0.1)
    Et, bt = torch.from_numpy(E.copy()), torch.from_numpy(b.copy())
    aEt, abt = torch.full_like(Et, 0.1), torch.full_like(bt, 0.1)
    for k in range(2):
        og.step_adagrad(En, bn, aE, ab, ids[k, 0], ids[k, 1], counts[k], 0.05, bias_mode)
        ogt.step_adagrad_sparse(Et, bt, aEt, abt, torch.from_numpy(ids[k, 0].astype(np.int64)),
                                torch.from_numpy(ids[k, 1].astype(np.int64)), torch.from_numpy(counts[k]), 0.05, bias_mode)
    np.testing.assert_allclose(Et.numpy(), En, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(aEt.numpy(), aE, rtol=1e-5, atol=1e-7)
    assert oopt.ADAGRAD_EPS == ogt.ADAGRAD_EPS


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--vocab", "20000", "--dim", "64", "--batch", "4096"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_inbatch_cfg_layout_and_workspace_query():
    """EsrInbatchCfg matches the header (u32, i32, 3 x i64, 2 x i32, 3 x f32, i32) and the workspace query is pure."""
    from esrecsys_b200 import _lib
    assert C.sizeof(_lib.EsrInbatchCfg) == 8 + 24 + 8 + 12 + 4
    assert _lib.EsrInbatchCfg.D.offset == 32 and _lib.EsrInbatchCfg.b_norm.offset == 48
    h = _lib.lib()
    cfg = _lib.EsrInbatchCfg()
    cfg.struct_size = C.sizeof(_lib.EsrInbatchCfg)
    cfg.loss_kind, cfg.Bq, cfg.Bk, cfg.D, cfg.margin, cfg.scale, cfg.b_norm = 0, 8192, 8192, 128, 1.0, 1.0, 8192.0
    full = h.esr_inbatch_workspace_bytes(C.byref(cfg))
    assert full > 8192 * 8192 * 2                      # holds the bf16 dL/dS matrix
    cfg.chunk_rows = 1024
    assert 0 < h.esr_inbatch_workspace_bytes(C.byref(cfg)) < full
    cfg.D = 100
    assert h.esr_inbatch_workspace_bytes(C.byref(cfg)) == 0     # unsupported width is refused, not rounded


def test_argument_validation_needs_no_gpu():
    """Every entry point rejects null pointers / bad shapes / unknown ranks with ESR_EINVAL before it touches CUDA
    (include/esr.h error convention), so these calls are safe on a box without a GPU."""
    from esrecsys_b200 import _lib
    h = _lib.lib()
    EINVAL = _lib.ESR_EINVAL
    assert b"inval" in h.esr_strerror(EINVAL).lower() or b"argument" in h.esr_strerror(EINVAL).lower()
    # null table / plan / cfg
    assert h.esr_table_gather_f32(None, None, 4, None, None) == EINVAL
    assert h.esr_table_export_f32(None, None, None) == EINVAL
    assert h.esr_plan_build_i32(None, None, 0, None) == EINVAL
    assert h.esr_glove_prep_f32(None, None, None, None, None, None, 0, None) == EINVAL
    assert h.esr_glove_step_f32(None, None, None, None, None, None, None, None, 0, None) == EINVAL
    # a table whose row width is not a multiple of 4 floats (rows must be 16-byte aligned)
    t = _lib.EsrTable()
    t.struct_size = C.sizeof(_lib.EsrTable)
    t.D, t.V = 6, 10
    assert h.esr_table_gather_f32(C.byref(t), None, 0, None, None) == EINVAL
    # a struct older than the library's (struct_size too small)
    p = _lib.EsrPlan()
    p.struct_size = 8
    assert h.esr_plan_build_i32(C.byref(p), None, 0, None) == EINVAL
    # dense optimizers: step count starts at 1 (optax bias correction), null params
    assert h.esr_dense_adam_f32(None, None, None, None, 16, 1e-3, 0.9, 0.999, 1e-8, 0, None) == EINVAL
    assert h.esr_dense_adam_f32(None, None, None, None, 16, 1e-3, 0.9, 0.999, 1e-8, 1, None) == EINVAL
    assert h.esr_dense_sgdm_f32(None, None, None, 16, 0.1, 0.9, None) == EINVAL
    # ranking / retrieval
    assert h.esr_sort_cols_f32(None, 10, 1, 0, 10, None, None, None, 0, None) == EINVAL
    assert h.esr_sample_uniform_i32(1, 1, 8, 0, None, None) == EINVAL
    assert h.esr_rowwise_dot_f32(None, None, 4, 0, None, None) == EINVAL
    # peer path: more ranks than one NVSwitch domain, rank outside the group, too many floats
    assert h.esr_peer_gather_f32(None, None, 9, None, None, 4, 128, None, None, None) == EINVAL
    assert h.esr_peer_gather_f32(None, None, 2, None, None, 4, 128, None, None, None) == EINVAL
    assert h.esr_peer_allreduce_f32(None, 2, 2, None, None, 0, None, None) == EINVAL
    assert h.esr_peer_allreduce_f32(None, 2, 0, None, None, 7, None, None) == EINVAL
    assert h.esr_route_plan_i32(None, None, 4, 2, None, None, None, None, None, 0, None) == EINVAL
    # in-batch scorer: unsupported width -> no workspace, and the call itself refuses
    cfg = _lib.EsrInbatchCfg()
    cfg.struct_size = C.sizeof(_lib.EsrInbatchCfg)
    cfg.Bq = cfg.Bk = 256
    cfg.D = 96
    cfg.scale = 1.0
    cfg.b_norm = 256.0
    assert h.esr_inbatch_workspace_bytes(C.byref(cfg)) == 0
    assert h.esr_inbatch_fwd_bwd_bf16(None, None, C.byref(cfg), None, None, None, None, 0, None) in (EINVAL, _lib.ESR_ENOTSUP)
    # host decoders: null output buffers
    n = C.c_int64(0)
    used = C.c_size_t(0)
    assert h.esr_decode_cooccur_b64(b"", 0, None, None, None, 0, C.byref(n), C.byref(used)) in (0, EINVAL)


def test_sass_of_the_hot_kernels():
    """Static guard on the built library (cuobjdump, no GPU): the default row pass is spill-free and stages rows with
    cp.async (LDGSTS) + packed f32x2 math; the in-batch kernels really are tcgen05 / TMEM / TMA code."""
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_evidence.py")], capture_output=True, text=True,
                         check=True).stdout
    rows = {}
    for line in out.splitlines():
        if line.startswith("#") or "|" not in line:
            continue
        name, regs, smem, local, n_ins, mn = [x.strip() for x in line.split("|")]
        rows[name] = dict(regs=int(regs), local=int(local), mn=mn)
    default = rows["k_glove_rows_grp_async<8, 4, 2, true>"]          # D = 128: 8 lanes x 4 float4 per row
    assert default["local"] == 0 and default["regs"] <= 128                 # 2 CTAs of 256 threads per SM
    assert "LDGSTS" in default["mn"] and "FFMA2" in default["mn"]
    scores = [v for k, v in rows.items() if k.startswith("k_inbatch_scores<")]
    bwd = [v for k, v in rows.items() if k.startswith("k_inbatch_bwd<")]
    assert scores and bwd
    for v in scores + bwd:
        assert v["local"] == 0
        assert "UTCHMMA" in v["mn"] and "LDTM" in v["mn"] and "UTMALDG" in v["mn"]
    # TMA-store epilogue of the passes that write the bf16 dL/dS block (the softmax statistics pass, mode 1, writes none)
    assert all("UTMASTG" in v["mn"] for k, v in rows.items() if k.startswith("k_inbatch_scores<") and not k.endswith(", 1>"))
    assert any("UBLKCP" in v["mn"] for k, v in rows.items() if k.startswith("k_glove_rows_tma<"))


def test_mlp_tower_parameters_are_views_of_flat_buffers():
    """MLPTower (configs[3]): parameters, gradients and Adam moments are 16-byte aligned views of four flat buffers, so
    one esr_dense_adam_f32 launch and one gradient all-reduce per tower cover every tensor (host layout only: no GPU)."""
    import torch
    from esrecsys_b200.inbatch import MLPTower
    gen = torch.Generator().manual_seed(1)
    t = MLPTower(10, 6, 5, gen, "cpu")                      # odd sizes: W1 60, b1 6 (-> 8), W2 30 (-> 32), b2 5 (-> 8)
    assert t.flat["p"].numel() == 60 + 8 + 32 + 8
    for name, buf in t.flat.items():
        group = getattr(t, {"p": "p", "g": "g", "mu": "mu", "nu": "nu"}[name])
        for k, v in group.items():
            off = (v.data_ptr() - buf.data_ptr()) // 4
            assert 0 <= off and off + v.numel() <= buf.numel() and off % 4 == 0, (name, k, off)
            assert v.is_contiguous()
    t.g["W2"].fill_(3.0)
    t.g["b1"].fill_(-1.0)
    assert float(t.flat["g"].sum()) == 3.0 * 30 - 6.0        # writes through the views land in the flat buffer
    x = torch.randn(7, 10, generator=gen)
    y = t.forward(x)
    want = torch.relu(x @ t.p["W1"] + t.p["b1"]) @ t.p["W2"] + t.p["b2"]
    assert torch.allclose(y, want, atol=1e-6)
    dy = torch.randn(7, 5, generator=gen)
    dx = t.backward(dy)
    h = torch.relu(x @ t.p["W1"] + t.p["b1"])
    assert torch.allclose(t.g["b2"], dy.sum(0), atol=1e-5) and torch.allclose(t.g["W2"], h.t() @ dy, atol=1e-5)
    dh = (dy @ t.p["W2"].t()) * (h > 0)
    assert torch.allclose(t.g["b1"], dh.sum(0), atol=1e-5) and torch.allclose(dx, dh @ t.p["W1"].t(), atol=1e-5)
