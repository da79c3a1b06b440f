"""Host-side arithmetic of tools/table_sweep.py (BASELINE.json configs[4]): the closed-form row function used for the
full-size lookup property and the SURVEY.md 8(d) K7 byte accounting.  CPU only."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("table_sweep", os.path.join(ROOT, "tools", "table_sweep.py"))
ts = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ts)


def test_row_function_is_exact_and_id_dependent():
    g = torch.tensor([0, 1, 7, 99_999_999, 2**31 - 1], dtype=torch.int64)
    f = ts.row_function(g, 128, "cpu")
    assert f.shape == (5, 128) and f.dtype == torch.float32
    ref = ((g.numpy()[:, None] * 131 + np.arange(128)[None, :] * 7919) & 0xFFFFF) / float(1 << 20) - 0.5
    assert np.array_equal(f.numpy().astype(np.float64), ref)          # every value is exactly representable in f32
    assert not torch.equal(f[0], f[1]) and not torch.equal(f[2], f[3])
    b = ts.bias_function(g, "cpu")
    assert np.array_equal(b.numpy().astype(np.float64), ((g.numpy() * 40503) & 0xFFFF) / float(1 << 20))


def test_lookup_accounting():
    uniq = torch.tensor([0, 1, 2, 3, 4, 5, 8, 9, 16], dtype=torch.int32)
    assert ts.remote_unique(uniq, 0, 4) == 9 - 4                        # 0, 4, 8, 16 are rank 0's rows
    assert ts.remote_unique(uniq, 1, 4) == 9 - 3
    assert ts.remote_unique(uniq, 0, 1) == 0
    assert ts.lookup_bytes(10, 128) == 10 * (4 + 512)


def test_bench_reference_arm_line_matches_the_contract():
    """`bench.py --impl reference` (CPU only): one JSON line with impl = reference, the metric / unit of our arm, the SAME
    config keys our N = 1 line carries, the requested steps, and an e2e object with zero copy bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--vocab", "20000", "--batch", "4096", "--dim", "32"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pairs/sec" and line["unit"] == "pairs/s"
    assert line["steps"] == 3 and line["warmup"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, root)
    import bench
    argv, sys.argv = sys.argv, ["bench.py", "--vocab", "20000", "--batch", "4096", "--dim", "32"]
    try:
        ours = bench.config_single(bench.parse())
    finally:
        sys.argv = argv
    assert set(line["config"]) == set(ours) and line["config"]["workload"] == ours["workload"]
