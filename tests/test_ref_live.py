"""LIVE randomised pin of the oracle against the reference's own source files (CPU, build container only).

When /root/reference exists (it does in the build container, not on the GPU box) the reference's unmodified
`models.py` / `train_*.py` are executed on the jax/flax/optax stand-in of tests/golden/refshim for a few dozen random
shapes, id patterns (heavy duplication, i == j, counts of 0 and exactly x_max, ragged m / o, D = 1 ... 33) and
hyper-parameters per trainer, and the oracle must agree with every returned value to 1e-11
(tests/golden/ref_live_check.py).  Skipped where the reference tree is absent; the committed ref_*.npz vectors
(tests/test_ref_golden.py) are the travelling subset of the same check."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "wikipedia")), reason="reference tree not present")


@pytest.mark.timeout(300)
@pytest.mark.parametrize("trainer,seed", [("wikipedia", 1), ("spotify", 2), ("pinterest", 3), ("records", 4)])
def test_oracle_matches_reference_source_on_random_cases(trainer, seed):
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "ref_live_check.py"), trainer, "20", str(seed)],
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK 20"), (r.stdout[-2000:], r.stderr[-4000:])
