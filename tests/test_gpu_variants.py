"""Every kernel variant against the oracle: the parity sweep of test_gpu_parity.py re-run in a child
process with the dispatch pinned (MLSTM_FORCE_VARIANT is read once per process): single-pass forward
+ single-pass backward, and two-phase forward + chunk-parallel backward, whatever the shape would pick."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["11", "22", "12", "21"])
def test_parity_sweep_with_pinned_variant(variant):
    env = dict(os.environ, MLSTM_FORCE_VARIANT=variant)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q",
                        "-x", "-k", "test_cuda_matches_oracle or test_initial_and_last_states or sigmoid_input_gate", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
