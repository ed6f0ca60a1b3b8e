"""BASELINE configs[0]: the reference's own detector (its unmodified nn/tasks.py parsing an n-scale xLSTM-YOLO YAML)
forwards and back-propagates one 640x640 image on CPU with this repo's drop-ins plugged in at the operator seam (B2),
the module level (B1) and the layer-stack level (B0), against the reference's own cell + PyTorch mLSTM.
Runs only where the reference tree is mounted (this container); tests/golden/run_reference_model.py does the work in a
child process because it has to mock matplotlib and install a stand-in `mlstm_kernels` before importing the reference."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nn/modules/vision_lstm"), reason="reference tree not mounted")
def test_reference_detector_with_dropins_at_three_levels(tmp_path):
    env = dict(os.environ, YOLO_CONFIG_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "run_reference_model.py")], cwd=str(tmp_path),
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["cells"] == 4 and d["out_shape"] == [1, 84, 8400]
    for level in ("B2", "B1", "B0"):
        # fp32 end to end through the whole detector (conv stem, C3k2, two ViL fusion blocks, PAN head, Detect)
        assert d[f"{level}_y"] < 5e-4, d
        assert d[f"{level}_dx"] < 5e-4, d
