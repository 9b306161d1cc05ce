"""north_star: "train.py, inference.py and config.yaml run unchanged".  The reference's own train.py (byte copy under
oracle/_ref, populated by oracle/make_ref.py where the upstream checkout exists) drives the drop-in models through
its own loop: utils.train / utils.evaluate, torch Adam, ReduceLROnPlateau, checkpoint + summary files
(scripts/run_reference_train.py has the stand-ins and the synthetic dataset)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exp,precision", [("airfoil_mgn", "single"), ("airfoil_mgn", "bf16"),
                                           ("airfoil_pooling_mgn", "single"), ("airfoil_fourier_mgn", "bf16")])
def test_reference_train_py_runs_on_the_drop_in_models(exp, precision):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "train.py")):
        pytest.skip("oracle/_ref is not populated (no reference checkout where the tree was built)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_train.py"), exp, "6", "6", precision],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["epochs"] == 6 and len(out["train_losses"]) == 6
    assert all(v == v and v < 1e6 for v in out["train_losses"] + out["val_losses"])          # finite
    assert min(out["train_losses"][1:]) < out["train_losses"][0]                              # it trains (4 shuffled tiny meshes: noisy)
    assert "model_weights.pt" in out["run_dir_files"] and "training_summary.txt" in out["run_dir_files"]
    assert out["state_dict_tensors"] > 100
