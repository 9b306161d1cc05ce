"""Run the reference's OWN train.py (oracle/_ref/train.py, byte copy) on the drop-in models of this repository.

north_star: "the same nn.Module constructors and forward signatures stay in place, so train.py, inference.py and
config.yaml run unchanged".  train.py imports `models.<file>` (train.py:85,106,131,155): here `models` is the
repository's shim package, so every model train.py builds from config.yaml is an aero_gnn_b200 module running the
sm_100a kernels, while train.main(), utils.train / utils.evaluate (utils.py:171-219), the Adam / ReduceLROnPlateau
set-up and the checkpoint writing are the reference's code, unmodified.  What this image lacks is stood in
(oracle/standins.py: torch_scatter, torch_geometric incl. a PyG-style DataLoader, pyvista, matplotlib), and the
dataset -- the reference reads .vtk files that do not exist here -- is replaced by synthetic airfoil meshes with the
same per-sample attributes (dataset.py:52-104).

    python scripts/run_reference_train.py [experiment=airfoil_mgn] [epochs=2] [n_meshes=6] [precision=single]
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import make_ref, standins  # noqa: E402

exp = sys.argv[1] if len(sys.argv) > 1 else "airfoil_mgn"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_meshes = int(sys.argv[3]) if len(sys.argv) > 3 else 6
precision = sys.argv[4] if len(sys.argv) > 4 else "single"
if not make_ref.available() or not os.path.exists(os.path.join(REF, "train.py")):
    sys.exit("oracle/_ref is not populated (run `python oracle/make_ref.py` where /root/reference exists)")

standins.install(REF, pin_models=False)      # `models` stays the repository's shim package
standins.install_full()
sys.path.append(REF)                         # train, utils, dataset, inference: the reference's files
import yaml  # noqa: E402
import utils as ref_utils  # noqa: E402
import train as ref_train  # noqa: E402
import models.mgn  # noqa: E402
assert "aero_gnn_b200" in models.mgn.MeshGraphNet.__module__, "models.* must resolve to the drop-in package"

from aero_gnn_b200.meshes import airfoil_o_mesh  # noqa: E402


def synthetic_datasets(data_dir, dataset_type, params, dtype):
    """Stands in for dataset.create_datasets (reads VTK files): lists of Data with x, edge_attr, edge_index, y, pos."""
    items = []
    for s in range(n_meshes):
        m = airfoil_o_mesh(36, 14, seed=s)
        items.append(standins.Data(x=m.node_attr.to(dtype), edge_attr=m.edge_attr.to(dtype), edge_index=m.edge_index,
                                   y=m.target.to(dtype), pos=m.pos.to(dtype), airfoil=f"synthetic-{s}"))
    n_val = max(1, n_meshes // 6)
    stats = {k: torch.zeros(1) for k in ("node_mean", "node_std", "edge_mean", "edge_std", "target_mean", "target_std")}
    return items[: n_meshes - 2 * n_val], items[n_meshes - 2 * n_val: n_meshes - n_val], items[n_meshes - n_val:], stats


ref_train.create_datasets = synthetic_datasets
cfg = yaml.safe_load(open(os.path.join(REF, "config.yaml")))
params = ref_utils.get_experiment_config(cfg["experiments"][exp], cfg)
params["training"].update(epochs=epochs, batch_size=2, precision=precision, device="cuda:0")
params["experiment_name"] = exp
work = tempfile.mkdtemp(prefix="aero_ref_train_")
os.chdir(work)
ref_train.main(params)
runs = []
for d, _, files in os.walk(work):
    if "training_losses.json" in files:
        runs.append(d)
assert runs, "train.py wrote no training_losses.json"
losses = json.load(open(os.path.join(runs[0], "training_losses.json")))
sd = torch.load(os.path.join(runs[0], "model_weights.pt"), map_location="cpu")
print(json.dumps({"experiment": exp, "model_class": params["model"]["name"], "epochs": losses["total_epochs"],
                  "train_losses": losses["train_losses"], "val_losses": losses["val_losses"],
                  "state_dict_tensors": len(sd), "run_dir_files": sorted(os.listdir(runs[0]))}))
