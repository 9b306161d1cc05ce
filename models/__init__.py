"""Drop-in `models` package: the reference's train.py / inference.py / utils.py import `models.<file>`
(train.py:85,106,131,155; utils.py:285-353).  Each module here re-exports the B200-native class of the same name
from aero_gnn_b200.models, so those scripts run unchanged with this repository root on sys.path."""
