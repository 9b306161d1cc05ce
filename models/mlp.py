"""Shim for the reference module path `models.mlp` -> aero_gnn_b200.models.mlp."""
from aero_gnn_b200.models.mlp import MLP  # noqa: F401
