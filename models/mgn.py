"""Shim for the reference module path `models.mgn` -> aero_gnn_b200.models.mgn."""
from aero_gnn_b200.models.mgn import MeshGraphNet  # noqa: F401
