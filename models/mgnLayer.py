"""Shim for the reference module path `models.mgnLayer` -> aero_gnn_b200.models.mgnLayer."""
from aero_gnn_b200.models.mgnLayer import EdgeBlock, EdgeBlockSum, NodeBlock, MeshGraphNetLayer  # noqa: F401
