"""Shim for the reference module path `models.bsms_mgn` -> aero_gnn_b200.models.bsms_mgn."""
from aero_gnn_b200.models.bsms_mgn import BiStridedMeshGraphNet  # noqa: F401
