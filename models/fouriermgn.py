"""Shim for the reference module path `models.fouriermgn` -> aero_gnn_b200.models.fouriermgn."""
from aero_gnn_b200.models.fouriermgn import FourierMeshGraphNet  # noqa: F401
