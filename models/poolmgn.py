"""Shim for the reference module path `models.poolmgn` -> aero_gnn_b200.models.poolmgn."""
from aero_gnn_b200.models.poolmgn import poolMGN  # noqa: F401
