"""Shim for the reference module path `models.bistride_ops` (bytecode-only upstream) -> aero_gnn_b200.models.bistride_ops."""
from aero_gnn_b200.models.bistride_ops import BistridePooling, Unpool, WeightedEdgeConv, GMP  # noqa: F401
