"""Shim for the reference module path `models.bsms_mgn` -> aero_gnn_b200.models.bsms_mgn."""
from aero_gnn_b200.models.bsms_mgn import BiStridedMeshGraphNet  # noqa: F401
# the older BFS-bistride design that upstream's stale bytecode of this module still holds
from aero_gnn_b200.models.bsms_gmp import (  # noqa: F401,E402
    MultiScaleGraphPreprocessor, BSMSGMP, BSMS_MeshGraphNet, create_bsms_model_from_config)
