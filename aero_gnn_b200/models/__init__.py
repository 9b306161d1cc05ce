"""Drop-in mirror of the reference `models` package for the message-passing path."""
from .mlp import MLP
from .mgnLayer import EdgeBlock, EdgeBlockSum, NodeBlock, MeshGraphNetLayer
from .mgn import MeshGraphNet
from .bsms_mgn import BiStridedMeshGraphNet
from .poolmgn import poolMGN
from .fouriermgn import FourierMeshGraphNet
from .bistride_ops import BistridePooling, Unpool, WeightedEdgeConv, GMP
from .bsms_gmp import MultiScaleGraphPreprocessor, BSMSGMP, BSMS_MeshGraphNet, create_bsms_model_from_config

__all__ = [
    "MLP", "EdgeBlock", "EdgeBlockSum", "NodeBlock", "MeshGraphNetLayer", "MeshGraphNet",
    "BiStridedMeshGraphNet", "poolMGN", "FourierMeshGraphNet",
    "BistridePooling", "Unpool", "WeightedEdgeConv", "GMP",
    "MultiScaleGraphPreprocessor", "BSMSGMP", "BSMS_MeshGraphNet", "create_bsms_model_from_config",
]
