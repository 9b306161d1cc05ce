"""Drop-in mirror of the reference `models` package for the message-passing path."""
from .mlp import MLP
from .mgnLayer import EdgeBlock, EdgeBlockSum, NodeBlock, MeshGraphNetLayer
from .mgn import MeshGraphNet
from .bsms_mgn import BiStridedMeshGraphNet
from .poolmgn import poolMGN
from .fouriermgn import FourierMeshGraphNet

__all__ = [
    "MLP", "EdgeBlock", "EdgeBlockSum", "NodeBlock", "MeshGraphNetLayer", "MeshGraphNet",
    "BiStridedMeshGraphNet", "poolMGN", "FourierMeshGraphNet",
]
