"""aero_gnn_b200 -- B200-native (sm_100a) implementation of the aero-gnn message-passing hot path.

Public surface:
  aero_gnn_b200.models.*   nn.Module mirror of the reference's models/{mlp,mgnLayer,mgn,bsms_mgn,
                           poolmgn,fouriermgn}.py (same class names, constructors, forward
                           signatures and state_dict keys)
  aero_gnn_b200.ops        tensor-level wrappers over the C ABI (include/aero_gnn.h)
  aero_gnn_b200.build      nvcc build of libaero_sm100.so
"""
__version__ = "0.1.0"
