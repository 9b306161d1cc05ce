"""Shared MLP + LayerNorm block (interface of reference models/mlp.py:5-51).

Parameter names are the checkpoint contract: `layers.{i}.weight|bias` and `layer_norm.weight|bias`
(SURVEY.md section 8b).  Inside a processor step the Linear/LayerNorm chain is executed by the fused
block kernels (processor.py reads the parameters through `split_first` / `tail`); called on its own
(encoders) its thin first Linear (K <= 16 raw features) is a streaming kernel (ops.ThinLinearFn) and everything
after it runs on the same fused block kernel (processor.DenseTailFn); a
128-wide input (the decoder) runs whole on it, a narrower last Linear zero-padded (processor.DenseMLPFn); shapes
the kernel does not cover stay a dense row-wise chain of library ops on the GPU.  CPU tensors are refused.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_hidden_layers: int = 1,
                 activation_fn: str = "relu", dropout: float = 0.0, use_layer_norm: bool = True):
        super().__init__()
        self.use_layer_norm = use_layer_norm
        self.activation_name = activation_fn
        # getattr raises AttributeError for unknown names, like the reference (mlp.py:37)
        self.activation = getattr(F, activation_fn)
        if num_hidden_layers > 0:
            widths = [input_dim] + [hidden_dim] * (num_hidden_layers + 1) + [output_dim]
        else:
            widths = [input_dim, output_dim]  # a single Linear, no activation (mlp.py:29-32)
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        if use_layer_norm:
            self.layer_norm = nn.LayerNorm(output_dim)
        self.dropout = nn.Dropout(dropout)

    def _fusable(self, x: torch.Tensor) -> bool:
        """Everything after the first Linear can run on the fused block kernel: CUDA rows, 128-wide hidden and
        output layers, at least one hidden Linear, a supported activation, no active dropout."""
        from .. import lib as _l
        from ..ops import D
        if not (x.is_cuda and x.dim() == 2 and x.size(0) > 0 and x.dtype in (torch.bfloat16, torch.float32)):
            return False
        if len(self.layers) < 3 or self.activation_name not in _l.ACT_CODES:
            return False
        if self.training and self.dropout.p > 0:
            return False
        if not all(l.out_features == D for l in self.layers[:-1]) or not all(l.in_features == D for l in self.layers[1:]):
            return False
        # last Linear: D wide, or narrower without LayerNorm (zero-padded to D columns; needs a D-wide input so the
        # first Linear is the kernel's own first GEMM -- the decoder, mgn.py:130)
        out = self.layers[-1].out_features
        return out == D or (out < D and not self.use_layer_norm and self.layers[0].in_features == D)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..ops import _require_cuda
        _require_cuda(x)                                   # like every module of the path: no CPU fallback
        if self._fusable(x):
            from ..ops import D
            from ..processor import dense_mlp, dense_tail
            hidden, w_out, b_out, gamma, beta = self.tail()
            if self.layers[0].in_features == D:
                out = dense_mlp(len(hidden), self.activation_name, self.use_layer_norm, x, self.layers[0].weight,
                                self.layers[0].bias, hidden, w_out, b_out, gamma, beta)
                return out if w_out.size(0) == D else out[:, : w_out.size(0)]
            from .. import ops as _ops
            lin0 = self.layers[0]
            if _ops.thin_linear_ok(x, lin0):           # K <= 16 raw features: streaming kernel, d(W) + d(b) in one pass
                z = _ops.ThinLinearFn.apply(x, lin0.weight, lin0.bias)
            else:
                z = lin0(x)                            # [rows, in] x [in, 128]: a library GEMM, bias included
            return dense_tail(len(hidden), self.activation_name, self.use_layer_norm, z, hidden, w_out, b_out,
                              gamma, beta)
        # shapes / activations the fused kernel does not cover: a dense row-wise chain of library ops on the GPU,
        # announced once and counted (never silent; ops.Fallbacks)
        from ..ops import Fallbacks
        Fallbacks.note("mlp_library_chain",
                       f"MLP {[l.in_features for l in self.layers]} -> {self.layers[-1].out_features} "
                       f"(activation '{self.activation_name}') is outside the fused block kernel (128-wide layers, "
                       f"relu/tanh/sigmoid/elu/leaky_relu)")
        last = len(self.layers) - 1
        for i, lin in enumerate(self.layers):
            x = lin(x)
            if i < last:
                x = self.dropout(self.activation(x))
        return self.layer_norm(x) if self.use_layer_norm else x

    # ---- views used by the fused processor path -------------------------------------------------
    @property
    def num_hidden(self) -> int:
        """Number of Linear(hidden, hidden) layers between the first and the last Linear."""
        return len(self.layers) - 2

    def tail(self):
        """(hidden [(W, b)...], W_out, b_out, gamma, beta) -- everything after the first Linear."""
        hidden = [(l.weight, l.bias) for l in self.layers[1:-1]]
        out = self.layers[-1]
        if self.use_layer_norm:
            gamma, beta = self.layer_norm.weight, self.layer_norm.bias
        else:
            gamma = torch.ones(out.out_features, device=out.weight.device, dtype=out.weight.dtype)
            beta = torch.zeros_like(gamma)
        return hidden, out.weight, out.bias, gamma, beta
