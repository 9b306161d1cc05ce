"""Helpers shared by the model classes: run a list of MeshGraphNetLayer modules as one fused stack."""
from __future__ import annotations

from typing import Iterable

import torch

from .. import ops
from ..processor import run_stack


def run_layers(layers: Iterable, plan: ops.GraphPlan, x: torch.Tensor, e_csr: torch.Tensor):
    """Apply processor layers in order (reference loop mgn.py:127-128); edge latents stay in CSR order."""
    layers = list(layers)
    if not layers:
        return x, e_csr
    cfg = layers[0].stack_config()
    for layer in layers[1:]:
        if layer.stack_config() != cfg:
            raise RuntimeError("all layers of one processor stack must share depth, activation and aggregation")
    steps = [layer.step_weights(x.dtype) for layer in layers]
    return run_stack(cfg, plan, x, e_csr, steps)


def encoder_kwargs(activation_fn: str, dropout: float):
    return dict(activation_fn=activation_fn, dropout=dropout, use_layer_norm=True)
