"""Training-step tail on the device: fused MSE loss (+ backward seed) and multi-tensor Adam.

Mirrors what the reference's training loop does around the model call (utils.py:191-195: `loss = loss_fn(pred, y);
loss.backward(); optimizer.step(); optimizer.zero_grad(); total_loss += loss.item()`, with `nn.MSELoss()` train.py:222
and `torch.optim.Adam(params, lr, weight_decay)` train.py:207-211) in three launches and without the per-batch host
sync: the loss stays a device scalar (accumulate it on the device, read it once per epoch).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import lib as _l
from . import ops


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: torch.Tensor, target: torch.Tensor, world_scale: float):
        ops._require_cuda(pred, target)
        if pred.dim() != 2 or pred.stride(1) != 1 or pred.shape != target.shape:
            raise RuntimeError("mse_loss: pred must be [rows, cols] with unit column stride and match the target's shape")
        lib = _l.load()
        rows, cols = pred.shape
        target = target.contiguous().float()
        grad = torch.empty((rows, cols), dtype=pred.dtype, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        ws = ops._workspace(lib.aero_mse_workspace_bytes(), pred.device)
        n = max(rows * cols, 1)
        with torch.cuda.device(pred.device):
            rc = lib.aero_mse_loss_grad(ops._ptr(pred), ops._ptr(target), ops._ptr(grad), ops._ptr(loss), rows, cols,
                                        pred.stride(0) if rows > 1 else max(pred.stride(0), cols), cols,
                                        ops.dtype_code(pred), world_scale / n, 2.0 * world_scale / n, ops._ptr(ws),
                                        ws.numel(), ops._stream())
        _l.check(rc, "aero_mse_loss_grad")
        ops.LaunchCounter.add()
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g.to(grad.dtype), None, None      # g is the seed (1 for loss.backward()): one small launch


def mse_loss(pred: torch.Tensor, target: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """nn.MSELoss()(pred, target) (mean reduction) as a device scalar, with dL/dpred produced in the same pass.
    `scale` multiplies loss and gradient (a rank of a partitioned mesh passes n_own / N so the ranks' losses add up to
    the global mean)."""
    return _MSEFn.apply(pred, target, float(scale))


class FusedAdam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics, every tensor updated by ONE kernel launch.

    `master_weights=True` keeps an fp32 master copy of bf16 parameters (the update is applied to the master and the
    parameter is its rounding), which the reference's pure-bf16 mode (train.py:30-33) does not have; default off =
    the reference's behaviour.  Gradients that are `None` are skipped, like torch.  `zero_grad()` follows
    torch.optim (set_to_none=True by default, utils.py:194)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, master_weights: bool = False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam got an empty parameter list")
        ops._require_cuda(*self.params)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay or 0.0)
        self.step_count = 0
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.master = None
        if master_weights and any(p.dtype != torch.float32 for p in self.params):
            self.master = torch.cat([p.detach().reshape(-1).float() for p in self.params])
        self._table_dev = torch.empty(len(self.params) * C.sizeof(_l.AdamSeg), dtype=torch.uint8, device=dev)
        # pinned staging for the table upload; four buffers rotate so a rebuild never rewrites bytes an earlier
        # asynchronous upload may still be reading (no host-side event wait, which a CUDA-graph capture forbids)
        self._table_hosts = [torch.empty(len(self.params) * C.sizeof(_l.AdamSeg), dtype=torch.uint8).pin_memory()
                             for _ in range(4)]
        self._table_turn = 0
        self._key = None
        self._steps = [torch.zeros(len(self.params), dtype=torch.int32, device=dev) for _ in range(2)]
        self.max_elems = max(p.numel() for p in self.params)
        self.param_groups = [{"lr": self.lr, "params": self.params}]     # what lr schedulers touch

    def _table(self):
        """Device segment table; rebuilt (one small pinned-host -> device copy) only when a pointer changed."""
        key = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0,
                     0 if p.grad is None else ops.dtype_code(p.grad)) for p in self.params)
        if key == self._key:
            return
        host = self._table_hosts[self._table_turn % 4]
        self._table_turn += 1
        segs = (_l.AdamSeg * len(self.params)).from_buffer(host.numpy())
        off = 0
        for i, p in enumerate(self.params):
            n = p.numel()
            if not p.is_contiguous() or (p.grad is not None and not p.grad.is_contiguous()):
                raise RuntimeError("FusedAdam: parameters and gradients must be contiguous")
            s = segs[i]
            s.param, s.grad = p.data_ptr(), (p.grad.data_ptr() if p.grad is not None else None)
            s.m, s.v = self.m.data_ptr() + 4 * off, self.v.data_ptr() + 4 * off
            s.master = (self.master.data_ptr() + 4 * off) if (self.master is not None and p.dtype != torch.float32) else None
            s.n, s.p_dtype = n, ops.dtype_code(p)
            s.g_dtype = ops.dtype_code(p.grad) if p.grad is not None else s.p_dtype
            off += n
        self._table_dev.copy_(host, non_blocking=True)
        self._key = key

    @torch.no_grad()
    def step(self) -> None:
        self.lr = float(self.param_groups[0]["lr"])
        self._table()
        self.step_count += 1
        lib = _l.load()
        with torch.cuda.device(self._table_dev.device):
            rc = lib.aero_adam_step(ops._ptr(self._table_dev), len(self.params), self.max_elems,
                                    ops._ptr(self._steps[0]), ops._ptr(self._steps[1]), self.lr, self.betas[0],
                                    self.betas[1], self.eps, self.weight_decay, ops._stream())
        self._steps.reverse()
        _l.check(rc, "aero_adam_step")
        ops.LaunchCounter.add()

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self):
        return {"step": self.step_count, "steps": self._steps[0], "m": self.m, "v": self.v, "master": self.master, "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd) -> None:
        self.step_count = int(sd["step"])
        self._steps[0].copy_(sd["steps"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        if self.master is not None and sd.get("master") is not None:
            self.master.copy_(sd["master"])
