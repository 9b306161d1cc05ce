import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|_inf, tiny): relative error of a tensor against its reference scale."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def rrmse(pred: torch.Tensor, ref: torch.Tensor) -> float:
    """The reference's own relative error of node predictions (inference.py:113-126): per-feature RMSE divided by
    the per-feature mean |reference|, averaged over features.  Used for the bf16 tolerance (<= 1e-2)."""
    pred, ref = pred.detach().double().cpu(), ref.detach().double().cpu()
    rmse = ((pred - ref) ** 2).mean(dim=0).sqrt()
    scale = ref.abs().mean(dim=0)
    return float(torch.where(scale > 1e-8, rmse / scale, torch.zeros_like(rmse)).mean())


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
