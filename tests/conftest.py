import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


VARIANTS = ("A", "B")  # published variants of k2.get_rnnt_prune_ranges; B is the default


def load_golden(name: str, tag: str = "f32", variant: str = "B"):
    """Golden vectors of a case: <name>.<variant>.<tag>.npz for the pruned cases (one file per variant of the
    prune-range selection), <name>.<tag>.npz for the vanilla one."""
    path = os.path.join(GOLDEN, f"{name}.{variant}.{tag}.npz")
    if not os.path.exists(path):
        path = os.path.join(GOLDEN, f"{name}.{tag}.npz")
    return dict(np.load(path))


def check_summary(t: torch.Tensor, gold: dict, key: str, rtol: float, what: str = ""):
    """Compare a tensor with the strided sub-sample / sum / norm summary written by
    oracle/make_golden.py.  Error is measured relative to the largest golden magnitude."""
    stride = int(gold[f"{key}.stride"])
    sub = torch.from_numpy(gold[f"{key}.sub"]).to(torch.float64)
    got = t.detach().reshape(-1).to(torch.float64).cpu()[::stride]
    assert got.numel() == sub.numel(), (what, key, got.numel(), sub.numel())
    scale = max(sub.abs().max().item(), 1e-30)
    err = (got - sub).abs().max().item() / scale
    assert err <= rtol, f"{what} {key}: max rel err {err:.3e} > {rtol:.1e}"
    if f"{key}.norm" in gold:
        n = t.detach().to(torch.float64).norm().item()
        gn = float(gold[f"{key}.norm"])
        assert abs(n - gn) <= rtol * max(gn, 1e-30) * 10, f"{what} {key}: norm {n} vs {gn}"
    return err


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().to(torch.float64).cpu()
    b = b.detach().to(torch.float64).cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
