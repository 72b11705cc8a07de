"""Shared helpers for the parity tests (test infrastructure, may import oracle/)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Return dict[str, torch.Tensor] from tests/golden/<name>."""
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files if z[k].dtype.kind in "fiub"}  # skips string tags


def section(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def rel_fro(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def assert_close_fro(a, b, rtol, atol=0.0, what=""):
    """||a-b||_F <= rtol*||b||_F + atol*sqrt(numel).  The atol term covers gradients that are
    analytically zero (e.g. every item-side bias: sum_j dS_ij = 0) and therefore pure rounding noise."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    err = float((a - b).norm())
    bound = rtol * float(b.norm()) + atol * (b.numel() ** 0.5)
    assert err <= bound, f"{what}: |a-b|={err:.3e} > {bound:.3e} (|b|={float(b.norm()):.3e})"
