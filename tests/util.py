"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def params_of(g, prefix="param."):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def quat_err(a, b):
    """max over quaternions of min(|a-b|_inf, |a+b|_inf): q and -q are the same rotation."""
    a = a.detach().double().reshape(-1, 4)
    b = b.detach().double().reshape(-1, 4)
    d = torch.minimum((a - b).abs().amax(-1), (a + b).abs().amax(-1))
    return float(d.max()) if d.numel() else 0.0


def max_err(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max()) if a.numel() else 0.0


def rel_err(a, b):
    """max abs error scaled by the largest magnitude of the reference tensor."""
    scale = max(float(b.double().abs().max()), 1e-12) if b.numel() else 1.0
    return max_err(a, b) / scale
