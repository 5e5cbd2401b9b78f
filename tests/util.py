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


def sum_of_terms_scale(agreement, agreement_grad):
    """conditioning scale of d loss / d beta of the OUTPUT layer: that gradient is sum_c gz_c with
    gz_c = dL/dagreement_c * agreement_c (1 - agreement_c) (reference qec_module.py:193), a near-cancelling sum under
    the spread loss; a sum computed in FP32 (here: FP32 atomics over ~260 CTAs in nondeterministic order) is accurate
    relative to sum_c |gz_c|, not relative to its own value"""
    a = agreement.detach().double()
    return float((agreement_grad.double() * a * (1.0 - a)).abs().sum())


def grads_close(named_grads, ref_grads, tol, term_scales=None):
    """Every parameter gradient within tol of the reference, relative to
    max(|ref_k|_max, 25% of the largest gradient entry of the whole network).
    The floor matters for scalars that are near-cancelling sums: under the spread loss
    the activation gradients of a sample sum to zero, so d loss / d pool2.beta is pure
    cancellation residue ~100x below the other gradients (in the reference as well): one
    float32 ulp of difference in an output activation moves it by more than 1e-4 of itself."""
    scale = max(float(v.double().abs().max()) for v in ref_grads.values())
    bad = []
    for k, g in named_grads.items():
        ref = ref_grads[k]
        denom = max(float(ref.double().abs().max()), 0.25 * scale, 1e-30)
        if term_scales and k in term_scales:  # a sum of cancelling terms: accurate relative to the sum of |terms|
            denom = max(denom, term_scales[k])
        e = max_err(g, ref) / denom
        print("grad margin %s err=%.2e tol=%.2e" % (k, e, tol))
        if not e < tol:
            bad.append((k, e))
    return bad
