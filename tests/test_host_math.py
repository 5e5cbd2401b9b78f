"""CPU: the device math header (csrc/qec_math.cuh) compiled for the host with g++ and checked
against float64 numpy / the oracle.  Test scaffolding only -- lets the arithmetic of the Jacobi
eigensolver, the acos polynomial, the Cholesky adjoint and the qrotv adjoint be verified on the
GPU-less build container; the product never runs this host build."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import qec_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "math_host.cpp")
LIB = os.path.join(HERE, "host", "libqecmath_host.so")
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
IU = np.triu_indices(4)


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["g++", "-O2", "-mfma", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return ctypes.CDLL(LIB)


def weighted_outer(n, J, spread, rng):
    base = rng.standard_normal((n, 1, 4))
    q = base + spread * rng.standard_normal((n, J, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    q *= np.where(q[..., :1] < 0, -1, 1)
    w = rng.random((n, J))
    w /= w.sum(-1, keepdims=True)
    return np.einsum("nj,njc,njd->ncd", w, q, q).astype(np.float32)


@pytest.mark.parametrize("spread,J", [(0.1, 9), (0.5, 9), (3.0, 9), (100.0, 9), (0.3, 256), (100.0, 512)])
def test_jacobi_top_eigenvector(lib, spread, J):
    rng = np.random.default_rng(int(spread * 10) + J)
    M = weighted_outer(4000, J, spread, rng)
    M10 = np.ascontiguousarray(M[:, IU[0], IU[1]])
    n = len(M10)
    v = np.zeros((n, 4), np.float32)
    lam = np.zeros(n, np.float32)
    sw = np.zeros(n, np.int32)
    lib.h_top_eigvec(n, P(M10), P(v), P(lam), P(sw))
    w64, V64 = np.linalg.eigh(M.astype(np.float64))
    top = V64[:, :, 3]
    gap = w64[:, 3] - w64[:, 2]
    err = np.minimum(np.abs(v - top).max(-1), np.abs(v + top).max(-1))
    assert sw.max() <= 6                       # quadratic convergence: a handful of sweeps
    assert np.abs(lam - w64[:, 3]).max() < 1e-6
    assert np.abs(np.linalg.norm(v, axis=-1) - 1).max() < 1e-6
    # eigenvector error = rounding floor + eps/gap (same scaling as LAPACK ssyevd)
    assert (err <= 1e-6 + 5e-7 / np.maximum(gap, 1e-12)).all()


@pytest.mark.parametrize("fn", ["h_top_eigvec_psd", "h_top_eigvec_psd2"])
@pytest.mark.parametrize("spread,J", [(0.1, 9), (0.5, 9), (3.0, 9), (100.0, 9), (0.3, 256), (100.0, 512)])
def test_squaring_top_eigenvector(lib, fn, spread, J):
    """the repeated-squaring solvers of the routing kernels (psd2 = the low-latency variant of the
    packed cluster kernels) against float64 LAPACK, incl. the zero matrix and tiny / huge scales"""
    rng = np.random.default_rng(int(spread * 10) + J + 1)
    M = weighted_outer(4000, J, spread, rng)
    M[0] = 0
    M[1] *= 1e-30
    M[2] *= 1e30
    M10 = np.ascontiguousarray(M[:, IU[0], IU[1]])
    n = len(M10)
    v = np.zeros((n, 4), np.float32)
    steps = np.zeros(n, np.int32)
    getattr(lib, fn)(n, P(M10), P(v), P(steps))
    w64, V64 = np.linalg.eigh(M.astype(np.float64))
    top = V64[:, :, 3]
    gap = (w64[:, 3] - w64[:, 2]) / np.maximum(w64[:, 3], 1e-300)
    err = np.minimum(np.abs(v - top).max(-1), np.abs(v + top).max(-1))
    assert np.array_equal(v[0], np.array([0, 0, 0, 1], np.float32))
    assert np.abs(np.linalg.norm(v, axis=-1) - 1).max() < 1e-6
    ok = err <= 1e-6 + 5e-7 / np.maximum(gap, 1e-12)
    ok[0] = True
    assert ok.all(), (err[~ok][:5], gap[~ok][:5])
    assert steps.max() <= (13 if fn.endswith("2") else 26)


def test_jacobi_full_decomposition_and_edge_cases(lib):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((500, 4, 4))
    A = (A + A.transpose(0, 2, 1)).astype(np.float32)
    A[0] = 0                                       # zero matrix: no rotation, V = I
    A[1] = np.eye(4, dtype=np.float32)             # degenerate spectrum
    A[2] = np.diag([4, 3, 2, 1]).astype(np.float32)
    bug = np.array([[2.7378, -2.7378, -0.4963, -0.2952], [-2.7378, 2.7378, 0.4963, 0.2952],
                    [-0.4963, 0.4963, 1.8648, 1.1091], [-0.2952, 0.2952, 1.1091, 0.6596]], np.float32)
    A[3] = bug                                     # reference pytorch-autograd-solver/test.py:186-190
    M10 = np.ascontiguousarray(A[:, IU[0], IU[1]])
    w = np.zeros((500, 4), np.float32)
    V = np.zeros((500, 4, 4), np.float32)
    lib.h_jacobi_full(500, P(M10), P(w), P(V))
    Asym = np.triu(A) + np.triu(A, 1).transpose(0, 2, 1)
    rec = np.einsum("nij,nj,nkj->nik", V, w, V)
    assert np.abs(rec - Asym).max() < 2e-5
    assert np.abs(np.einsum("nij,nik->njk", V, V) - np.eye(4)).max() < 1e-5
    assert np.abs(np.sort(w, -1) - np.linalg.eigvalsh(Asym.astype(np.float64))).max() < 2e-5
    assert np.array_equal(V[0], np.eye(4, dtype=np.float32))


def test_distance_polynomial(lib):
    x = np.concatenate([np.linspace(-1, 1, 200001), [0.9999, 0.99991, -0.9999, 0.0, 1e-8]]).astype(np.float32)
    out = np.zeros_like(x)
    grad = np.zeros_like(x)
    lib.h_one_minus_dist(len(x), P(x), P(out), P(grad))
    xt = torch.from_numpy(x).clone().requires_grad_(True)  # float32 like the reference
    d = 2 * torch.acos(torch.clamp(xt.abs(), max=0.9999)) / np.pi
    d.sum().backward()
    assert np.abs(out - (1 - d.detach().numpy())).max() < 3e-7
    g = xt.grad.numpy()
    # next to the clamp 1 - x^2 cancels: one float32 ulp of x is 3e-5 of the derivative
    assert (np.abs(grad - g) <= 2e-6 + 1e-4 * np.abs(g)).all()
    assert grad[np.abs(x) > np.float32(0.9999)].max() == 0.0 and grad[x == 0].max() == 0.0
    # the Estrin form the cluster kernels evaluate (dependency depth 3 instead of 8): same accuracy
    est = np.zeros_like(x)
    lib.h_one_minus_dist_estrin(len(x), P(x), P(est))
    d64 = 1 - 2 * np.arccos(np.minimum(np.abs(x.astype(np.float64)), np.float64(np.float32(0.9999)))) / np.pi
    assert np.abs(est - d64).max() < 2e-7 and np.abs(out - d64).max() < 2e-7
    assert np.abs(est - out).max() < 1.5e-7


def test_top_eigenvector_adjoint(lib):
    rng = np.random.default_rng(5)
    for spread in (0.1, 3.0):
        M = weighted_outer(3000, 9, spread, rng)
        M10 = np.ascontiguousarray(M[:, IU[0], IU[1]])
        n = len(M)
        w64, V64 = np.linalg.eigh(M.astype(np.float64))
        top = V64[:, :, 3] * np.where(V64[:, :1, 3] < 0, -1, 1)
        g = rng.standard_normal((n, 4)).astype(np.float32)
        u = np.zeros((n, 4), np.float32)
        p = top.astype(np.float32)
        lam = w64[:, 3].astype(np.float32)
        lib.h_adjoint(n, P(M10), P(p), P(lam), P(g), P(u))
        uref = sum(((V64[:, :, i] * g).sum(-1) / (w64[:, 3] - w64[:, i]))[:, None] * V64[:, :, i] for i in range(3))
        rel = np.abs(u - uref).max(-1) / np.abs(uref).max(-1)
        gap = w64[:, 3] - w64[:, 2]
        assert (rel <= 2e-6 + 1e-6 / gap).all()
        # and it is what autograd of eigh gives: dL/dM = u p^T (symmetrised)
        Mt = torch.from_numpy(M[:8].astype(np.float64)).requires_grad_(True)
        _, V = torch.linalg.eigh(Mt)
        (orc.fix(V[:, :, 3]) * torch.from_numpy(g[:8]).double()).sum().backward()
        gM = np.einsum("ni,nj->nij", u[:8].astype(np.float64), top[:8])
        sym = lambda a: 0.5 * (a + a.transpose(0, 2, 1))  # noqa: E731
        assert np.abs(sym(gM) - sym(Mt.grad.numpy())).max() < 1e-4 * np.abs(Mt.grad.numpy()).max()


def test_qrotv_and_adjoint(lib):
    rng = np.random.default_rng(7)
    n = 2000
    q = rng.standard_normal((n, 4)).astype(np.float32)
    v = rng.standard_normal((n, 3)).astype(np.float32)
    g = rng.standard_normal((n, 3)).astype(np.float32)
    x = np.zeros((n, 3), np.float32)
    gq = np.zeros((n, 4), np.float32)
    gv = np.zeros((n, 3), np.float32)
    lib.h_qrotv(n, P(q), P(v), P(g), P(x), P(gq), P(gv))
    qt = torch.from_numpy(q).double().requires_grad_(True)
    vt = torch.from_numpy(v).double().requires_grad_(True)
    xr = orc.qrotv(qt, vt)
    (xr * torch.from_numpy(g).double()).sum().backward()
    assert np.abs(x - xr.detach().numpy()).max() < 2e-5
    assert np.abs(gq - qt.grad.numpy()).max() < 2e-5
    assert np.abs(gv - vt.grad.numpy()).max() < 2e-5


@pytest.mark.parametrize("weight_decay,scale", [(0.0, 1.0), (0.01, 0.125)])
def test_adam_element_update_matches_torch_adam(lib, weight_decay, scale):
    """csrc/adam_math.cuh (the element update and the double-precision bias corrections of qec_adam_flat) against
    torch.optim.Adam -- the optimizer of the reference's train_cls.py:44-51 -- over 25 steps with gradients
    spanning five decades"""
    n, steps = 1003, 25
    g = torch.Generator().manual_seed(4)
    p0 = torch.randn(n, generator=g)
    grads = torch.stack([torch.randn(n, generator=g) * (10.0 ** (t % 5 - 2)) for t in range(steps)])
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([q], lr=1e-3, weight_decay=weight_decay)
    for t in range(steps):
        q.grad = grads[t].clone()
        opt.step()
    p = p0.numpy().copy()
    m = np.zeros(n, dtype=np.float32)
    v = np.zeros(n, dtype=np.float32)
    gin = np.ascontiguousarray((grads / scale).numpy())  # the kernel reads grad * grad_scale
    lib.h_adam_steps.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 + [
        ctypes.c_float, ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    lib.h_adam_steps(n, steps, P(p), P(gin), P(m), P(v), 1e-3, 0.9, 0.999, 1e-8, weight_decay, scale)
    st = opt.state[q]
    assert np.abs(p - q.detach().numpy()).max() <= 2e-6 * max(1.0, float(q.detach().abs().max()))
    assert np.abs(m - st["exp_avg"].numpy()).max() <= 1e-6 * max(1.0, float(st["exp_avg"].abs().max()))
    assert np.abs(v - st["exp_avg_sq"].numpy()).max() <= 1e-6 * max(1.0, float(st["exp_avg_sq"].abs().max()))
