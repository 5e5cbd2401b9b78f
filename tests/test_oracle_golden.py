"""CPU: the oracle restatement vs the golden vectors produced by the unmodified
reference modules (tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import qec_oracle as orc
from tests.util import load_golden, params_of, quat_err, max_err, rel_err, grads_close

TOL = 1e-4  # BASELINE.json north_star tolerance (fp32, quaternions up to sign)


def test_quat_ops_golden():
    g = load_golden("quat_ops")
    assert max_err(orc.ham(g["q"], g["r"]), g["qmul"]) < 1e-6
    assert max_err(orc.qrotv(g["qn"], g["v"]), g["qrotv"]) < 1e-6


def test_symeig_golden():
    g = load_golden("symeig")
    w, V = orc.symeig_rows(g["cov"])
    assert max_err(w, g["w"]) < 1e-5
    # rows are eigenvectors: A = V^T diag(w) V  (pytorch-cusolver/test.py:28-31)
    rec = V.transpose(1, 2) @ torch.diag_embed(w) @ V
    assert max_err(rec, g["cov"]) < 1e-5
    top = orc.fix(V[:, 3])
    assert quat_err(top, g["top"]) < 1e-5
    wb, Vb = orc.symeig_rows(g["bug"].unsqueeze(0))
    assert max_err(wb, g["bug_w"]) < 1e-5


def test_symeig_top_gradient_golden():
    """pytorch-autograd-solver/test.py:101-141: gradient of mean(sign-fixed top
    eigenvector of sum_8 n(x) n(x)^T) w.r.t. x, batched vs per-matrix loop."""
    g = load_golden("symeig")
    x = g["x"].clone().requires_grad_(True)
    q = x / x.norm(dim=-1, keepdim=True)
    cov = torch.einsum("nkc,nkd->ncd", q, q)
    _, V = orc.symeig_rows(cov)
    orc.fix(V[:, 3]).mean().backward()
    assert rel_err(x.grad, g["grad_x"]) < 1e-4
    # and the closed-form backward restatement agrees with eigh's autograd
    w, Vr = orc.symeig_rows(cov.detach())
    gV = torch.zeros_like(Vr)
    gV[:, 3] = torch.sign(Vr[:, 3, :1]) / Vr[:, 3].numel()
    gM = orc.symeig_rows_backward(None, gV, w, Vr)
    gMs = 0.5 * (gM + gM.transpose(1, 2))
    cov2 = cov.detach().clone().requires_grad_(True)
    _, V2 = torch.linalg.eigh(cov2)
    orc.fix(V2[:, :, 3]).mean().backward()
    gA = 0.5 * (cov2.grad + cov2.grad.transpose(1, 2))
    assert rel_err(gMs, gA) < 1e-4


MODULE_CASES = ["module_small", "module_pool1like", "module_T1", "module_T2", "module_O1"]


@pytest.mark.parametrize("name", MODULE_CASES + ["module_cfg1_B2"])
def test_module_forward_golden(name):
    g = load_golden(name)
    T = int(g["dims"][5])
    pose, agr = orc.qec_module_forward(params_of(g), g["points"], g["lrfs"], g["act"], T)
    assert pose.shape == g["pose"].shape
    assert quat_err(pose, g["pose"]) < TOL
    assert max_err(agr, g["agreement"]) < TOL


@pytest.mark.parametrize("name", MODULE_CASES)
def test_module_backward_golden(name):
    g = load_golden(name)
    T = int(g["dims"][5])
    params = {k: v.clone().requires_grad_(True) for k, v in params_of(g).items()}
    points = g["points"].clone().requires_grad_(True)
    lrfs = g["lrfs"].clone().requires_grad_(True)
    act = g["act"].clone().requires_grad_(True)
    pose, agr = orc.qec_module_forward(params, points, lrfs, act, T)
    ((pose * g["cot.pose"]).sum() + (agr * g["cot.agreement"]).sum()).backward()
    assert rel_err(points.grad, g["grad.points"]) < TOL
    assert rel_err(lrfs.grad, g["grad.lrfs"]) < TOL
    assert rel_err(act.grad, g["grad.act"]) < TOL
    for k, v in params.items():
        assert rel_err(v.grad, g["grad." + k]) < TOL, k


def test_module_float64_close_to_golden():
    """the float64 oracle is the tie-breaker for fp32 rounding: it must sit
    within the same tolerance of the fp32 reference."""
    g = load_golden("module_small")
    T = int(g["dims"][5])
    p64 = {k: v.double() for k, v in params_of(g).items()}
    pose, agr = orc.qec_module_forward(p64, g["points"].double(), g["lrfs"].double(), g["act"].double(), T)
    assert quat_err(pose, g["pose"]) < TOL
    assert max_err(agr, g["agreement"]) < TOL


def test_net_golden():
    g = load_golden("net_small")
    T = int(g["dims"][3])
    params = {k: v.clone().requires_grad_(True) for k, v in params_of(g).items()}
    pose, a = orc.qec_net_forward(params, g["points"], g["lrfs"], g["act"], T)
    assert quat_err(pose, g["pose"]) < TOL
    assert max_err(a, g["agreement"]) < TOL
    loss = orc.spread_loss(a, g["labels"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    got = {k: v.grad for k, v in params.items()}
    assert grads_close(got, {k: g["grad." + k] for k in got}, TOL) == []


def test_sia_net_golden():
    g = load_golden("sia_net_small")
    T = int(g["dims"][3])
    params = {k: v.clone().requires_grad_(True) for k, v in params_of(g).items()}
    pose, a = orc.qec_sia_net_forward(params, g["points"], g["lrfs"], g["act"], T)
    assert pose.shape == g["pose"].shape
    assert quat_err(pose, g["pose"]) < TOL
    assert max_err(a, g["agreement"]) < TOL
    loss = orc.sia_spread_loss(a, g["labels"]) + orc.pose_diff_loss(pose)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    got = {k: v.grad for k, v in params.items()}
    assert grads_close(got, {k: g["grad." + k] for k in got}, TOL) == []


def test_equivariance():
    """rotate points by qrotv(q,.) and LRFs by q (x) .: poses become q (x) pose up
    to sign, activations unchanged."""
    params = orc.init_net_params(8, 5, seed=3, dtype=torch.float64)
    points, lrfs, act, _ = orc.synthetic_batch(2, seed=5, dtype=torch.float64)
    q = torch.tensor([0.3, -0.5, 0.7, 0.4], dtype=torch.float64)
    q = q / q.norm()
    qe = q.expand(2, 256, 9, 4)
    points_r = orc.qrotv(qe, points)
    lrfs_r = orc.ham(qe, lrfs)
    lrfs_r = lrfs_r * torch.where(lrfs_r[..., :1] < 0, -1.0, 1.0)
    p0, a0 = orc.qec_net_forward(params, points, lrfs, act, 3)
    p1, a1 = orc.qec_net_forward(params, points_r, lrfs_r, act, 3)
    assert max_err(a0, a1) < 1e-9
    assert quat_err(orc.ham(q.expand_as(p0), p0), p1) < 1e-9


def test_gather_golden():
    """the loader gather (SURVEY 8a row d) against outputs of the reference's own dataset class
    (tests/golden/make_golden_gather.py): bit-exact"""
    g = load_golden("gather")
    for n in range(3):
        p, l, a = orc.neighbour_gather(g["s%d.point_set" % n], g["s%d.lrf_set" % n], g["s%d.idx" % n],
                                       g["s%d.wrong_ids" % n])
        assert torch.equal(p, g["s%d.points" % n])
        assert torch.equal(l, g["s%d.lrfs" % n])
        assert torch.equal(a, g["s%d.act" % n])


def test_knn_oracle_properties():
    """oracle knn_indices (the checker of qec_knn): first neighbour of a centre that is a point is a point
    at distance 0, distances are non-decreasing, ties go to the smaller index, -1 padding when N < K"""
    g = torch.Generator().manual_seed(3)
    pts = torch.round(torch.randn(2, 50, 3, generator=g) * 2) / 2
    ci = torch.stack([torch.randperm(50, generator=g)[:7] for _ in range(2)])
    idx = orc.knn_indices(pts, 9, centre_idx=ci)
    assert idx.shape == (2, 7, 9)
    for b in range(2):
        for p in range(7):
            c = pts[b, ci[b, p]]
            d = ((pts[b, idx[b, p]] - c) ** 2).sum(-1)
            assert float(d[0]) == 0.0
            assert bool((d[1:] >= d[:-1]).all())
            ties = (d[1:] == d[:-1])
            assert bool((idx[b, p, 1:][ties] > idx[b, p, :-1][ties]).all())
    few = orc.knn_indices(pts[:, :4], 9, centre_idx=ci % 4)
    assert bool((few[..., 4:] == -1).all()) and bool((few[..., :4] >= 0).all())
