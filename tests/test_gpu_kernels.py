"""GPU parity tests, kernel by kernel: every C-ABI entry point against the CPU oracle
on the same seeded inputs (run on the B200 box: pytest -m gpu).

Tolerance: BASELINE.json north_star -- 1e-4 in fp32, quaternions compared up to sign,
gradients under the same tolerance (relative to the largest reference magnitude)."""
import pytest
import torch

from oracle import qec_oracle as orc
from tests.util import load_golden, quat_err, max_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def cuda():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch.device("cuda:0")


def clustered(shape, g, spread=0.3):
    base = torch.randn(shape[0], *([1] * (len(shape) - 1)), 4, generator=g, dtype=torch.float64)
    q = base + spread * torch.randn(*shape, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    return (q * torch.where(q[..., :1] < 0, -1.0, 1.0)).float()


# ---------------------------------------------------------------- symeig (L0 boundary)
def test_symeig_forward_golden():
    import qecnetworks_b200 as qb

    dev = cuda()
    g = load_golden("symeig")
    w, V = qb.symeig(g["cov"].to(dev))
    w, V = w.cpu(), V.cpu()
    assert max_err(w, g["w"]) < 1e-5
    rec = V.transpose(1, 2) @ torch.diag_embed(w) @ V  # rows are eigenvectors
    assert max_err(rec, g["cov"]) < 1e-5
    assert quat_err(orc.fix(V[:, 3]), g["top"]) < 1e-5
    assert max_err(V @ V.transpose(1, 2), torch.eye(4).expand(80, 4, 4)) < 1e-5
    wb, Vb = qb.symeig(g["bug"].unsqueeze(0).to(dev))
    assert max_err(wb.cpu(), g["bug_w"]) < 1e-5


def test_symeig_top_gradient_golden():
    """the reference's only live backward test (pytorch-autograd-solver/test.py:101-141)"""
    import qecnetworks_b200 as qb

    dev = cuda()
    g = load_golden("symeig")
    x = g["x"].to(dev).requires_grad_(True)
    q = x / x.norm(dim=-1, keepdim=True)
    cov = torch.einsum("nkc,nkd->ncd", q, q)
    _, V = qb.symeig(cov)
    top = V[:, 3]
    (top * torch.sign(top[:, :1])).mean().backward()
    assert rel_err(x.grad.cpu(), g["grad_x"]) < TOL


def test_symeig_backward_vs_oracle():
    import qecnetworks_b200 as qb

    dev = cuda()
    gen = torch.Generator().manual_seed(5)
    A = torch.randn(257, 4, 4, generator=gen)
    A = A @ A.transpose(1, 2)
    w, V = orc.symeig_rows(A.double())
    gw = torch.randn(257, 4, generator=gen)
    gV = torch.randn(257, 4, 4, generator=gen)
    ref = orc.symeig_rows_backward(gw.double(), gV.double(), w, V)
    Ad = A.to(dev).requires_grad_(True)
    wd, Vd = qb.symeig(Ad, fold=False)
    # eigenvector signs are arbitrary: feed a sign-consistent cotangent
    sgn = torch.sign((Vd.detach().cpu().double() * V).sum(-1, keepdim=True))
    torch.autograd.backward([wd, Vd], [gw.to(dev), (gV * sgn.float()).to(dev)])
    assert rel_err(Ad.grad.cpu(), ref) < 1e-3  # F = 1/(w_j - w_i) amplifies fp32 rounding on random spectra
    # folded variant: triu(g) + triu(g^T, 1)
    Af = A.to(dev).requires_grad_(True)
    wf, Vf = qb.symeig(Af, fold=True)
    torch.autograd.backward([wf, Vf], [gw.to(dev), (gV * sgn.float()).to(dev)])
    folded = torch.triu(ref) + torch.triu(ref.transpose(1, 2), 1)
    assert rel_err(Af.grad.cpu(), folded) < 1e-3


# ---------------------------------------------------------------- LRF mean + rotation
@pytest.mark.parametrize("B,P,K,I", [(2, 3, 5, 4), (3, 16, 9, 1), (2, 1, 256, 8), (1, 2, 40, 33)])
def test_lrf_mean_rotate(B, P, K, I):
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(B * 1000 + K)
    lrfs = clustered((B, P, K, I), gen)
    act = (torch.rand(B, P, K, I, generator=gen) < 0.8).float()
    act[0, 0, :, 0] = 0  # a channel with no active neighbour -> zero mean
    lrfs = lrfs * (torch.rand(B, P, K, I, 1, generator=gen) < 0.9)
    points = torch.randn(B, P, K, 3, generator=gen)
    G = torch.randn(B, P, K, I, 3, generator=gen)

    l64 = lrfs.double().requires_grad_(True)
    p64 = points.double().requires_grad_(True)
    mean_ref = orc.lrf_mean(l64, act.double())
    x_ref = orc.rotate_points(mean_ref, p64)
    (x_ref * G.double()).sum().backward()

    ld = lrfs.to(dev).requires_grad_(True)
    pd = points.to(dev).requires_grad_(True)
    mean, xrot = QF.lrf_mean_rotate(ld, act.to(dev), pd)
    assert quat_err(mean.cpu(), mean_ref) < TOL
    assert float(mean[0, 0, 0].abs().max()) == 0.0
    assert max_err(xrot.cpu(), x_ref) < TOL
    (xrot * G.to(dev)).sum().backward()
    assert rel_err(ld.grad.cpu(), l64.grad) < TOL
    assert rel_err(pd.grad.cpu(), p64.grad) < TOL


# ---------------------------------------------------------------- votes
@pytest.mark.parametrize("B,P,K,I,O", [(2, 3, 5, 4, 6), (2, 8, 9, 1, 128), (1, 1, 40, 16, 40), (1, 2, 3, 70, 65)])
def test_votes(B, P, K, I, O):
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(K * 100 + O)
    lrfs = clustered((B, P, K, I), gen)
    lrfs[0, 0, 0] = 0
    t_raw = torch.randn(B * P * K, 4 * I * O, generator=gen)
    t_raw[0, :O] = 0.0  # w == 0 -> sign(0) = 0 zeroes the transform
    G = torch.randn(B, P, O, K * I, 4, generator=gen)

    l64 = lrfs.double().requires_grad_(True)
    t64 = t_raw.double().requires_grad_(True)
    v_ref = orc.votes_from(l64, t64.view(B, P, K, 4, I, O))
    (v_ref * G.double()).sum().backward()

    ld = lrfs.to(dev).requires_grad_(True)
    td = t_raw.to(dev).requires_grad_(True)
    v = QF.votes(ld, td)
    assert max_err(v.cpu(), v_ref) < 1e-5
    (v * G.to(dev)).sum().backward()
    assert rel_err(td.grad.cpu(), t64.grad) < TOL
    assert rel_err(ld.grad.cpu(), l64.grad) < TOL


# ---------------------------------------------------------------- routing
def _routing_case(B, P, O, J, T, frac, seed, need_act=True):
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(seed)
    votes = clustered((B, P, O, J), gen, spread=0.35)
    if frac:
        act = torch.rand(B, P, J, generator=gen) * 0.6 + 0.3
    else:
        act = torch.ones(B, P, J)
    act = act * (torch.rand(B, P, J, generator=gen) < 0.85)
    act[:, :, 0] = 1.0
    votes[0, 0, :, min(1, J - 1)] = 0  # a zero vote (inactive neighbour's zero LRF)
    if B * P > 1:
        votes[-1, -1, 0] = 0  # a capsule with no votes at all -> pose 0
        act[-1, -1, :] = act[-1, -1, :]
    alpha = torch.tensor([1.3])
    beta = torch.tensor([0.7])
    Gp = torch.randn(B, P, O, 4, generator=gen)
    Ga = torch.randn(B, P, O, generator=gen)

    def run_oracle(dt):
        v = votes.clone().to(dt).requires_grad_(True)  # clone: .to() of the same dtype aliases
        a = act.clone().to(dt).requires_grad_(need_act)
        al = alpha.clone().to(dt).requires_grad_(True)
        be = beta.clone().to(dt).requires_grad_(True)
        pose_o, agr_o = orc.routing(v, a, al, be, T)
        ((pose_o * Gp.to(dt)).sum() + (agr_o * Ga.to(dt)).sum()).backward()
        return pose_o.detach(), agr_o.detach(), v.grad, (a.grad if need_act else None), al.grad, be.grad

    ref = run_oracle(torch.float64)
    r32 = run_oracle(torch.float32)  # the reference's own arithmetic: its distance to float64 is the noise floor

    vd = votes.to(dev).requires_grad_(True)
    ad = act.to(dev).requires_grad_(need_act)
    ald = alpha.to(dev).requires_grad_(True)
    bed = beta.to(dev).requires_grad_(True)
    pose, agr = QF.routing(vd, ad, ald, bed, T)
    assert quat_err(pose.cpu(), ref[0]) < TOL
    assert max_err(agr.cpu(), ref[1]) < TOL
    if B * P > 1:
        assert float(pose[-1, -1, 0].abs().max()) == 0.0
    ((pose * Gp.to(dev)).sum() + (agr * Ga.to(dev)).sum()).backward()

    def close(got, k):
        # 1e-4, or 3x the float32 reference's own deviation from float64 where that is
        # larger: D'(x) = -(2/pi)/sqrt(1-x^2) next to the 0.9999 clamp turns one float32
        # ulp of <pose,vote> into a 3e-4 relative change of that vote's gradient.
        tol = max(TOL, 3.0 * rel_err(r32[k], ref[k]))
        e = rel_err(got.cpu(), ref[k])
        print("routing margin k=%d err=%.2e tol=%.2e" % (k, e, tol))
        assert e < tol, (k, e, tol)

    close(vd.grad, 2)
    if need_act:
        close(ad.grad, 3)
    close(ald.grad, 4)
    close(bed.grad, 5)


@pytest.mark.parametrize(
    "B,P,O,J,T,frac",
    [
        (2, 3, 6, 20, 3, True),      # thread per capsule
        (2, 8, 16, 9, 3, False),     # pool1-like
        (2, 2, 5, 9, 1, True),       # T = 1
        (2, 2, 5, 9, 2, True),       # T = 2
        (1, 2, 7, 64, 3, True),      # largest thread-per-capsule J
        (2, 2, 3, 300, 3, True),     # warp per capsule
        (1, 2, 3, 2048, 3, True),    # largest warp-per-capsule J
        (1, 2, 5, 4100, 3, True),    # CTA per capsule, ragged tail
        (1, 1, 3, 100, 5, True),     # T > 3 instantiation
    ],
)
def test_routing(B, P, O, J, T, frac):
    _routing_case(B, P, O, J, T, frac, seed=J * 10 + T)


def test_routing_no_activation_grad():
    _routing_case(2, 4, 8, 9, 3, False, seed=77, need_act=False)
    _routing_case(1, 1, 4, 3000, 3, True, seed=78, need_act=False)


def test_routing_pool2_shape():
    """pool2 geometry of BASELINE config 2 at B=1: 40 capsules x 32768 votes"""
    _routing_case(1, 1, 40, 32768, 3, True, seed=79)


# ---------------------------------------------------------------- gather (row d)
def test_gather_bit_exact():
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(9)
    B, N, P, K = 3, 500, 256, 9
    ps = torch.randn(B, N, 3, generator=gen)
    ls = torch.randn(B, N, 4, generator=gen)
    idx = torch.randint(0, N, (B, P, K), generator=gen)
    idx[:, 200:, :] = -1
    idx[:, :, 5:][torch.rand(B, P, 4, generator=gen) < 0.3] = -1
    idx[0, 3, 2] = 0          # id 0: act 0 but a live LRF
    idx[1, 7, 1] = N - 1      # the zeroed last row
    wrong = torch.stack([idx[b, :50].reshape(-1)[torch.randperm(450, generator=gen)[:6]] for b in range(B)])
    points, lrfs, act, err = QF.neighbour_gather(ps.to(dev), ls.to(dev), idx.to(dev), wrong.to(dev))
    assert int(err.item()) == 0
    for b in range(B):
        w = wrong[b][wrong[b] > 0]
        rp, rl, ra = orc.neighbour_gather(ps[b], ls[b], idx[b], wrong[b])
        assert torch.equal(points[b].cpu(), rp)
        assert torch.equal(lrfs[b].cpu(), rl)
        assert torch.equal(act[b].cpu(), ra)
    bad = idx.clone()
    bad[0, 0, 0] = N + 3
    _, _, _, err = QF.neighbour_gather(ps.to(dev), ls.to(dev), bad.to(dev))
    assert int(err.item()) == 1


@pytest.mark.parametrize("B,N,P,K", [(3, 1024, 256, 9), (2, 2048, 64, 64), (2, 37, 5, 9), (1, 5, 3, 9), (2, 300, 16, 1)])
def test_knn_indices_bit_exact(B, N, P, K):
    """qec_knn against the float32 CPU statement: indices bit-exact, incl. exact distance ties (points on
    a grid, duplicated points), fewer points than K, and centres given as coordinates"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(N * 7 + K)
    pts = torch.randn(B, N, 3, generator=gen)
    pts[0] = torch.round(pts[0] * 4) / 4          # a coarse grid: many exact ties
    if N > 8:
        pts[-1, N // 2:N // 2 + 4] = pts[-1, 0]    # duplicates of one point
    ci = torch.stack([torch.randperm(N, generator=gen)[:P] if N >= P else torch.randint(0, N, (P,), generator=gen)
                      for _ in range(B)])
    ref = orc.knn_indices(pts, K, centre_idx=ci)
    idx, err = QF.knn_indices(pts.to(dev), K, centre_idx=ci.to(dev))
    assert int(err.item()) == 0
    assert torch.equal(idx.cpu(), ref)
    cen = torch.randn(B, P, 3, generator=gen)
    ref2 = orc.knn_indices(pts, K, centres=cen)
    idx2, _ = QF.knn_indices(pts.to(dev), K, centres=cen.to(dev))
    assert torch.equal(idx2.cpu(), ref2)
    bad = ci.clone()
    bad[0, 0] = N + 1
    _, err = QF.knn_indices(pts.to(dev), K, centre_idx=bad.to(dev))
    assert int(err.item()) == 1
    if N >= K:   # the indices feed the loader gather: neighbourhoods of the centres, centre first
        ps, ls = pts.to(dev), torch.randn(B, N, 4, generator=gen).to(dev)
        gp, gl, ga, gerr = QF.neighbour_gather(ps, ls, idx)
        assert int(gerr.item()) == 0 and gp.shape == (B, P, K, 3)


# ---------------------------------------------------------------- fused cluster-resident routing
def _oqi_perm(I, O):
    o = torch.arange(O).view(O, 1, 1)
    q = torch.arange(4).view(1, 4, 1)
    i = torch.arange(I).view(1, 1, I)
    return (q * (I * O) + i * O + o).reshape(-1)


@pytest.mark.parametrize(
    "B,P,K,I,O,T",
    [
        (2, 1, 64, 32, 5, 3),     # J = 2048, one CTA per capsule
        (1, 2, 100, 37, 3, 3),    # J = 3700, ragged channels
        (1, 1, 256, 24, 4, 3),    # J = 6144, cluster of 2
        (1, 1, 250, 50, 3, 2),    # J = 12500, cluster of 4, ragged split, T = 2
        (1, 1, 256, 128, 6, 3),   # J = 32768 (pool2 of BASELINE config 2): 4 CTAs x 512 threads, every CTA full
        (1, 1, 256, 128, 2, 1),   # T = 1
        (1, 2, 250, 100, 3, 3),   # J = 25000: the 512-thread build with a ragged split
        (2, 1, 129, 128, 2, 2),   # J = 16512: just above two full 8 192-vote CTAs, last CTA nearly empty
    ],
)
@pytest.mark.parametrize("wide", [0, 1])
def test_routing_fused(B, P, K, I, O, T, wide, fused_build):
    from qecnetworks_b200 import functional as QF

    if wide and K * I <= 16384:
        pytest.skip("the 512-thread build only serves capsules of more than 16 384 votes")
    fused_build(wide)

    dev = cuda()
    assert QF.routing_fused_supported(K, I, O, T)
    gen = torch.Generator().manual_seed(K * 7 + I)
    J = K * I
    lrfs = clustered((B, P, K, I), gen, spread=0.3)
    act = (torch.rand(B, P, J, generator=gen) * 0.6 + 0.3) * (torch.rand(B, P, J, generator=gen) < 0.85)
    act[:, :, 0] = 1.0
    lrfs = lrfs * (act.view(B, P, K, I, 1) != 0)  # inactive neighbours carry a zero LRF
    base = torch.randn(1, 4, 1, O, generator=gen)
    t_raw = (base + 0.25 * torch.randn(B * P * K, 4, I, O, generator=gen)).reshape(B * P * K, 4 * I * O)
    perm = _oqi_perm(I, O)
    alpha, beta = torch.tensor([1.2]), torch.tensor([0.8])
    Gp = torch.randn(B, P, O, 4, generator=gen)
    Ga = torch.randn(B, P, O, generator=gen)

    def run_oracle(dt):
        l = lrfs.clone().to(dt).requires_grad_(True)
        t = t_raw.clone().to(dt).requires_grad_(True)
        a = act.clone().to(dt).requires_grad_(True)
        al = alpha.clone().to(dt).requires_grad_(True)
        be = beta.clone().to(dt).requires_grad_(True)
        v = orc.votes_from(l, t.view(B, P, K, 4, I, O))
        pose_o, agr_o = orc.routing(v, a, al, be, T)
        ((pose_o * Gp.to(dt)).sum() + (agr_o * Ga.to(dt)).sum()).backward()
        return pose_o.detach(), agr_o.detach(), t.grad, l.grad, a.grad, al.grad, be.grad

    ref = run_oracle(torch.float64)
    r32 = run_oracle(torch.float32)

    ld = lrfs.to(dev).requires_grad_(True)
    td = t_raw[:, perm].contiguous().to(dev).requires_grad_(True)
    ad = act.to(dev).requires_grad_(True)
    ald = alpha.to(dev).requires_grad_(True)
    bed = beta.to(dev).requires_grad_(True)
    pose, agr = QF.routing_fused(td, ld, ad, ald, bed, T)
    assert quat_err(pose.cpu(), ref[0]) < TOL
    assert max_err(agr.cpu(), ref[1]) < TOL
    ((pose * Gp.to(dev)).sum() + (agr * Ga.to(dev)).sum()).backward()
    g_traw = torch.empty_like(t_raw)
    g_traw[:, perm] = td.grad.cpu()

    def close(got, k):
        tol = max(TOL, 3.0 * rel_err(r32[k], ref[k]))
        e = rel_err(got, ref[k])
        print("fused margin k=%d err=%.2e tol=%.2e" % (k, e, tol))
        assert e < tol, (k, e, tol)

    close(g_traw, 2)
    close(ld.grad.cpu(), 3)
    close(ad.grad.cpu(), 4)
    close(ald.grad.cpu(), 5)
    close(bed.grad.cpu(), 6)

    # and the fused path agrees with the general path (materialised votes) on the same input
    l2 = lrfs.to(dev).requires_grad_(True)
    t2 = t_raw.to(dev).requires_grad_(True)
    pose2, agr2 = QF.routing(QF.votes(l2, t2), act.to(dev), alpha.to(dev), beta.to(dev), T)
    assert quat_err(pose.detach().cpu(), pose2.detach().cpu()) < 1e-5
    assert max_err(agr.detach().cpu(), agr2.detach().cpu()) < 1e-5


@pytest.mark.parametrize(
    "B,P,K,I,O,T",
    [
        (2, 1, 256, 128, 3, 3),   # J = 32768, every CTA full (the guard-free instantiation of the 512-thread build)
        (1, 1, 256, 128, 3, 2),   # ... T = 2
        (1, 1, 250, 90, 2, 3),    # J = 22500: 512-thread build, ragged
        (1, 1, 128, 128, 2, 3),   # cluster of 4, full
        (1, 1, 64, 128, 2, 3),    # cluster of 2, full
        (1, 2, 32, 128, 2, 3),    # one CTA per capsule, full
        (1, 2, 100, 30, 3, 3),    # ragged
        (2, 1, 37, 6, 5, 1),      # T = 1, tiny
        (1, 1, 1, 2, 1, 3),       # a single pair, one of its votes dead
    ],
)
@pytest.mark.parametrize("wide", [0, 1])
def test_routing_fused_packed_vs_scalar_and_split(B, P, K, I, O, T, wide, fused_build):
    """the packed-pair kernels (routing_fused2.cu) against the scalar ones (routing_fused.cu) through the
    C ABI on the same inputs, incl. a capsule with a single live vote (|<pose, vote>| = 1: the distance
    gradient is infinite there and must be selected away, not multiplied by zero); and the pre-split
    gradient of qec_routing_fused_bwd_split: hi on the TF32 grid, hi + lo == g bit-exactly"""
    from qecnetworks_b200 import functional as QF, _lib

    if wide and K * I <= 16384:
        pytest.skip("the 512-thread build only serves capsules of more than 16 384 votes")
    fused_build(wide)
    dev = cuda()
    assert QF.routing_fused_packed_supported(K, I, O, T)
    gen = torch.Generator().manual_seed(K * 3 + I)
    J = K * I
    lrfs = clustered((B, P, K, I), gen, spread=0.3)
    act = (torch.rand(B, P, J, generator=gen) * 0.6 + 0.3) * (torch.rand(B, P, J, generator=gen) < 0.85)
    act[:, :, 0] = 1.0
    if J == 2:
        act[:, :, 1] = 0.0
    lrfs = (lrfs * (act.view(B, P, K, I, 1) != 0)).to(dev)
    act = act.to(dev)
    base = torch.randn(1, O, 4, 1, generator=gen)
    t = (base + 0.25 * torch.randn(B * P * K, O, 4, I, generator=gen)).reshape(B * P * K, 4 * I * O).to(dev)
    alpha, beta = torch.tensor([1.2], device=dev), torch.tensor([0.8], device=dev)
    Gp = torch.randn(B, P, O, 4, generator=gen).to(dev)
    Ga = torch.randn(B, P, O, generator=gen).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    Nc = B * P * O

    def fwd(name):
        pose = torch.empty(Nc, 4, device=dev); agr = torch.empty(Nc, device=dev)
        pa = torch.empty(T, Nc, 4, device=dev); stt = torch.empty(Nc, 2, device=dev)
        _lib.call(name, t.data_ptr(), lrfs.data_ptr(), act.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                  pose.data_ptr(), agr.data_ptr(), pa.data_ptr(), stt.data_ptr(), B * P, K, I, O, T, st)
        return pose, agr, pa, stt

    def bwd(name, f, split=False, chain=True):
        g_t = [torch.full(tuple(t.shape), 7.0, device=dev)]
        if split:
            g_t.append(torch.full(tuple(t.shape), 7.0, device=dev, dtype=torch.bfloat16))
        g_l = torch.zeros_like(lrfs); g_a = torch.zeros_like(act); g_ab = torch.zeros(2, device=dev)
        outs = [x.data_ptr() for x in g_t]
        _lib.call(name, t.data_ptr(), lrfs.data_ptr(), act.data_ptr(), alpha.data_ptr(), beta.data_ptr(),
                  f[2].data_ptr(), f[3].data_ptr(), Gp.data_ptr(), Ga.data_ptr(), *outs, g_l.data_ptr(),
                  g_a.data_ptr() if chain else None, g_ab.data_ptr(), B * P, K, I, O, T, st)
        torch.cuda.synchronize()
        return g_t, g_l, g_a, g_ab

    fp, fs = fwd("qec_routing_fused_fwd"), fwd("qec_routing_fused_fwd_scalar")
    assert quat_err(fp[0].cpu(), fs[0].cpu()) < 2e-6 and max_err(fp[1].cpu(), fs[1].cpu()) < 2e-6
    assert quat_err(fp[2].cpu(), fs[2].cpu()) < 2e-6 and max_err(fp[3].cpu(), fs[3].cpu()) < 1e-5
    bp, bs = bwd("qec_routing_fused_bwd", fs), bwd("qec_routing_fused_bwd_scalar", fs)
    bp, bs = ([bp[0][0]] + list(bp[1:])), ([bs[0][0]] + list(bs[1:]))
    for k, (x, y) in enumerate(zip(bp, bs)):
        assert torch.isfinite(x).all(), k
        # dL/dlrfs keeps the radial component the normalisation projects out of dL/dt: it inherits the
        # 1/gap conditioning of the adjoint solve, so the two summation orders differ by ~1e-4 there
        assert rel_err(x.cpu(), y.cpu()) < (5e-4 if k == 1 else 2e-5), (k, rel_err(x.cpu(), y.cpu()))
    # activation without gradient (pool1-like callers): only the top level is processed
    np_, ns_ = bwd("qec_routing_fused_bwd", fs, chain=False), bwd("qec_routing_fused_bwd_scalar", fs, chain=False)
    assert rel_err(np_[0][0].cpu(), ns_[0][0].cpu()) < 2e-5 and rel_err(np_[1].cpu(), ns_[1].cpu()) < 5e-4
    assert float(np_[2].abs().max()) == 0.0
    sp = bwd("qec_routing_fused_bwd_split", fs, split=True)
    hi, lo, g = sp[0][0], sp[0][1].float(), bp[0]
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0          # hi is on the TF32 grid
    assert bool(((hi - g).abs() <= g.abs() * 2.0 ** -11).all())           # nearest TF32 value of the same gradient
    assert bool(((hi + lo - g).abs() <= g.abs() * 2.0 ** -19 + 1e-37).all())   # BF16 remainder: ~2^-20 |g| left


@pytest.mark.parametrize("tcgen05,fwd_tc", [(True, False), (False, False), (True, True)])
def test_transform_routing_node_matches_plain_path(tcgen05, fwd_tc, monkeypatch):
    """_TransformRoutingFused -- backward of quater_gen2 on the package's tcgen05 kernel (qec_gen2_bwd) or on cuBLAS
    over the pre-split gradient, forward on cuBLAS or on qec_gen2_fwd -- against the plain path (cuBLAS 3xTF32
    forward, FP32 SGEMM backward) on a pool2-shaped layer: same outputs, same gradients.  With the cuBLAS forward on
    both sides the transforms are bit-identical and only the backward differs (2e-5); with the own forward the
    transforms differ in the last bits (5e-7) and the 1/gap conditioning of the routing amplifies that."""
    from qecnetworks_b200 import functional as QF
    from qecnetworks_b200 import qec_module as QM

    monkeypatch.setattr(QM, "GEN2_BACKWARD_TCGEN05", tcgen05)
    monkeypatch.setattr(QM, "GEN2_FORWARD_TCGEN05", fwd_tc)
    dev = cuda()
    B, P, K, I, O, T = 2, 1, 256, 128, 5, 3
    gen = torch.Generator().manual_seed(5)
    lrfs = clustered((B, P, K, I), gen, spread=0.3).to(dev)
    act = (torch.rand(B, P, K * I, generator=gen) * 0.6 + 0.3).to(dev)
    h = torch.randn(B * P * K, O, generator=gen).to(dev)
    W = (0.3 * torch.randn(4 * I * O, O, generator=gen)).to(dev)
    b = (0.3 * torch.randn(4 * I * O, generator=gen)).to(dev)
    alpha, beta = torch.tensor([1.1], device=dev), torch.tensor([0.9], device=dev)
    Gp = torch.randn(B, P, O, 4, generator=gen).to(dev)
    Ga = torch.randn(B, P, O, generator=gen).to(dev)

    perm = _oqi_perm(I, O).to(dev)

    def run(node):
        hh, WW, bb = (x.clone().requires_grad_(True) for x in (h, W, b))
        ll, aa = lrfs.clone().requires_grad_(True), act.clone().requires_grad_(True)
        if node:
            pose, agr = QM._TransformRoutingFused.apply(hh, WW, bb, perm, ll, aa, alpha, beta, T)
        else:
            xa, wa = QM._augment(hh, WW, bb, perm)
            pose, agr = QF.routing_fused(QM._SplitTf32Matmul.apply(xa, wa), ll, aa, alpha, beta, T)
        ((pose * Gp).sum() + (agr * Ga).sum()).backward()
        return pose.detach(), agr.detach(), hh.grad, WW.grad, bb.grad, ll.grad, aa.grad

    a, c = run(True), run(False)
    otol = 1e-5 if fwd_tc else 1e-6
    assert quat_err(a[0].cpu(), c[0].cpu()) < otol and max_err(a[1].cpu(), c[1].cpu()) < otol
    for k in range(2, 7):
        tol = (5e-4 if k == 5 else 2e-5) * (5.0 if fwd_tc else 1.0)
        e = rel_err(a[k].cpu(), c[k].cpu())
        print("node margin k=%d bwd_tc=%s fwd_tc=%s err=%.2e tol=%.2e" % (k, tcgen05, fwd_tc, e, tol))
        assert e < tol, (k, e)


def test_spread_loss_kernel():
    """qec_spread_loss against the reference expression (models/qec_net.py:53-59), value and gradient"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(11)
    for B, C in [(32, 40), (5, 10), (1, 3), (64, 70)]:
        x = torch.rand(B, C, generator=gen) * 0.3 + 0.5
        y = torch.randint(0, C, (B,), generator=gen)
        xr = x.clone().double().requires_grad_(True)
        onehot = torch.zeros(B, C, dtype=torch.double).scatter_(1, y.view(-1, 1), 1.0)
        at = (xr * onehot).sum(1)
        ref = ((torch.relu(0.1 - (at.view(-1, 1) - xr)) ** 2) * (1 - onehot)).sum(1).mean()
        ref.backward()
        xd = x.to(dev).requires_grad_(True)
        loss = QF.spread_loss(xd, y.to(dev), 0.1)
        (loss * 3.0).backward()
        assert abs(float(loss) - float(ref)) < 1e-6
        assert max_err(xd.grad.cpu(), 3.0 * xr.grad) < 1e-6


# ---------------------------------------------------------------- register-resident pool1 kernel
@pytest.mark.parametrize("B,P,K,O,T", [(2, 8, 9, 16, 3), (1, 5, 9, 130, 3), (2, 3, 4, 7, 1), (1, 4, 13, 40, 2), (1, 2, 16, 33, 5)])
def test_routing_small(B, P, K, O, T):
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    assert QF.routing_small_supported(K, 1, O, T)
    gen = torch.Generator().manual_seed(K * 13 + O)
    lrfs = clustered((B, P, K, 1), gen, spread=0.3)
    act = (torch.rand(B, P, K, generator=gen) < 0.85).float()
    act[:, :, 0] = 1.0
    lrfs = lrfs * act.view(B, P, K, 1, 1)
    if B * P > 1:
        act[-1, -1] = 0
        lrfs[-1, -1] = 0  # a patch with nothing in it: pose 0, agreement 0
    xrot = torch.randn(B, P, K, 1, 3, generator=gen)
    W1, b1 = torch.randn(O, 3, generator=gen) * 0.5, torch.randn(O, generator=gen) * 0.5
    W2, b2 = torch.randn(4 * O, O, generator=gen) / O ** 0.5, torch.randn(4 * O, generator=gen) * 0.3
    b2 = b2 + torch.tensor([2.0, 0.3, -0.4, 0.5]).repeat_interleave(O)  # a dominant mean transform: clustered votes
    alpha, beta = torch.tensor([1.1]), torch.tensor([0.9])
    Gp = torch.randn(B, P, O, 4, generator=gen)
    Ga = torch.randn(B, P, O, generator=gen)

    def run_oracle(dt):
        ps = [t.clone().to(dt).requires_grad_(True) for t in (W1, b1, W2, b2, alpha, beta)]
        t_raw = orc.transform_quats(xrot.to(dt), ps[0], ps[1], ps[2], ps[3])
        v = orc.votes_from(lrfs.to(dt), t_raw)
        pose_o, agr_o = orc.routing(v, act.to(dt), ps[4], ps[5], T)
        ((pose_o * Gp.to(dt)).sum() + (agr_o * Ga.to(dt)).sum()).backward()
        return [pose_o.detach(), agr_o.detach()] + [p.grad for p in ps]

    ref = run_oracle(torch.float64)
    r32 = run_oracle(torch.float32)
    ps = [t.to(dev).requires_grad_(True) for t in (W1, b1, W2, b2, alpha, beta)]
    Wc = ps[2] @ ps[0]
    bc = torch.addmv(ps[3], ps[2], ps[1])
    pose, agr = QF.routing_small(xrot.to(dev), lrfs.to(dev), act.to(dev), Wc, bc, ps[4], ps[5], T)
    assert quat_err(pose.cpu(), ref[0]) < TOL
    assert max_err(agr.cpu(), ref[1]) < TOL
    if B * P > 1:
        assert float(pose[-1, -1].abs().max()) == 0.0 and float(agr[-1, -1].abs().max()) == 0.0
    ((pose * Gp.to(dev)).sum() + (agr * Ga.to(dev)).sum()).backward()
    for k, p in enumerate(ps):
        tol = max(TOL, 3.0 * rel_err(r32[k + 2], ref[k + 2]))
        e = rel_err(p.grad.cpu(), ref[k + 2])
        print("small margin k=%d err=%.2e tol=%.2e" % (k, e, tol))
        assert e < tol, (k, e, tol)


def test_routing_negative_activations():
    """outside the reference's domain (activations are sigmoid outputs / masks) but inside the API's:
    negative weights make M indefinite; the LARGEST eigenvalue must still be the one that is followed"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(123)
    B, P, O, J, T = 2, 3, 5, 12, 2
    votes = clustered((B, P, O, J), gen, spread=0.3)
    act = torch.rand(B, P, J, generator=gen) * 0.6 + 0.3
    act[:, :, 3::4] *= -0.5
    alpha, beta = torch.tensor([1.0]), torch.tensor([1.0])
    pose_ref, agr_ref = orc.routing(votes.double(), act.double(), alpha.double(), beta.double(), T)
    pose, agr = QF.routing(votes.to(dev), act.to(dev), alpha.to(dev), beta.to(dev), T)
    assert quat_err(pose.cpu(), pose_ref) < TOL
    assert max_err(agr.cpu(), agr_ref) < TOL


def test_gather_golden_gpu():
    """qec_gather against the outputs of the reference's own dataset class: bit-exact"""
    from qecnetworks_b200 import functional as QF
    from tests.util import load_golden as lg

    dev = cuda()
    g = lg("gather")
    ps = torch.stack([g["s%d.point_set" % n] for n in range(3)]).to(dev)
    ls = torch.stack([g["s%d.lrf_set" % n] for n in range(3)]).to(dev)
    idx = torch.stack([g["s%d.idx" % n] for n in range(3)]).to(dev)
    W = max(int(g["s%d.wrong_ids" % n].numel()) for n in range(3))
    wrong = torch.full((3, W), -12345, dtype=torch.int64)
    cnt = torch.zeros(3, dtype=torch.int32)
    for n in range(3):
        w = g["s%d.wrong_ids" % n]
        wrong[n, : w.numel()] = w
        cnt[n] = w.numel()
    points, lrfs, act, err = QF.neighbour_gather(ps, ls, idx, wrong.to(dev), cnt.to(dev))
    assert int(err.item()) == 0
    for n in range(3):
        assert torch.equal(points[n].cpu(), g["s%d.points" % n])
        assert torch.equal(lrfs[n].cpu(), g["s%d.lrfs" % n])
        assert torch.equal(act[n].cpu(), g["s%d.act" % n])


# ---------------------------------------------------------------- losses / eval ops as single launches (8 f.3, f.4)
@pytest.mark.parametrize("B,C", [(16, 40), (5, 10), (1, 3), (300, 7)])
def test_class_pose_loss_kernel(B, C):
    """qec_class_pose_loss against the reference expressions in float64: Siamese spread loss
    (qec_sia_net.py:69-77) + 0.1 x pose_diff_loss (qec_sia_net.py:79-82) of the GT-class poses selected as in
    train_cls_sia.py:72-78, values and both gradients; also the one-branch and the pose-only forms"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(B * 31 + C)
    x = torch.rand(2, B, C, generator=gen) * 0.3 + 0.5
    y = torch.randint(0, C, (B,), generator=gen)
    pose = torch.randn(2, B, C, 4, generator=gen)
    pose = pose / pose.norm(dim=-1, keepdim=True)
    pose[1, 0, y[0]] = pose[0, 0, y[0]]           # <p0,p1> = 1: beyond the 0.9999 clamp, zero gradient
    if B > 2:
        pose[1, 1, y[1]] = -pose[0, 1, y[1]]      # and its mirror image
    xr = x.double().requires_grad_(True)
    pr = pose.double().requires_grad_(True)
    spread = orc.sia_spread_loss(xr, y)
    sel = torch.stack([pr[:, i, y[i]] for i in range(B)], 1)   # the reference's selection loop
    pd = orc.pose_diff_loss(sel)
    (spread + 0.1 * pd).backward()
    xd = x.to(dev).requires_grad_(True)
    pdv = pose.to(dev).requires_grad_(True)
    out = QF.class_pose_loss(xd, y.to(dev), pdv, margin=0.1, w_pose=0.1)
    (out[0] * 2.0).backward()
    assert abs(float(out[0]) - float(spread + 0.1 * pd)) < 2e-6
    assert abs(float(out[1]) - float(spread)) < 2e-6 and abs(float(out[2]) - float(pd)) < 2e-6
    assert max_err(xd.grad.cpu(), 2.0 * xr.grad) < 2e-6
    # d acos / dx ~ 70 at the clamp: relative tolerance on the pose gradient
    assert rel_err(pdv.grad.cpu(), 2.0 * pr.grad) < 1e-5
    # one branch, no pose term == QecNet.spread_loss
    x1 = x[0].to(dev).requires_grad_(True)
    l1 = QF.class_pose_loss(x1.unsqueeze(0), y.to(dev))[0]
    ref1 = orc.spread_loss(x[0].double(), y)
    assert abs(float(l1) - float(ref1)) < 2e-6
    # pose term only on pre-selected poses == pose_diff_loss(pose_out_act)
    import qecnetworks_b200 as qb
    net = qb.QecSiaNet(2048, 4, C, 3)
    got = net.pose_diff_loss(sel.detach().float().to(dev))
    assert abs(float(got) - float(pd)) < 2e-6
    tot, cls, pl = net.class_pose_loss(x.to(dev), pose.to(dev), torch.stack([y, y], 1).to(dev))
    assert abs(float(tot) - float(spread + 0.1 * pd)) < 2e-6 and abs(float(pl) - 0.1 * float(pd)) < 2e-6
    bad = y.clone()
    bad[0] = C
    QF.class_pose_loss(x.to(dev), bad.to(dev), pose.to(dev))
    # (out-of-range labels set the device flag instead of raising: no host sync on the training path)


def test_knn_then_gather_composition():
    """qec_knn -> qec_gather composed through knn_neighbourhoods: the id contracts differ (plain point numbers vs
    the loader's "id <= 0 inactive, last row zeroed"), so point 0 and point N-1 must survive the composition"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(3)
    B, N, P, K = 2, 64, 8, 9
    pts = torch.randn(B, N, 3, generator=gen)
    lr = torch.randn(B, N, 4, generator=gen)
    ci = torch.stack([torch.tensor([0, N - 1] + list(range(5, 5 + P - 2))) for _ in range(B)])
    gp, gl, ga, idx, err = QF.knn_neighbourhoods(pts.to(dev), lr.to(dev), K, ci.to(dev))
    assert int(err.item()) == 0
    ref = orc.knn_indices(pts, K, centre_idx=ci)
    assert torch.equal(idx.cpu(), ref)
    for b in range(B):
        assert torch.equal(gp[b].cpu(), pts[b][ref[b].reshape(-1)].view(P, K, 3))
        assert torch.equal(gl[b].cpu(), lr[b][ref[b].reshape(-1)].view(P, K, 4))
    assert bool((ga == 1).all())
    # the centre itself is its own nearest neighbour, also for point 0 and point N-1
    assert torch.equal(gp[:, :, 0].cpu(), torch.stack([pts[b][ci[b]] for b in range(B)]))
    # fewer points than K: the missing slots read zeros with activation 0
    gp2, gl2, ga2, idx2, _ = QF.knn_neighbourhoods(pts[:, :5].contiguous().to(dev), lr[:, :5].contiguous().to(dev), K,
                                                   ci[:, :1].clamp(max=4).to(dev))
    assert bool((idx2[..., 5:] == -1).all()) and float(ga2[..., 5:].abs().max()) == 0.0
    assert float(gp2[..., 5:, :].abs().max()) == 0.0 and bool((ga2[..., :5] == 1).all())


def test_quat_eval_kernel():
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(7, 33, 4, generator=gen)
    b = torch.randn(7, 33, 4, generator=gen)
    a = a / a.norm(dim=-1, keepdim=True)
    b = b / b.norm(dim=-1, keepdim=True)
    b[0, 0] = a[0, 0]
    dist, rel = QF.quat_eval(a.to(dev), b.to(dev))
    assert dist.shape == (7, 33) and rel.shape == (7, 33, 4)
    assert max_err(dist.cpu(), orc.q_distance(a.double(), b.double())) < 5e-4  # acos at 1 is ill-conditioned in FP32
    assert max_err(dist.cpu()[1:], orc.q_distance(a.double(), b.double())[1:]) < 1e-5
    assert max_err(rel.cpu(), orc.ham(orc.conj(a.double()), b.double())) < 1e-6


@pytest.mark.parametrize("K,I,O", [(256, 16, 8), (256, 128, 3)])
def test_routing_fused_uniform_random_votes_residual(K, I, O):
    """ill-conditioned cluster-kernel case: LRFs and transforms uniformly random, so every vote
    is a uniformly random rotation and M = mean v v^T is I/4 + O(1/sqrt(J)): relative eigengaps of a few 1e-3
    (the slow case of the squaring solver: eigenvalue ratio ~0.99).  T = 1, so the device and the float64 oracle
    solve the same matrix: eigen-residual and eigenvalue within 1e-5 for every capsule; the vector itself within
    1e-4 only where the relative gap is >= 1e-2"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    B, P, T = 1, 2, 1
    J = K * I
    gen = torch.Generator().manual_seed(K + I + O)
    lrfs = torch.randn(B, P, K, I, 4, generator=gen)
    lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
    lrfs = lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)
    act = torch.ones(B, P, J)
    t_raw = torch.randn(B * P * K, 4 * I * O, generator=gen)
    perm = _oqi_perm(I, O)
    alpha, beta = torch.tensor([1.0]), torch.tensor([1.0])
    v = orc.votes_from(lrfs.double(), t_raw.double().view(B, P, K, 4, I, O))
    M = torch.einsum("bpojc,bpojd->bpocd", v, v) / J
    lam, V = torch.linalg.eigh(M)
    gap = (lam[..., 3] - lam[..., 2]) / lam[..., 3]
    pose, agr, _, _ = QF.routing_fused_forward_raw(t_raw[:, perm].contiguous().to(dev), lrfs.to(dev), act.to(dev),
                                                   alpha.to(dev), beta.to(dev), T)
    p = pose.cpu().double()
    assert float((p.norm(dim=-1) - 1).abs().max()) < 1e-5 and float(p[..., 0].min()) >= 0.0
    Mp = torch.einsum("bpocd,bpod->bpoc", M, p)
    lg = (p * Mp).sum(-1)
    resid = (Mp - lg.unsqueeze(-1) * p).abs().amax(-1)
    ref = V[..., :, 3]
    d = torch.minimum((p - ref).abs().amax(-1), (p + ref).abs().amax(-1))
    well = gap >= 1e-2
    print("uniform votes J=%d: gaps %.1e .. %.1e, residual %.2e, dlambda %.2e, pose err (gap >= 1e-2: %d of %d) %.2e, all %.2e"
          % (J, float(gap.min()), float(gap.max()), float(resid.max()), float((lg - lam[..., 3]).abs().max()),
             int(well.sum()), well.numel(), float(d[well].max()) if well.any() else 0.0, float(d.max())))
    assert float(resid.max()) < 1e-5
    assert float((lg - lam[..., 3]).abs().max()) < 1e-5
    if well.any():
        assert float(d[well].max()) < TOL


# ---------------------------------------------------------------- third-generation cluster forward (opt-in)
@pytest.mark.parametrize(
    "B,P,K,I,O,T",
    [
        (2, 1, 32, 128, 5, 3),     # one CTA per capsule pair, every slot full, odd number of capsules per patch
        (1, 1, 256, 128, 6, 3),    # J = 32768: clusters of 8, the geometry of pool2 at BASELINE config 2
        (1, 2, 100, 30, 3, 2),     # ragged slice, T = 2
        (2, 1, 37, 6, 5, 1),       # T = 1, tiny
        (1, 1, 250, 100, 3, 3),    # J = 25000: ragged split over 8 CTAs
    ],
)
def test_routing_fused_pp_forward(B, P, K, I, O, T):
    """qec_routing_fused_pp_fwd (two capsules in flight per cluster, solver warps; csrc/routing_fused3.cu) against the
    float64 oracle and against the second-generation kernel, and bit-reproducible run to run"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    from qecnetworks_b200 import _lib

    assert _lib.query("qec_routing_fused_pp_supported", K, I, O, T) == 1
    gen = torch.Generator().manual_seed(K * 5 + I)
    J = K * I
    lrfs = clustered((B, P, K, I), gen, spread=0.3)
    act = (torch.rand(B, P, J, generator=gen) * 0.6 + 0.3) * (torch.rand(B, P, J, generator=gen) < 0.85)
    act[:, :, 0] = 1.0
    lrfs = lrfs * (act.view(B, P, K, I, 1) != 0)
    base = torch.randn(1, 4, 1, O, generator=gen)
    t_raw = (base + 0.25 * torch.randn(B * P * K, 4, I, O, generator=gen)).reshape(B * P * K, 4 * I * O)
    alpha, beta = torch.tensor([1.2]), torch.tensor([0.8])
    v = orc.votes_from(lrfs.double(), t_raw.double().view(B, P, K, 4, I, O))
    pose_ref, agr_ref = orc.routing(v, act.double(), alpha.double(), beta.double(), T)
    args = (t_raw[:, _oqi_perm(I, O)].contiguous().to(dev), lrfs.to(dev), act.to(dev), alpha.to(dev), beta.to(dev), T)
    prev = QF.set_fused_pp(0)
    try:
        ref2 = QF.routing_fused_forward_raw(*args)
        QF.set_fused_pp(1)
        got = QF.routing_fused_forward_raw(*args)
        again = QF.routing_fused_forward_raw(*args)
    finally:
        QF.set_fused_pp(prev)
    assert quat_err(got[0].cpu(), pose_ref) < TOL and max_err(got[1].cpu(), agr_ref) < TOL
    for a, b in zip(got, ref2):      # pose, agreement, poses_all, stats: the two generations agree to rounding
        assert rel_err(a.cpu(), b.cpu()) < 2e-6
    assert all(torch.equal(a, b) for a, b in zip(got, again))


# ---------------------------------------------------------------- tcgen05 backward of quater_gen2
@pytest.mark.parametrize("M,N,Kc", [(128, 128, 40), (256, 384, 7), (512, 20480, 40), (2048, 5120, 47)])
def test_gen2_backward_tcgen05(M, N, Kc):
    """qec_gen2_bwd (TMA + tcgen05 tf32 x3, accumulator in tensor memory; csrc/gen2_bwd.cu) against float64 matmuls:
    dL/dh = g Wp, dL/dW2 = (g^T h) scattered through the row permutation, dL/db2 = column sums of g -- the backward of
    the reference's quater_gen2 = nn.Linear(O, 4*I*O) (qec_module.py:124), within 2e-5 of the largest entry"""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    assert QF.gen2_bwd_supported(M, N, Kc)
    gen = torch.Generator().manual_seed(M + N + Kc)
    g = torch.randn(M, N, generator=gen).to(dev) * torch.rand(M, 1, generator=gen).to(dev)
    h = torch.randn(M, Kc, generator=gen).to(dev)
    W2 = (torch.randn(N, Kc, generator=gen) / 6).to(dev)
    perm = torch.randperm(N, generator=gen).to(dev)
    g_h, g_W2, g_b2 = QF.gen2_backward(g, h, W2, perm)
    ref_h = g.double() @ W2.double()[perm]
    ref_w = torch.empty(N, Kc, device=dev, dtype=torch.float64)
    ref_w[perm] = g.double().t() @ h.double()
    ref_b = torch.empty(N, device=dev, dtype=torch.float64)
    ref_b[perm] = g.double().sum(0)
    for got, ref, name in ((g_h, ref_h, "dh"), (g_W2, ref_w, "dW2"), (g_b2, ref_b, "db2")):
        e = rel_err(got.cpu(), ref.cpu())
        print("gen2 margin %s M=%d N=%d err=%.2e tol=2e-5" % (name, M, N, e))
        assert e < 2e-5, (name, e)
    only_h = QF.gen2_backward(g, h, W2, perm, need_w=False)
    assert only_h[1] is None and torch.equal(only_h[0], g_h)
    again = QF.gen2_backward(g, h, W2, perm)
    assert all(torch.equal(a, b) for a, b in zip(again, (g_h, g_W2, g_b2)))   # no atomics: bit-reproducible
    assert not QF.gen2_bwd_supported(100, 128, 40) and not QF.gen2_bwd_supported(128, 128, 48)


@pytest.mark.parametrize("M,C,O", [(8192, 384, 40), (100, 9, 5), (130, 7, 64), (64, 130, 17), (1, 3, 1)])
def test_linear_small_vs_float64(M, C, O):
    """qec_linear_fwd / _bwd (quater_gen1, reference qec_module.py:123 = nn.Linear + its autograd backward) against a
    float64 matmul: full and ragged row blocks, in-feature counts off the 4-float grid, O up to the 64 served"""
    from qecnetworks_b200 import functional as QF

    DEV = cuda()
    gen = torch.Generator().manual_seed(M * 7 + C)
    x = torch.randn(M, C, generator=gen).to(DEV).requires_grad_(True)
    W = (torch.randn(O, C, generator=gen) / C ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(O, generator=gen).to(DEV).requires_grad_(True)
    g = torch.randn(M, O, generator=gen).to(DEV)
    assert QF.linear_supported(C, O)
    y = QF.LinearSmall.apply(x, W, b)
    y.backward(g)
    xd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (x, W, b))
    yd = torch.nn.functional.linear(xd, Wd, bd)
    yd.backward(g.double())
    for name, got, ref in (("y", y, yd), ("g_x", x.grad, xd.grad), ("g_W", W.grad, Wd.grad), ("g_b", b.grad, bd.grad)):
        e = float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        print("linear margin %s M=%d C=%d O=%d err=%.2e tol=1.00e-05" % (name, M, C, O, e))
        assert e < 1e-5, (name, e)
    # partial outputs: input gradient only / parameter gradients only
    x2 = x.detach().clone().requires_grad_(True)
    QF.LinearSmall.apply(x2, W.detach(), b.detach()).backward(g)
    assert torch.equal(x2.grad, x.grad)
    W2 = W.detach().clone().requires_grad_(True)
    QF.LinearSmall.apply(x.detach(), W2, b.detach()).backward(g)
    assert torch.equal(W2.grad, W.grad)


@pytest.mark.parametrize("R,O,C", [(512, 128, 3), (20, 5, 3), (64, 33, 12), (4, 1, 16)])
def test_compose_linears_vs_float64(R, O, C):
    """qec_compose_fwd / _bwd: Wc = W2 W1, bc = W2 b1 + b2 (reference qec_module.py:123-124 composed) and its adjoint"""
    from qecnetworks_b200 import functional as QF

    DEV = cuda()
    gen = torch.Generator().manual_seed(R + O + C)
    W1 = torch.randn(O, C, generator=gen).to(DEV).requires_grad_(True)
    b1 = torch.randn(O, generator=gen).to(DEV).requires_grad_(True)
    W2 = (torch.randn(R, O, generator=gen) / O ** 0.5).to(DEV).requires_grad_(True)
    b2 = torch.randn(R, generator=gen).to(DEV).requires_grad_(True)
    gW = torch.randn(R, C, generator=gen).to(DEV)
    gb = torch.randn(R, generator=gen).to(DEV)
    assert QF.compose_supported(C)
    Wc, bc = QF.ComposeLinears.apply(W1, b1, W2, b2)
    torch.autograd.backward([Wc, bc], [gW, gb])
    d = [t.detach().double().requires_grad_(True) for t in (W1, b1, W2, b2)]
    Wcd, bcd = d[2] @ d[0], d[2] @ d[1] + d[3]
    torch.autograd.backward([Wcd, bcd], [gW.double(), gb.double()])
    pairs = [("Wc", Wc, Wcd), ("bc", bc, bcd)] + [("g_" + n, t.grad, td.grad) for n, t, td in zip(("W1", "b1", "W2", "b2"), (W1, b1, W2, b2), d)]
    for name, got, ref in pairs:
        e = float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        print("compose margin %s R=%d O=%d C=%d err=%.2e tol=1.00e-05" % (name, R, O, C, e))
        assert e < 1e-5, (name, e)


@pytest.mark.parametrize("M,N,Kc,use_perm", [(128, 128, 40, True), (256, 384, 10, False), (384, 640, 47, True),
                                              (128, 256, 1, True), (8192, 20480, 40, True)])
def test_gen2_forward_tcgen05(M, N, Kc, use_perm):
    """qec_gen2_fwd (quater_gen2 forward, reference qec_module.py:124 = nn.Linear with bias, output columns permuted)
    on the tcgen05 tensor cores against a float64 matmul: FP32 accuracy from the 3-term TF32 split.  Shapes: one
    tile, several tiles per CTA with a row-block change, contraction lengths 2 .. 48 (one to six k steps, both
    shared-memory panels), the BASELINE config-2 size (10 240 tiles over 148 persistent CTAs)."""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    gen = torch.Generator().manual_seed(M + N + Kc)
    h = torch.randn(M, Kc, generator=gen).to(dev)
    W2 = (torch.randn(N, Kc, generator=gen) / Kc ** 0.5).to(dev)
    b2 = torch.randn(N, generator=gen).to(dev)
    perm = torch.randperm(N, generator=gen).to(dev) if use_perm else None
    assert QF.gen2_fwd_supported(M, N, Kc)
    t = QF.gen2_forward(h, W2, b2, perm)
    Wd, bd = (W2.double(), b2.double()) if perm is None else (W2.double()[perm], b2.double()[perm])
    ref = h.double() @ Wd.t() + bd
    e = float((t.double() - ref).abs().max() / ref.abs().max())
    print("gen2fwd margin t M=%d N=%d Kc=%d err=%.2e tol=2.00e-06" % (M, N, Kc, e))
    assert e < 2e-6, e
    assert not QF.gen2_fwd_supported(100, 128, 40) and not QF.gen2_fwd_supported(128, 128, 64)


def test_routing_fused_backward_bit_reproducible():
    """dL/dt of the cluster-resident backward has no atomics on its path: the same launch must give the same bits every
    time, also with another kernel (qec_gen2_bwd: different shared-memory carve-out, cold L2 / TLB) running between
    the repetitions.  This is the regression guard of a race found in round 2 (operands staged in the reducer's
    scratch were overwritten by a warp a whole sweep ahead: one k x 32 channels x 4 components wrong about once per
    1 000 launches; scripts/race_hunt_bwd.py is the long form of this test)."""
    from qecnetworks_b200 import functional as QF

    dev = cuda()
    B, P, K, I, O, T = 2, 1, 256, 128, 40, 3
    g = torch.Generator().manual_seed(3)
    lrfs = clustered((B, P, K, I), g, spread=0.3).to(dev)
    act = (torch.rand(B, P, K * I, generator=g) * 0.5 + 0.4).to(dev)
    tb = torch.randn(1, O, 4, 1, generator=g)
    t = (tb + 0.25 * torch.randn(B * P * K, O, 4, I, generator=g)).reshape(B * P * K, 4 * I * O).to(dev)
    alpha, beta = torch.ones(1, device=dev), torch.ones(1, device=dev)
    Gp = torch.randn(B, P, O, 4, generator=g).to(dev)
    Ga = torch.randn(B, P, O, generator=g).to(dev)
    h = torch.randn(B * P * K, O, generator=g).to(dev)
    W2 = (torch.randn(4 * I * O, O, generator=g) / 6).to(dev)
    perm = torch.randperm(4 * I * O, generator=g).to(dev)
    pose, agr, poses_all, stats = QF.routing_fused_forward_raw(t, lrfs, act, alpha, beta, T)
    ref = ref2 = None
    for r in range(2000):
        gt, gl, ga, _ = QF.routing_fused_backward_raw(t, lrfs, act, alpha, beta, poses_all, stats, Gp, Ga, T, True, True)
        outs = QF.gen2_backward(gt, h, W2, perm)
        if ref is None:
            ref, ref2 = gt.clone(), [x.clone() for x in outs]
            continue
        assert torch.equal(gt, ref), "dL/dt differs in repetition %d (%d elements)" % (r, int((gt != ref).sum()))
        assert all(torch.equal(x, y) for x, y in zip(outs, ref2)), "qec_gen2_bwd differs in repetition %d" % r
