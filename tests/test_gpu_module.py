"""GPU parity of the drop-in QecModule / QecNet / QecSiaNet against the golden vectors
of the unmodified reference and against the CPU oracle (pytest -m gpu)."""
import pytest
import torch

from oracle import qec_oracle as orc
from tests.util import load_golden, params_of, quat_err, max_err, rel_err, grads_close, sum_of_terms_scale

pytestmark = pytest.mark.gpu
TOL = 1e-4


def cuda():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch.device("cuda:0")


def make_module(g, dev):
    import qecnetworks_b200 as qb

    B, P, K, I, O, T = [int(v) for v in g["dims"]]
    mod = qb.QecModule(I, O, num_iterations=T, num_neighbours=K, num_patches=P)
    mod.load_state_dict(params_of(g), strict=True)  # the reference's own keys and shapes
    return mod.to(dev)


CASES = ["module_small", "module_pool1like", "module_T1", "module_T2", "module_O1"]


@pytest.mark.parametrize("name", CASES + ["module_cfg1_B2"])
def test_module_forward_golden(name):
    dev = cuda()
    g = load_golden(name)
    mod = make_module(g, dev)
    with torch.no_grad():
        pose, agr = mod(g["points"].to(dev), g["lrfs"].to(dev), g["act"].to(dev))
    assert pose.shape == g["pose"].shape and agr.shape == g["agreement"].shape
    assert quat_err(pose.cpu(), g["pose"]) < TOL
    assert max_err(agr.cpu(), g["agreement"]) < TOL


@pytest.mark.parametrize("name", CASES)
def test_module_backward_golden(name):
    dev = cuda()
    g = load_golden(name)
    mod = make_module(g, dev)
    points = g["points"].to(dev).requires_grad_(True)
    lrfs = g["lrfs"].to(dev).requires_grad_(True)
    act = g["act"].to(dev).requires_grad_(True)
    pose, agr = mod(points, lrfs, act)
    ((pose * g["cot.pose"].to(dev)).sum() + (agr * g["cot.agreement"].to(dev)).sum()).backward()
    assert rel_err(points.grad.cpu(), g["grad.points"]) < TOL
    assert rel_err(lrfs.grad.cpu(), g["grad.lrfs"]) < TOL
    assert rel_err(act.grad.cpu(), g["grad.act"]) < TOL
    for k, v in mod.named_parameters():
        assert rel_err(v.grad.cpu(), g["grad." + k]) < TOL, k


def _net(g, dev, sia):
    import qecnetworks_b200 as qb

    B, C, ncls, T = [int(v) for v in g["dims"]]
    net = (qb.QecSiaNet if sia else qb.QecNet)(2048, C, ncls, T)
    net.load_state_dict(params_of(g), strict=True)
    return net.to(dev)


def test_net_golden():
    dev = cuda()
    g = load_golden("net_small")
    net = _net(g, dev, False)
    pose, a = net(g["points"].to(dev), g["lrfs"].to(dev), g["act"].to(dev), None)
    assert quat_err(pose.cpu(), g["pose"]) < TOL
    assert max_err(a.cpu(), g["agreement"]) < TOL
    loss = net.spread_loss(a, g["labels"].to(dev), 0)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    got = {k: v.grad.cpu() for k, v in net.named_parameters()}
    assert grads_close(got, {k: g["grad." + k] for k in got}, TOL) == []


def test_sia_net_golden():
    dev = cuda()
    g = load_golden("sia_net_small")
    net = _net(g, dev, True)
    pose, a = net(g["points"].to(dev), g["lrfs"].to(dev), g["act"].to(dev), None)
    assert pose.shape == g["pose"].shape
    assert quat_err(pose.cpu(), g["pose"]) < TOL
    assert max_err(a.cpu(), g["agreement"]) < TOL
    tgt = torch.stack([g["labels"], g["labels"]], 1).to(dev)
    loss = net.spread_loss(a, tgt, 0) + net.pose_diff_loss(pose)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    got = {k: v.grad.cpu() for k, v in net.named_parameters()}
    assert grads_close(got, {k: g["grad." + k] for k in got}, TOL) == []


def test_net_vs_oracle_seeded():
    """same seeded synthetic ModelNet40-shaped inputs through the CUDA net and the
    float64 oracle (B=2, C=32, 10 classes)."""
    import qecnetworks_b200 as qb

    dev = cuda()
    params = orc.init_net_params(32, 10, seed=4)
    points, lrfs, act, labels = orc.synthetic_batch(2, ncls=10, seed=21)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    pose_ref, a_ref = orc.qec_net_forward(p64, points.double(), lrfs.double(), act.double(), 3)
    orc.spread_loss(a_ref, labels).backward()
    net = qb.QecNet(2048, 32, 10, 3)
    net.load_state_dict(params)
    net = net.to(dev)
    pose, a = net(points.to(dev), lrfs.to(dev), act.to(dev), None)
    assert quat_err(pose.cpu(), pose_ref) < TOL
    assert max_err(a.cpu(), a_ref) < TOL
    net.spread_loss(a, labels.to(dev)).backward()
    got = {k: v.grad.cpu() for k, v in net.named_parameters()}
    assert grads_close(got, {k: p64[k].grad for k in got}, TOL) == []


def test_net_vs_oracle_full_geometry():
    """BASELINE config 2 geometry per sample (C=128, 40 classes, T=3: pool1 = 256 x 9 votes per capsule,
    pool2 = 32 768 votes per capsule on the full 8-CTA clusters of the packed kernels) at B=2 against the
    float64 oracle: outputs and every parameter gradient."""
    import qecnetworks_b200 as qb

    dev = cuda()
    params = orc.init_net_params(128, 40, seed=9)
    points, lrfs, act, labels = orc.synthetic_batch(2, ncls=40, seed=33)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    pose_ref, a_ref = orc.qec_net_forward(p64, points.double(), lrfs.double(), act.double(), 3)
    orc.spread_loss(a_ref, labels).backward()
    net = qb.QecNet(2048, 128, 40, 3)
    net.load_state_dict(params)
    net = net.to(dev)
    pose, a = net(points.to(dev), lrfs.to(dev), act.to(dev), None)
    assert quat_err(pose.cpu(), pose_ref) < TOL
    assert max_err(a.cpu(), a_ref) < TOL
    net.spread_loss(a, labels.to(dev)).backward()
    got = {k: v.grad.cpu() for k, v in net.named_parameters()}
    assert grads_close(got, {k: p64[k].grad for k in got}, TOL) == []


def test_full_size_properties():
    """BASELINE config 2 shapes (B=32, C=128, 40 classes, T=3): size-independent
    properties -- unit-norm poses in the w>=0 hemisphere, activations in (0,1),
    rotation equivariance (poses become q (x) pose up to sign, activations
    unchanged), finite gradients for every parameter."""
    import qecnetworks_b200 as qb

    dev = cuda()
    torch.manual_seed(0)
    net = qb.QecNet(2048, 128, 40, 3).to(dev)
    points, lrfs, act, labels = orc.synthetic_batch(32, seed=5)
    points, lrfs, act, labels = points.to(dev), lrfs.to(dev), act.to(dev), labels.to(dev)
    pose, a = net(points, lrfs, act, None)
    assert pose.shape == (32, 40, 4) and a.shape == (32, 40)
    assert float((pose.norm(dim=-1) - 1).abs().max()) < 1e-5
    assert float(pose[..., 0].min()) >= 0.0
    assert float(a.min()) > 0.0 and float(a.max()) < 1.0
    net.spread_loss(a, labels).backward()
    for k, v in net.named_parameters():
        assert torch.isfinite(v.grad).all(), k
        assert float(v.grad.abs().max()) > 0.0, k
    q = torch.tensor([0.3, -0.5, 0.7, 0.4], device=dev)
    q = q / q.norm()
    qe = q.expand(32, 256, 9, 4)

    def ham(a_, b_):
        return orc.ham(a_, b_)

    points_r = orc.qrotv(qe, points)
    lrfs_r = ham(qe, lrfs)
    lrfs_r = lrfs_r * torch.where(lrfs_r[..., :1] < 0, -1.0, 1.0)
    with torch.no_grad():
        pose_r, a_r = net(points_r, lrfs_r, act, None)
    assert max_err(a_r.cpu(), a.detach().cpu()) < TOL
    assert quat_err(ham(q.expand_as(pose), pose.detach()).cpu(), pose_r.cpu()) < 5e-4


@pytest.mark.parametrize("K,O,T", [(9, 32, 1), (16, 64, 10), (32, 48, 5), (64, 512, 2), (9, 512, 7)])
def test_stress_sweep_corners(K, O, T):
    """BASELINE config 5 geometry (P=64, I=1, K in 9..64, O in 32..512, T in 1..10) at a small batch,
    inference only: the drop-in module against the oracle"""
    import qecnetworks_b200 as qb

    dev = cuda()
    B, P = 2, 64
    params = orc.init_module_params(1, O, seed=K + O)
    params["quater_gen2.bias"] = params["quater_gen2.bias"] + torch.tensor([1.5, 0.2, -0.3, 0.4]).repeat_interleave(O)
    points, lrfs, act, _ = orc.synthetic_batch(B, P=P, K=K, seed=K * 3 + T)
    pose_ref, agr_ref = orc.qec_module_forward({k: v.double() for k, v in params.items()}, points.double(),
                                               lrfs.unsqueeze(-2).double(), act.unsqueeze(-1).double(), T)
    mod = qb.QecModule(1, O, num_iterations=T, num_neighbours=K, num_patches=P)
    mod.load_state_dict(params)
    mod = mod.to(dev)
    with torch.no_grad():
        pose, agr = mod(points.to(dev), lrfs.unsqueeze(-2).to(dev), act.unsqueeze(-1).to(dev))
    assert quat_err(pose.cpu(), pose_ref) < TOL
    assert max_err(agr.cpu(), agr_ref) < TOL


def test_eval_ops_vs_oracle():
    """evaluation-side ops of test_rot_sia.py (ave_pose, q_distance, relative pose) batched on the device
    against their restatement (float64), incl. zeroed poses and the reference's sum-to-zero mask"""
    import qecnetworks_b200 as qb
    from qecnetworks_b200 import eval_ops as E

    dev = cuda()
    gen = torch.Generator().manual_seed(17)
    base = torch.randn(6, 1, 4, generator=gen)
    q = base + 0.3 * torch.randn(6, 300, 4, generator=gen)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[..., :1] < 0, -1.0, 1.0)
    q[:, ::7] = 0                                        # invalid LRFs
    q[0, 1] = torch.tensor([0.5, -0.5, 0.5, -0.5])      # non-zero but excluded by the reference's mask
    ref = orc.ave_pose(q.double())
    got = E.ave_pose(q.to(dev))
    assert quat_err(got.cpu(), ref) < 1e-5
    assert float(got[:, 0].min()) >= 0.0
    a, b = q[:, 3].to(dev), q[:, 5].to(dev)
    assert max_err(E.q_distance(a, b).cpu(), orc.q_distance(q[:, 3].double(), q[:, 5].double())) < 1e-5
    rel = E.relative_pose(a, b)
    assert max_err(rel.cpu(), orc.ham(orc.conj(q[:, 3]), q[:, 5])) < 1e-6
    assert qb.eval_ops is E


def test_training_trajectory_matches_oracle():
    """eight Adam steps of QecNet (C=16, 5 classes, B=2) on the device and in the float64 oracle from the
    same initialisation and batch: the loss trajectories agree step by step and the loss goes down --
    forward, backward and the update loop end to end"""
    import qecnetworks_b200 as qb

    dev = cuda()
    params = orc.init_net_params(16, 5, seed=2)
    points, lrfs, act, labels = orc.synthetic_batch(2, ncls=5, seed=8)
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    net = qb.QecNet(2048, 16, 5, 3)
    net.load_state_dict(params)
    net = net.to(dev)
    keys = [k for k, _ in net.named_parameters()]
    opt_d = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt_o = torch.optim.Adam([p64[k] for k in keys], lr=1e-3)
    pd, ld, ad, yd = points.to(dev), lrfs.to(dev), act.to(dev), labels.to(dev)
    got, ref = [], []
    for _ in range(8):
        opt_d.zero_grad()
        _, a = net(pd, ld, ad, None)
        loss = net.spread_loss(a, yd)
        loss.backward()
        opt_d.step()
        got.append(float(loss.detach()))
        opt_o.zero_grad()
        _, a_ref = orc.qec_net_forward(p64, points.double(), lrfs.double(), act.double(), 3)
        loss_ref = orc.spread_loss(a_ref, labels)
        loss_ref.backward()
        opt_o.step()
        ref.append(float(loss_ref.detach()))
    assert max(abs(g - r) for g, r in zip(got, ref)) < 1e-4 * max(1.0, max(ref)), (got, ref)
    assert got[-1] < got[0]


# ---------------------------------------------------------------- BASELINE config 4: the Siamese step at its stated shape
def _sia_inputs(B, ncls, seed):
    from qecnetworks_b200.synthetic import siamese_pair

    points, lrfs, act, labels = orc.synthetic_batch(B, ncls=ncls, seed=seed)
    p2, l2, a2 = siamese_pair(points, lrfs, act, seed=seed + 1)
    return p2, l2, a2, labels


def test_sia_net_vs_oracle_full_geometry():
    """BASELINE config 4 per-sample geometry (QecSiaNet, C=128, 40 classes, T=3; train_cls_sia.py:70-79: spread
    loss of both branches + 0.1 x pose_diff_loss of the GT-class poses) at B=2 pairs against the float64 oracle:
    outputs, the three loss values and every parameter gradient"""
    import qecnetworks_b200 as qb

    dev = cuda()
    params = orc.init_net_params(128, 40, seed=12)
    p2, l2, a2, labels = _sia_inputs(2, 40, seed=41)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    pose_ref, a_ref = orc.qec_sia_net_forward(p64, p2.double(), l2.double(), a2.double(), 3)
    a_ref.retain_grad()
    cls_ref = orc.sia_spread_loss(a_ref, labels)
    sel = torch.stack([pose_ref[:, i, labels[i]] for i in range(2)], 1)   # train_cls_sia.py:76-77
    pose_loss_ref = 0.1 * orc.pose_diff_loss(sel)
    (cls_ref + pose_loss_ref).backward()
    net = qb.QecSiaNet(2048, 128, 40, 3)
    net.load_state_dict(params)
    net = net.to(dev)
    pose, a = net(p2.to(dev), l2.to(dev), a2.to(dev), None)
    assert pose.shape == (2, 2, 40, 4) and a.shape == (2, 2, 40)
    assert quat_err(pose.cpu(), pose_ref) < TOL
    assert max_err(a.cpu(), a_ref) < TOL
    tgt = torch.stack([labels, labels], 1).to(dev)
    total, cls, pl = net.class_pose_loss(a, pose, tgt)
    assert abs(float(cls) - float(cls_ref)) < 1e-5 and abs(float(pl) - float(pose_loss_ref)) < 1e-5
    # the reference's own formulation (two calls + the selection loop) gives the same value
    sel_d = torch.stack([pose[:, i, labels[i]] for i in range(2)], 1)
    alt = net.spread_loss(a, tgt, 0) + 0.1 * net.pose_diff_loss(sel_d)
    assert abs(float(alt) - float(total)) < 1e-6
    total.backward()
    got = {k: v.grad.cpu() for k, v in net.named_parameters()}
    # d loss / d pool2.beta = sum over the 160 output capsules of terms that cancel to ~1e-2 of their magnitudes
    assert grads_close(got, {k: p64[k].grad for k in got}, TOL,
                       term_scales={"pool2.beta": sum_of_terms_scale(a_ref, a_ref.grad)}) == []


def test_sia_full_size_properties_and_rotation_error():
    """BASELINE config 4 at its stated size (B=16 pairs, C=128, 40 classes): branch 1 is branch 0 rotated by a
    known quaternion, so by equivariance pose_1 = q (x) pose_0 up to sign and the activations of the two branches
    agree; the relative-rotation error metric of test_rot_sia.py (batched on the device) is ~0 for every class;
    gradients finite and non-zero"""
    import qecnetworks_b200 as qb
    from qecnetworks_b200 import eval_ops as E
    from qecnetworks_b200.synthetic import siamese_pair

    dev = cuda()
    torch.manual_seed(1)
    net = qb.QecSiaNet(2048, 128, 40, 3).to(dev)
    points, lrfs, act, labels = orc.synthetic_batch(16, ncls=40, seed=77)
    gq = torch.Generator().manual_seed(78)
    q = torch.randn(16, 1, 1, 4, generator=gq)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[..., :1] < 0, -1.0, 1.0)
    p2, l2, a2 = siamese_pair(points, lrfs, act, seed=78)   # the same generator state: rotation q per sample
    pose, a = net(p2.to(dev), l2.to(dev), a2.to(dev), None)
    assert pose.shape == (2, 16, 40, 4) and a.shape == (2, 16, 40)
    assert float((pose.norm(dim=-1) - 1).abs().max()) < 1e-5
    assert max_err(a[0].detach().cpu(), a[1].detach().cpu()) < TOL
    err = E.relative_rotation_error(pose.detach(), labels.to(dev), q.view(16, 4).to(dev))
    assert err.shape == (16,)
    assert float(err.max()) < 2e-3, float(err.max())   # (2/pi) acos near 1: 1e-6 in the dot product is 9e-4 here
    total, _, _ = net.class_pose_loss(a, pose, torch.stack([labels, labels], 1).to(dev))
    total.backward()
    for k, v in net.named_parameters():
        assert torch.isfinite(v.grad).all(), k
        assert float(v.grad.abs().max()) > 0.0, k


# ---------------------------------------------------------------- robustness: ill-conditioned (uniform-random) LRFs
def _oracle_iteration_matrices(aux, act, T):
    """float64 M^t of every routing iteration of the oracle run (rebuilt from its votes and poses), their
    eigenvalues, and the relative gap of the top eigenvalue per iteration"""
    v = aux["votes"]
    B, P, O, J, _ = v.shape
    a = act.double().reshape(B, P, 1, J).expand(B, P, O, J)
    b = a.clone()
    Ms, lams, gaps = [], [], []
    for t in range(T):
        w = b / b.abs().sum(-1, keepdim=True).clamp_min(1e-12)
        M = torch.einsum("bpoj,bpojc,bpojd->bpocd", w, v, v)
        lam = torch.linalg.eigvalsh(M)
        Ms.append(M)
        lams.append(lam)
        gaps.append((lam[..., 3] - lam[..., 2]) / lam[..., 3].clamp_min(1e-30))
        if t < T - 1:
            b = b * a * (1.0 - orc.angular_distance(aux["poses"][t], v))
    return Ms, lams, torch.stack(gaps, 0)


@pytest.mark.parametrize("T", [1, 3])
def test_random_lrfs_gap_aware(T):
    """with uniform-random LRFs the routing matrices have eigengaps down to ~1e-4, where an FP32
    eigenvector is only defined to ~eps/gap.  Gap-aware check per output capsule against the float64 oracle:
    pose and agreement within 1e-4 where the relative gap of EVERY iteration is >= 1e-2 (an ill-conditioned early
    iteration moves the weights of the later ones); for T = 1, where the device and the oracle solve the same
    matrix, additionally for ALL capsules the eigen-residual |M p - lambda p| <= 1e-5 and
    |lambda - lambda_ref| <= 1e-5 (the vector may rotate inside a near-degenerate eigenspace, the residual may not
    grow).  Exercises the squaring solver's round caps on small gaps."""
    import qecnetworks_b200 as qb

    dev = cuda()
    B, P, K, O = 2, 64, 9, 64
    params = orc.init_module_params(1, O, seed=5)
    points, lrfs, act, _ = orc.synthetic_batch(B, P=P, K=K, seed=55, clustered=False)
    p64 = {k: v.double() for k, v in params.items()}
    pose_ref, agr_ref, aux = orc.qec_module_forward(p64, points.double(), lrfs.unsqueeze(-2).double(),
                                                    act.unsqueeze(-1).double(), T, return_all=True)
    mod = qb.QecModule(1, O, num_iterations=T, num_neighbours=K, num_patches=P)
    mod.load_state_dict(params)
    mod = mod.to(dev)
    with torch.no_grad():
        pose, agr = mod(points.to(dev), lrfs.unsqueeze(-2).to(dev), act.unsqueeze(-1).to(dev))
    pose = pose.cpu().double()
    Ms, lams, gaps = _oracle_iteration_matrices(aux, act, T)
    well = gaps.min(0).values >= 1e-2
    d = torch.minimum((pose - pose_ref).abs().amax(-1), (pose + pose_ref).abs().amax(-1))
    print("random LRFs T=%d: min gap %.2e, well-conditioned %d / %d, max pose err (well) %.2e, (all) %.2e"
          % (T, float(gaps.min()), int(well.sum()), well.numel(), float(d[well].max()), float(d.max())))
    assert int(well.sum()) > 0.5 * well.numel()
    assert float(d[well].max()) < TOL, float(d[well].max())
    assert max_err(agr.cpu()[well], agr_ref[well]) < TOL
    if T == 1:
        M, lam = Ms[0], lams[0]
        Mp = torch.einsum("bpocd,bpod->bpoc", M, pose)
        lam_got = (pose * Mp).sum(-1)
        resid = (Mp - lam_got.unsqueeze(-1) * pose).abs().amax(-1)
        print("   max eigen-residual %.2e, max |lambda - lambda_ref| %.2e"
              % (float(resid.max()), float((lam_got - lam[..., 3]).abs().max())))
        assert float(resid.max()) < 1e-5
        assert float((lam_got - lam[..., 3]).abs().max()) < 1e-5


def test_random_lrfs_full_net_residuals():
    """the same ill-conditioned distribution through the whole QecNet (C=32, 10 classes, B=2): the forward must
    stay finite, unit-norm and in the w >= 0 hemisphere, the gradients finite -- the pool2 solves see gaps down to
    a few 1e-4 here, beyond what a 1e-4 vector comparison can test"""
    import qecnetworks_b200 as qb

    dev = cuda()
    params = orc.init_net_params(32, 10, seed=6)
    points, lrfs, act, labels = orc.synthetic_batch(2, ncls=10, seed=56, clustered=False)
    net = qb.QecNet(2048, 32, 10, 3)
    net.load_state_dict(params)
    net = net.to(dev)
    pose, a = net(points.to(dev), lrfs.to(dev), act.to(dev), None)
    assert torch.isfinite(pose).all() and torch.isfinite(a).all()
    assert float((pose.norm(dim=-1) - 1).abs().max()) < 1e-5 and float(pose[..., 0].min()) >= 0.0
    net.spread_loss(a, labels.to(dev)).backward()
    for k, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), k


# ---------------------------------------------------------------- INTEGRATION.md section 3 executed
@pytest.mark.parametrize("name", ["module_small", "module_T2", "module_cfg1_B2"])
def test_reference_loop_around_our_symeig(name):
    """the reference's QecModule keeps its own ATen loop and only swaps the eigen-solver
    package for qecnetworks_b200.symeig.  Executed here on the GPU with the restatement of that loop
    (oracle.qec_module_forward: same call sites as qec_module.py:52,90) running in float32 ON THE DEVICE around
    qb.symeig -- forward against the golden vectors of the unmodified reference, and the gradients through
    BatchSymeig4.backward against its golden gradients"""
    import qecnetworks_b200 as qb

    dev = cuda()
    g = load_golden(name)
    B, P, K, I, O, T = [int(v) for v in g["dims"]]
    params = {k: v.to(dev).requires_grad_(True) for k, v in params_of(g).items()}
    lrfs = g["lrfs"].to(dev).requires_grad_(name != "module_cfg1_B2")
    prev = orc.SYMEIG
    orc.SYMEIG = lambda M: qb.symeig(M.contiguous())
    try:
        before = qb.kernel_launches()
        pose, agr = orc.qec_module_forward(params, g["points"].to(dev), lrfs, g["act"].to(dev), T)
        assert qb.kernel_launches() - before == 1 + T      # the reference's 1 + T S.symeig call sites
        assert quat_err(pose.cpu(), g["pose"]) < TOL
        assert max_err(agr.cpu(), g["agreement"]) < TOL
        if name != "module_cfg1_B2":
            ((pose * g["cot.pose"].to(dev)).sum() + (agr * g["cot.agreement"].to(dev)).sum()).backward()
            assert rel_err(lrfs.grad.cpu(), g["grad.lrfs"]) < TOL
            for k, v in params.items():
                assert rel_err(v.grad.cpu(), g["grad." + k]) < TOL, k
    finally:
        orc.SYMEIG = prev


def test_module_on_second_device_without_set_device():
    """a module moved with .to('cuda:1') while the current device stays cuda:0 must launch on cuda:1 (device
    guard in _lib.launch), like a native PyTorch op"""
    import qecnetworks_b200 as qb

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = load_golden("module_small")
    assert torch.cuda.current_device() == 0
    mod = make_module(g, torch.device("cuda:1"))
    with torch.no_grad():
        pose, agr = mod(g["points"].to("cuda:1"), g["lrfs"].to("cuda:1"), g["act"].to("cuda:1"))
    assert pose.device.index == 1 and torch.cuda.current_device() == 0
    assert quat_err(pose.cpu(), g["pose"]) < TOL and max_err(agr.cpu(), g["agreement"]) < TOL


@pytest.mark.parametrize("K,O,T", [(9, 32, 1), (64, 512, 10), (32, 128, 3)])
def test_stress_corners_full_batch(K, O, T):
    """BASELINE config 5 at its stated batch (B = 1024 inference, P = 64, I = 1): corners of the sweep at full size.
    Every capsule depends on one sample only, so three samples of the big batch are checked against the float64
    oracle run on those samples alone; plus unit norm / hemisphere / activation range over the whole output.
    K = 64, O = 512 is the corner whose vote tensor would be 34 GB: it runs on the regenerate-votes kernel."""
    import qecnetworks_b200 as qb

    dev = cuda()
    B, P = 1024, 64
    params = orc.init_module_params(1, O, seed=K + O)
    params["quater_gen2.bias"] = params["quater_gen2.bias"] + torch.tensor([1.5, 0.2, -0.3, 0.4]).repeat_interleave(O)
    points, lrfs, act, _ = orc.synthetic_batch(B, P=P, K=K, seed=K * 5 + T)
    mod = qb.QecModule(1, O, num_iterations=T, num_neighbours=K, num_patches=P)
    mod.load_state_dict(params)
    mod = mod.to(dev)
    before = torch.cuda.max_memory_allocated()
    with torch.no_grad():
        pose, agr = mod(points.to(dev), lrfs.unsqueeze(-2).to(dev), act.unsqueeze(-1).to(dev))
    assert pose.shape == (B, P, O, 4) and agr.shape == (B, P, O)
    # nothing of the size of the vote tensor (16 B x B P O K) was allocated
    assert torch.cuda.max_memory_allocated() - before < 3 * pose.numel() * 4 + (1 << 30)
    assert float((pose.norm(dim=-1) - 1).abs().max()) < 1e-5 and float(pose[..., 0].min()) >= 0.0
    assert float(agr.min()) > 0.0 and float(agr.max()) < 1.0
    pick = torch.tensor([0, 511, 1023])
    p64 = {k: v.double() for k, v in params.items()}
    pose_ref, agr_ref = orc.qec_module_forward(p64, points[pick].double(), lrfs[pick].unsqueeze(-2).double(),
                                               act[pick].unsqueeze(-1).double(), T)
    assert quat_err(pose[pick.to(dev)].cpu(), pose_ref) < TOL
    assert max_err(agr[pick.to(dev)].cpu(), agr_ref) < TOL


def test_general_path_inference_is_sliced_over_the_batch():
    """a geometry on the general path (I > 1, few votes per capsule): under no_grad a large batch runs in slices so
    that no per-vote temporary exceeds QecModule.INFERENCE_TEMP_BYTES, with results identical to the unsliced run"""
    import qecnetworks_b200 as qb

    dev = cuda()
    g = load_golden("module_small")
    mod = make_module(g, dev)
    rep = 37
    pts, lr, ac = [g[k].repeat(rep, *([1] * (g[k].dim() - 1))).to(dev) for k in ("points", "lrfs", "act")]
    with torch.no_grad():
        pose_a, agr_a = mod(pts, lr, ac)
        mod.INFERENCE_TEMP_BYTES = 32 * 3 * 5 * 4 * 6 * 5   # five samples per slice
        pose_b, agr_b = mod(pts, lr, ac)
    # (not bit-identical: cuBLAS picks its tiling by the GEMM's row count, i.e. by the slice size)
    assert quat_err(pose_a.cpu(), pose_b.cpu()) < 1e-5 and max_err(agr_a.cpu(), agr_b.cpu()) < 1e-5
    assert quat_err(pose_b[-2:].cpu(), g["pose"]) < TOL
