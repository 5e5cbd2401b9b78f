"""GPU parity of the drop-in QecModule / QecNet / QecSiaNet against the golden vectors
of the unmodified reference and against the CPU oracle (pytest -m gpu)."""
import pytest
import torch

from oracle import qec_oracle as orc
from tests.util import load_golden, params_of, quat_err, max_err, rel_err, grads_close

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
