#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED
reference modules (models/qec_module.py, qec_net.py, qec_sia_net.py,
quat_ops.py of tolgabirdal/qecnetworks) on CPU.

Run in the build container only (needs the reference checkout, default
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [/path/to/reference]

How the reference is made importable on torch 2.11 / CPU:
  * its two native packages cannot build (removed ATen/THC APIs) so
    ``torch_autograd_solver`` is replaced by a shim whose ``symeig`` is
    ``torch.linalg.eigh(UPLO='U')``, transposed to the cuSOLVER
    rows-are-eigenvectors convention for 3-D input;
  * ``Tensor.cuda`` is made a no-op (qec_module.py hard-codes ``.cuda()``);
  * the 1e-6 eigen-noise (qec_module.py:49-50,87-88) is zeroed by patching
    ``torch.randn_like`` for the duration of a reference call, because with it
    two runs of the reference itself differ by 6e-3.
Nothing from the reference is copied into the repo: only inputs, outputs and
gradients of its functions are stored.
"""
import os
import sys
import types
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

from oracle import qec_oracle as orc  # noqa: E402  (input generators only)


def _install_shim():
    shim = types.ModuleType("torch_autograd_solver")

    def symeig(a, b=None, **kw):
        w, V = torch.linalg.eigh(a, UPLO="U")
        return (w, V.transpose(-1, -2)) if a.dim() == 3 else (w, V)

    shim.symeig = symeig
    sys.modules["torch_autograd_solver"] = shim
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, os.path.join(REF, "models"))


@contextlib.contextmanager
def no_noise():
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: torch.zeros_like(t)
    try:
        yield
    finally:
        torch.randn_like = orig


def clustered_lrfs(shape, g, spread=0.3):
    """unit quaternions, w>=0, clustered around one random rotation per leading index."""
    base = torch.randn(shape[0], *([1] * (len(shape) - 1)), 4, generator=g, dtype=torch.float64)
    q = base + spread * torch.randn(*shape, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[..., :1] < 0, -1.0, 1.0)
    return q.float()


def np_dict(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def module_case(name, B, P, K, I, O, T, frac_act, grads, seed):
    import qec_module as ref_qm

    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mod = ref_qm.QecModule(I, O, num_iterations=T, num_neighbours=K, num_patches=P)
    points = 0.5 * torch.randn(B, P, K, 3, generator=g)
    lrfs = clustered_lrfs((B, P, K, I), g)
    if frac_act:
        act = torch.rand(B, P, K, I, generator=g) * 0.5 + 0.4
    else:
        act = torch.ones(B, P, K, I)
    # an inactive neighbour with a zero LRF, and one with act=0 but a live LRF
    act[0, 0, 1, :] = 0
    lrfs[0, 0, 1, :, :] = 0
    act[-1, -1, 2, 0] = 0
    if not frac_act:
        drop = torch.rand(B, P, K, I, generator=g) < 0.1
        drop[:, :, 0] = False
        act = act * (~drop)
        lrfs = lrfs * (~drop)[..., None]
    out = {"dims": np.array([B, P, K, I, O, T]), "points": points, "lrfs": lrfs, "act": act}
    for k, v in mod.state_dict().items():
        out["param." + k] = v.clone()
    if grads:
        points.requires_grad_(True)
        lrfs.requires_grad_(True)
        act.requires_grad_(True)
    with no_noise():
        pose, agr = mod(points, lrfs, act)
    out["pose"] = pose
    out["agreement"] = agr
    if grads:
        Gp = torch.randn(pose.shape, generator=g)
        Ga = torch.randn(agr.shape, generator=g)
        loss = (pose * Gp).sum() + (agr * Ga).sum()
        loss.backward()
        out["cot.pose"] = Gp
        out["cot.agreement"] = Ga
        out["grad.points"] = points.grad
        out["grad.lrfs"] = lrfs.grad
        out["grad.act"] = act.grad
        for k, v in mod.named_parameters():
            out["grad." + k] = v.grad
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **np_dict(out))
    print(name, "pose", tuple(pose.shape), "agr", tuple(agr.shape))


def net_case(name, sia, B, C, ncls, T, seed):
    import qec_net as ref_net
    import qec_sia_net as ref_sia

    torch.manual_seed(seed)
    net = (ref_sia.QecSiaNet if sia else ref_net.QecNet)(2048, C, ncls, T)
    points, lrfs, act, labels = orc.synthetic_batch(B, 256, 9, ncls, seed=seed)
    if sia:
        g = torch.Generator().manual_seed(seed + 7)
        q = torch.randn(B, 1, 1, 4, generator=g)
        q = q / q.norm(dim=-1, keepdim=True)
        q = q * torch.where(q[..., :1] < 0, -1.0, 1.0)
        import quat_ops as ref_q

        qe = q.expand(B, 256, 9, 4).contiguous()
        points2 = ref_q.qrotv(qe, points.contiguous())
        lrfs2 = ref_q.qmul(qe, lrfs.contiguous()) * act[..., None]
        lrfs2 = lrfs2 * torch.where(lrfs2[..., :1] < 0, -1.0, 1.0)
        points = torch.stack([points, points2], 1)
        lrfs = torch.stack([lrfs, lrfs2], 1)
        act = torch.stack([act, act], 1)
    out = {"dims": np.array([B, C, ncls, T]), "points": points, "lrfs": lrfs, "act": act, "labels": labels}
    for k, v in net.state_dict().items():
        out["param." + k] = v.clone()
    with no_noise():
        pose, a = net(points, lrfs, act, torch.zeros(B, 2, 1) if sia else None)
    out["pose"] = pose
    out["agreement"] = a
    tgt = labels.view(-1, 1)
    if sia:
        tgt = torch.stack([tgt, tgt], 1).view(B, 2)
    loss = net.spread_loss(a, tgt, 0)
    if sia:
        loss = loss + net.pose_diff_loss(pose)
    loss.backward()
    out["loss"] = loss.detach()
    for k, v in net.named_parameters():
        out["grad." + k] = v.grad
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **np_dict(out))
    print(name, "pose", tuple(pose.shape), "loss", float(loss.detach()))


def quat_case():
    import quat_ops as ref_q

    g = torch.Generator().manual_seed(11)
    q = torch.randn(64, 4, generator=g)
    r = torch.randn(64, 4, generator=g)
    v = torch.randn(64, 3, generator=g)
    qn = q / q.norm(dim=-1, keepdim=True)
    np.savez_compressed(
        os.path.join(HERE, "quat_ops.npz"),
        q=q.numpy(), r=r.numpy(), v=v.numpy(), qn=qn.numpy(),
        qmul=ref_q.qmul(q, r).numpy(), qrotv=ref_q.qrotv(qn, v).numpy(),
    )
    print("quat_ops")


def symeig_case():
    """The two constructions the reference's own package tests use
    (pytorch-cusolver/test.py:8-31, pytorch-autograd-solver/test.py:101-141,
    186-200): sum of 8 outer products of unit quaternions; gradient of the mean
    of the sign-fixed top eigenvector w.r.t. the raw quaternions through a
    per-matrix eigh loop (the side of the comparison that still runs); and the
    near-singular literal of test_bug."""
    import torch.nn.functional as F

    torch.manual_seed(4041)  # the seed test.py uses
    x = torch.randn(80, 8, 4)
    x.requires_grad_(True)
    q = F.normalize(x, p=2, dim=-1).view(-1, 4)
    cov = torch.bmm(q.unsqueeze(2), q.unsqueeze(1)).view(80, 8, 4, 4).sum(1)
    tops = []
    for i in range(cov.size(0)):
        w, V = torch.linalg.eigh(cov[i], UPLO="U")
        top = V[:, torch.argmax(w)]
        tops.append(top if top[0] >= 0 else -top)
    tops = torch.stack(tops)
    tops.mean().backward()
    w_all, V_all = sys.modules["torch_autograd_solver"].symeig(cov.detach())
    bug = torch.tensor(
        [[2.7378, -2.7378, -0.4963, -0.2952], [-2.7378, 2.7378, 0.4963, 0.2952],
         [-0.4963, 0.4963, 1.8648, 1.1091], [-0.2952, 0.2952, 1.1091, 0.6596]]
    )
    wb, Vb = sys.modules["torch_autograd_solver"].symeig(bug.unsqueeze(0))
    np.savez_compressed(
        os.path.join(HERE, "symeig.npz"),
        x=x.detach().numpy(), cov=cov.detach().numpy(), w=w_all.numpy(), V_rows=V_all.numpy(),
        top=tops.detach().numpy(), grad_x=x.grad.numpy(),
        bug=bug.numpy(), bug_w=wb.numpy(), bug_V_rows=Vb.numpy(),
    )
    print("symeig")


def main():
    _install_shim()
    torch.set_num_threads(8)
    quat_case()
    symeig_case()
    # general I>1 with fractional activations: the pool2 regime in miniature
    module_case("module_small", 2, 3, 5, 4, 6, 3, True, True, seed=101)
    # pool1 regime in miniature: I=1, 0/1 activations
    module_case("module_pool1like", 2, 8, 9, 1, 16, 3, False, True, seed=102)
    # one iteration / two iterations (loop edge cases)
    module_case("module_T1", 2, 2, 4, 2, 3, 1, True, True, seed=103)
    module_case("module_T2", 2, 2, 4, 2, 3, 2, True, True, seed=104)
    # out_channels == 1: pose.squeeze(-2) bites
    module_case("module_O1", 2, 2, 4, 3, 1, 3, True, True, seed=105)
    # BASELINE config 1 geometry at B=2, forward only
    module_case("module_cfg1_B2", 2, 64, 9, 1, 128, 3, False, False, seed=106)
    net_case("net_small", False, 2, 8, 5, 3, seed=201)
    net_case("sia_net_small", True, 2, 8, 5, 3, seed=202)


if __name__ == "__main__":
    main()
