"""CPU oracle for the QEC-Net quaternion dynamic-routing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``qecnetworks_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm, never as the product.

This is a plain-PyTorch (CPU, dtype-generic: float32 or float64) restatement of
what the reference computes on the hot path.  It is *not* a copy of the
reference: it is written from the behavioural specification in SURVEY.md
Appendix A/B and each function cites the reference lines it restates
(paths relative to the reference checkout):

  models/qec_module.py:35-194      QecModule (LRF mean, votes, routing, activation)
  models/quat_ops.py:30-68         qmul (Hamilton product), qrotv
  models/qec_net.py:31-64          QecNet glue, spread_loss, pose_diff_loss
  models/qec_sia_net.py:32-80      QecSiaNet glue
  models/pytorch-autograd-solver/torch_autograd_solver/__init__.py:68-98
  models/pytorch-autograd-solver/torch_autograd_solver.cpp:100-169   symeig fwd/bwd
  my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160     neighbour gather

Pinned: ``tests/golden/make_golden.py`` runs the *unmodified* reference modules
(imported from the reference checkout with a 5-line ``torch_autograd_solver``
shim, the 1e-6 eigen-noise zeroed) and stores their inputs/outputs/gradients
in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those fixtures.  The forward eigensolve of the reference is the
closed-source cuSOLVER ``cusolverDnSsyevjBatched`` (version unpinned); it is
restated as ``torch.linalg.eigh(UPLO='U')`` with rows-are-eigenvectors,
ascending order.

The one deliberate difference from the reference: the reference adds
``1e-6 * randn`` to every 4x4 matrix before the eigensolve
(qec_module.py:49-50,87-88).  That makes the reference nondeterministic at the
6e-3 level, so the oracle exposes it as ``noise`` (default 0).
"""
import math

import torch

# --------------------------------------------------------------------------
# quaternion algebra (models/quat_ops.py:30-68), (w, x, y, z) convention
# --------------------------------------------------------------------------


def ham(q, r):
    """Hamilton product q (x) r.  Restates quat_ops.qmul (quat_ops.py:30-48)."""
    q0, q1, q2, q3 = q.unbind(-1)
    r0, r1, r2, r3 = r.unbind(-1)
    return torch.stack(
        (
            q0 * r0 - q1 * r1 - q2 * r2 - q3 * r3,
            q0 * r1 + q1 * r0 + q2 * r3 - q3 * r2,
            q0 * r2 - q1 * r3 + q2 * r0 + q3 * r1,
            q0 * r3 + q1 * r2 - q2 * r1 + q3 * r0,
        ),
        dim=-1,
    )


def conj(q):
    """Quaternion conjugate; inlined in the reference at qec_module.py:106-110."""
    return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])


def fix(q):
    """Force the w>=0 hemisphere by multiplying with sign(w); sign(0)=0 zeroes
    the quaternion (qec_module.py:58-60, 94-96, 132-134, 143-145)."""
    return q * torch.sign(q[..., :1])


def qrotv(q, v):
    """Rotate v by q: v + 2 (w (u x v) + u x (u x v)), u = q.xyz
    (quat_ops.py:50-68)."""
    u = q[..., 1:]
    w = q[..., :1]
    uv = torch.linalg.cross(u, v, dim=-1)
    uuv = torch.linalg.cross(u, uv, dim=-1)
    return v + 2.0 * (w * uv + uuv)


def angular_distance(pose, votes):
    """d = (2/pi) acos(min(|<pose, v>|, 0.9999)) (qec_module.py:156-160).
    pose [..., 4] broadcast against votes [..., J, 4]."""
    dot = (pose.unsqueeze(-2) * votes).sum(-1)
    return (2.0 / math.pi) * torch.acos(torch.clamp(dot.abs(), max=0.9999))


# --------------------------------------------------------------------------
# eigen solver boundary
# --------------------------------------------------------------------------


def symeig_rows(M):
    """The L0 boundary ``torch_autograd_solver.symeig`` on CUDA: eigenvalues
    ascending, eigenvectors as ROWS of V (cuSOLVER column-major written into a
    row-major tensor), upper triangle of the row-major input
    (__init__.py:74-81; torch_cusolver.cpp:194-288)."""
    w, V = torch.linalg.eigh(M, UPLO="U")
    return w, V.transpose(-1, -2)


def symeig_rows_backward(gw, gV_rows, w, V_rows):
    """Restates batch_symeig_backward (torch_autograd_solver.cpp:133-152)
    WITHOUT the final upper-triangle fold (:154-169), which has an off-by-one
    in its chunk loop and does not change anything downstream of a symmetric
    construction.  Returns the unfolded gM = V (F o V^T gV) V^T
    + V diag(gw) V^T in the columns-are-eigenvectors convention."""
    V = V_rows.transpose(-1, -2)
    gV = gV_rows.transpose(-1, -2)
    F = w.unsqueeze(-2) - w.unsqueeze(-1)  # F[i, j] = w_j - w_i
    eye = torch.eye(w.shape[-1], dtype=torch.bool)
    F = torch.where(eye, torch.full_like(F, float("inf")), F).reciprocal()
    inner = F * (V.transpose(-1, -2) @ gV)
    gM = V @ inner @ V.transpose(-1, -2)
    if gw is not None:
        gM = gM + (V * gw.unsqueeze(-2)) @ V.transpose(-1, -2)
    return gM


_SAFE_DIAG = (1.0, 2.0, 3.0, 4.0)

# The eigen-solver the restated QecModule calls at the reference's ``S.symeig`` call sites
# (qec_module.py:52,90): [n,4,4] -> (w [n,4] ascending, V [n,4,4] rows = eigenvectors).  Tests may swap it
# (tests/test_gpu_module.py runs this restatement of the reference loop on the GPU around
# ``qecnetworks_b200.symeig``); the default is the LAPACK statement above.
SYMEIG = symeig_rows


def _masked_top(M, valid, noise=0.0, generator=None):
    """Top eigenvector (largest eigenvalue) of each valid 4x4, zeros elsewhere.

    The reference compacts the valid matrices with nonzero() before symeig and
    scatters the result back into zeros (qec_module.py:43-55, 72-93); here the
    invalid ones are replaced by a fixed well-separated diagonal so that eigh's
    autograd stays finite, and the result is masked."""
    safe = torch.diag(M.new_tensor(_SAFE_DIAG)).expand_as(M)
    Ms = torch.where(valid[..., None, None], M, safe)
    if noise:
        Ms = Ms + noise * torch.randn(Ms.shape, dtype=Ms.dtype, generator=generator)
    _, Vrows = SYMEIG(Ms.reshape(-1, 4, 4))
    return Vrows[:, -1, :].reshape(M.shape[:-1]) * valid[..., None].to(M.dtype)


# --------------------------------------------------------------------------
# QecModule pieces (models/qec_module.py)
# --------------------------------------------------------------------------


def lrf_mean(lrfs, act, noise=0.0):
    """mean_lrf_per_patch + averageQuaternions (qec_module.py:101-110, 35-61).

    lrfs [B,P,K,I,4], act [B,P,K,I] -> mean [B,P,I,4]: sign-fixed top
    eigenvector of the UNWEIGHTED (1/K) sum_k q q^T; patches/channels with no
    active neighbour give 0."""
    K = lrfs.shape[2]
    Mbar = torch.einsum("bpkic,bpkid->bpicd", lrfs, lrfs) / K
    valid = (act != 0).any(dim=2)
    return fix(_masked_top(Mbar, valid, noise))


def rotate_points(mean, points):
    """x[b,p,k,i] = qrotv(conj(mean[b,p,i]), points[b,p,k])
    (qec_module.py:115-120).  -> [B,P,K,I,3]"""
    B, P, K, _ = points.shape
    I = mean.shape[2]
    c = conj(mean)[:, :, None, :, :].expand(B, P, K, I, 4)
    p = points[:, :, :, None, :].expand(B, P, K, I, 3)
    return qrotv(c, p)


def transform_quats(x, W1, b1, W2, b2):
    """quater_gen2(quater_gen1(x)) (qec_module.py:123-124): two Linears with no
    nonlinearity between.  x [B,P,K,I,3] -> t_raw [B,P,K,4,I,O] (the reference's
    own feature order q*I*O + i*O + o, qec_module.py:126)."""
    B, P, K, I, _ = x.shape
    O = W1.shape[0]
    h = x.reshape(-1, I * 3) @ W1.t() + b1
    t = h @ W2.t() + b2
    return t.view(B, P, K, 4, I, O)


def votes_from(lrfs, t_raw):
    """get_votes after the Linears (qec_module.py:126-147):
    t = fix(normalise(t_raw)); v = fix(lrf (x) t); layout [B,P,O,K*I,4]."""
    B, P, K, _, I, O = t_raw.shape
    t = t_raw.permute(0, 1, 2, 5, 4, 3)  # [B,P,K,O,I,4]
    t = t / t.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    t = fix(t)
    v = ham(lrfs[:, :, :, None, :, :].expand(B, P, K, O, I, 4), t)
    v = v.permute(0, 1, 3, 2, 4, 5).reshape(B, P, O, K * I, 4)
    return fix(v)


def routing(votes, act, alpha, beta, num_iterations, noise=0.0, return_all=False):
    """The Weiszfeld/IRLS loop + activation epilogue (qec_module.py:172-194,
    64-98, 149-160).

    votes [B,P,O,J,4]; act [B,P,J] (broadcast over O); alpha, beta [1].
    Votes are detached everywhere except inside the last iteration's M and the
    final distance (qec_module.py:172,183-188)."""
    B, P, O, J, _ = votes.shape
    a = act[:, :, None, :].expand(B, P, O, J)
    mask = (a != 0).to(votes.dtype)
    valid = votes.sum(-1).sum(-1) != 0  # qec_module.py:72-73
    vd = votes.detach()
    b = a
    poses = []
    pose = None
    for it in range(num_iterations):
        last = it == num_iterations - 1
        vv = votes if last else vd
        w = b * mask
        w = w / w.abs().sum(-1, keepdim=True).clamp_min(1e-12)  # F.normalize(p=1)
        M = torch.einsum("bpoj,bpojc,bpojd->bpocd", w, vv, vv)
        pose = fix(_masked_top(M, valid, noise))
        poses.append(pose)
        if not last:
            b = b * a * (1.0 - angular_distance(pose, vd))
    neg = b * (1.0 - angular_distance(pose, votes))
    n = (b != 0).to(votes.dtype).sum(-1)
    m = (n != 0).to(votes.dtype)
    s = neg.sum(-1) / (n + (1.0 - m)) * m
    agreement = torch.sigmoid(alpha.view(1, 1, -1) * s + (beta.view(1, 1, -1) - 1.0)) * m
    if return_all:
        return pose, agreement, torch.stack(poses, 0)
    return pose, agreement


def qec_module_forward(params, points, lrfs, act, num_iterations, noise=0.0, return_all=False):
    """QecModule.forward (qec_module.py:167-194).

    params: dict with the reference's state_dict keys
      quater_gen1.weight [O,3I], quater_gen1.bias [O],
      quater_gen2.weight [4IO,O], quater_gen2.bias [4IO], alpha [1], beta [1].
    points [B,P,K,3], lrfs [B,P,K,I,4], act [B,P,K,I]
    -> pose [B,P,O,4] (O squeezed if 1, as pose.squeeze(-2) does), agreement [B,P,O]."""
    B, P, K, I, _ = lrfs.shape
    mean = lrf_mean(lrfs, act, noise)
    x = rotate_points(mean, points)
    t_raw = transform_quats(
        x,
        params["quater_gen1.weight"],
        params["quater_gen1.bias"],
        params["quater_gen2.weight"],
        params["quater_gen2.bias"],
    )
    v = votes_from(lrfs, t_raw)
    out = routing(
        v, act.reshape(B, P, K * I), params["alpha"], params["beta"], num_iterations, noise, return_all
    )
    pose = out[0].squeeze(-2)
    if return_all:
        return pose, out[1], {"mean": mean, "x": x, "t_raw": t_raw, "votes": v, "poses": out[2]}
    return pose, out[1]


# --------------------------------------------------------------------------
# networks (models/qec_net.py, models/qec_sia_net.py)
# --------------------------------------------------------------------------


def _sub(params, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in params.items() if k.startswith(prefix)}


def qec_net_forward(params, points4pool1, lrfs4pool1, act4pool1, num_iterations, noise=0.0):
    """QecNet.forward (qec_net.py:31-44).  params uses the reference state_dict
    keys pool1.* / pool2.*.  points [B,256,9,3], lrfs [B,256,9,4], act [B,256,9]."""
    lrfs = lrfs4pool1.unsqueeze(-2)
    act = act4pool1.unsqueeze(-1)
    pose1, a1 = qec_module_forward(_sub(params, "pool1."), points4pool1, lrfs, act, num_iterations, noise)
    points2 = points4pool1[:, :, 0, :].unsqueeze(1)
    pose2, a2 = qec_module_forward(
        _sub(params, "pool2."), points2, pose1.unsqueeze(1), torch.relu(a1.unsqueeze(1)), num_iterations, noise
    )
    return pose2.squeeze(), a2.squeeze()


def qec_sia_net_forward(params, points, lrfs, act, num_iterations, noise=0.0):
    """QecSiaNet.forward (qec_sia_net.py:32-59): the same two layers over the two
    branches [:,0] and [:,1], stacked on a new leading axis."""
    poses, acts = [], []
    for i in range(2):
        p, a = qec_net_forward(params, points[:, i], lrfs[:, i], act[:, i], num_iterations, noise)
        poses.append(p)
        acts.append(a)
    return torch.stack(poses, 0), torch.stack(acts, 0)


def spread_loss(x, target, margin=0.1):
    """QecNet.spread_loss (qec_net.py:53-59).  x [B,ncls], target [B] int64."""
    onehot = torch.zeros_like(x).scatter_(1, target.view(-1, 1), 1.0)
    act_t = (x * onehot).sum(dim=1, keepdim=True)
    loss = torch.relu(margin - (act_t - x)) ** 2 * (1.0 - onehot)
    return loss.sum(1).mean()


def sia_spread_loss(x, target, margin=0.1):
    """QecSiaNet.spread_loss (qec_sia_net.py:67-75).  x [2,B,ncls]."""
    return spread_loss(x[0], target, margin) + spread_loss(x[1], target, margin)


def pose_diff_loss(pose):
    """pose_diff_loss (qec_net.py:61-64, qec_sia_net.py:77-80)."""
    dot = (pose[0] * pose[1]).sum(-1)
    return ((2.0 / math.pi) * torch.acos(torch.clamp(dot.abs(), max=0.9999))).mean()


# --------------------------------------------------------------------------
# loader gather (my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160)
# --------------------------------------------------------------------------


def neighbour_gather(point_set, lrf_set, idx, wrong_ids=None):
    """idx [P,K] int64 original point ids, -1 = empty slot.
    act = [idx > 0] further zeroed where idx is in wrong_ids (:139-150);
    the last row of point_set/lrf_set is zeroed and -1 picks it (:154-160).
    Note idx == 0 gets act = 0 but a non-zero LRF (reference behaviour)."""
    point_set = point_set.clone()
    lrf_set = lrf_set.clone()
    act = torch.clamp(torch.sign(idx), min=0)
    if wrong_ids is not None and wrong_ids.numel():
        # the reference only scans the first pool2_size rows (:142-150), where
        # pool2_size = number of rows with a non-empty first slot
        pool2_size = int((act[:, 0] != 0).sum())
        bad = torch.isin(idx, wrong_ids)
        bad[pool2_size:] = False
        act = torch.where(bad, torch.zeros_like(act), act)
    point_set[-1] = 0
    lrf_set[-1] = 0
    flat = idx.reshape(-1)
    P, K = idx.shape
    return point_set[flat].view(P, K, 3), lrf_set[flat].view(P, K, 4), act.to(point_set.dtype)


# --------------------------------------------------------------------------
# synthetic ModelNet40-shaped inputs -- shared by tests/bench
# --------------------------------------------------------------------------


def ave_pose(pose):
    """test_rot_sia.py:16-29, batched restatement: top eigenvector of the mean outer product of the poses
    whose components do not sum to zero, sign-fixed.  pose [B,n,4] -> [B,4]"""
    out = []
    for b in range(pose.shape[0]):
        keep = pose[b].sum(-1) != 0
        q = pose[b][keep]
        M = (q.unsqueeze(2) * q.unsqueeze(1)).mean(0)
        w, V = torch.linalg.eigh(M, UPLO="U")
        v = V[:, 3]
        out.append(v * torch.sign(v[0]))
    return torch.stack(out, 0)


def q_distance(pose1, pose2):
    """test_rot_sia.py:32-41"""
    p1 = pose1 * torch.sign(pose1[..., :1])
    p2 = pose2 * torch.sign(pose2[..., :1])
    return 2.0 * torch.acos((p1 * p2).sum(-1).abs().clamp(max=1.0)) / math.pi


def knn_indices(points, K, centre_idx=None, centres=None):
    """K nearest neighbours of the patch centres, plain float32 statement (the checker of qec_knn).
    The reference has no neighbour search of its own: its neighbourhoods are read from index files
    written offline by my_dataloader/gen_downsample_m40.py:86-150 (scipy cdist + greedy grouping) and
    consumed by my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160.  SURVEY.md 8 f.2 fixes the
    on-device contract instead: squared distances in float32 with every operation rounded separately,
    d = ((cx-x)^2 + (cy-y)^2) + (cz-z)^2, neighbours ordered by (d, point index) -- a stable sort.
    points [B,N,3]; centre_idx [B,P] int64 or centres [B,P,3] -> idx [B,P,K] int64 (-1 where N < K)."""
    import numpy as np

    pts = points.detach().cpu().numpy().astype(np.float32)
    B, N, _ = pts.shape
    if centre_idx is not None:
        ci = centre_idx.detach().cpu().numpy()
        c = np.stack([pts[b][ci[b]] for b in range(B)], 0)
    else:
        c = centres.detach().cpu().numpy().astype(np.float32)
    P = c.shape[1]
    out = np.full((B, P, K), -1, np.int64)
    for b in range(B):
        dx = (c[b, :, None, 0] - pts[b, None, :, 0]).astype(np.float32)
        dy = (c[b, :, None, 1] - pts[b, None, :, 1]).astype(np.float32)
        dz = (c[b, :, None, 2] - pts[b, None, :, 2]).astype(np.float32)
        d = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32)
        d = (d + (dz * dz).astype(np.float32)).astype(np.float32)
        order = np.argsort(d, axis=1, kind="stable")[:, :K]
        out[b, :, : order.shape[1]] = order
    return torch.from_numpy(out)


def _small_rotation(rotvec):
    ang = rotvec.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    axis = rotvec / ang.clamp_min(1e-12)
    return torch.cat([torch.cos(half), torch.sin(half) * axis], dim=-1)


def synthetic_batch(B, P=256, K=9, ncls=40, seed=1234, clustered=True, p_active=0.9, dtype=torch.float32):
    """Seeded ModelNet40-shaped synthetic sample batch: patch centres on a
    radius-0.8 sphere, K neighbours = centre + N(0, 0.03^2); LRFs = one random
    base rotation per sample (x) small-angle noise (clustered) or uniformly
    random (ill-conditioned); activation Bernoulli(p_active) with slot 0 always
    on; inactive slots carry a zero LRF (mirrors loader :139-160)."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(B, P, 3, generator=g, dtype=torch.float64)
    centres = 0.8 * centres / centres.norm(dim=-1, keepdim=True)
    points = centres[:, :, None, :] + 0.03 * torch.randn(B, P, K, 3, generator=g, dtype=torch.float64)
    points[:, :, 0, :] = centres
    if clustered:
        base = torch.randn(B, 1, 1, 4, generator=g, dtype=torch.float64)
        base = base / base.norm(dim=-1, keepdim=True)
        small = _small_rotation(0.15 * torch.randn(B, P, K, 3, generator=g, dtype=torch.float64))
        lrfs = ham(base.expand(B, P, K, 4), small)
    else:
        lrfs = torch.randn(B, P, K, 4, generator=g, dtype=torch.float64)
    lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
    lrfs = lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)
    act = (torch.rand(B, P, K, generator=g) < p_active).to(torch.float64)
    act[:, :, 0] = 1.0
    lrfs = lrfs * act[..., None]
    labels = torch.randint(0, ncls, (B,), generator=g)
    return points.to(dtype), lrfs.to(dtype), act.to(dtype), labels


def init_module_params(in_channels, out_channels, seed=0, dtype=torch.float32):
    """nn.Linear default init (kaiming-uniform a=sqrt(5) == U(+-1/sqrt(fan_in)))
    for quater_gen1/2, alpha = beta = 1 (qec_module.py:24-33)."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        W = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        return W.to(dtype), b.to(dtype)

    W1, b1 = lin(out_channels, 3 * in_channels)
    W2, b2 = lin(4 * in_channels * out_channels, out_channels)
    return {
        "quater_gen1.weight": W1,
        "quater_gen1.bias": b1,
        "quater_gen2.weight": W2,
        "quater_gen2.bias": b2,
        "alpha": torch.ones(1, dtype=dtype),
        "beta": torch.ones(1, dtype=dtype),
    }


def init_net_params(inter_out_channels, num_of_class, seed=0, dtype=torch.float32):
    p1 = init_module_params(1, inter_out_channels, seed, dtype)
    p2 = init_module_params(inter_out_channels, num_of_class, seed + 1, dtype)
    out = {"pool1." + k: v for k, v in p1.items()}
    out.update({"pool2." + k: v for k, v in p2.items()})
    return out
