"""Seeded synthetic ModelNet40-shaped batches for the bench and smoke
runs: there is no network for ModelNet40 itself and the reference's inputs are
FLARE LRF files that are not in the repository.

Per sample: P patch centres on a radius-0.8 sphere, K neighbour points = centre +
N(0, 0.03^2) with slot 0 the centre itself (qec_net.py:39 uses slot 0 as the pool2
point); LRFs = one random base rotation per sample (x) a small-angle perturbation
(rotation vector ~ N(0, 0.15^2)), i.e. clustered like FLARE frames on a smooth
surface, in the w>=0 hemisphere; activation Bernoulli(p_active) with slot 0 always
active; inactive slots carry a zero LRF as the loader produces them
(modelnet_with_lrf_sample_index_loader.py:139-160)."""
import torch


def _ham(q, r):
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


def synthetic_batch(B, P=256, K=9, ncls=40, seed=1234, p_active=0.9, noise=0.15, pin=False):
    """-> points [B,P,K,3], lrfs [B,P,K,4], activation [B,P,K] (float32, CPU), labels [B] int64"""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(B, P, 3, generator=g, dtype=torch.float64)
    centres = 0.8 * centres / centres.norm(dim=-1, keepdim=True)
    points = centres[:, :, None, :] + 0.03 * torch.randn(B, P, K, 3, generator=g, dtype=torch.float64)
    points[:, :, 0, :] = centres
    base = torch.randn(B, 1, 1, 4, generator=g, dtype=torch.float64)
    base = base / base.norm(dim=-1, keepdim=True)
    rv = noise * torch.randn(B, P, K, 3, generator=g, dtype=torch.float64)
    ang = rv.norm(dim=-1, keepdim=True)
    small = torch.cat([torch.cos(0.5 * ang), torch.sin(0.5 * ang) * rv / ang.clamp_min(1e-12)], dim=-1)
    lrfs = _ham(base.expand(B, P, K, 4), small)
    lrfs = lrfs / lrfs.norm(dim=-1, keepdim=True)
    lrfs = lrfs * torch.where(lrfs[..., :1] < 0, -1.0, 1.0)
    act = (torch.rand(B, P, K, generator=g) < p_active).to(torch.float64)
    act[:, :, 0] = 1.0
    lrfs = lrfs * act[..., None]
    labels = torch.randint(0, ncls, (B,), generator=g)
    out = [points.float(), lrfs.float(), act.float(), labels]
    if pin:
        out = [t.pin_memory() for t in out]
    return tuple(out)


def siamese_pair(points, lrfs, act, seed=99):
    """second branch = first rotated by one random unit quaternion per sample (qrotv on the
    points, Hamilton product on the LRFs), as the reference loaders' augmentation does
    (modelnet_with_lrf_sample_index_loader.py:112-119).  -> stacked [B,2,...] tensors"""
    g = torch.Generator().manual_seed(seed)
    B = points.shape[0]
    q = torch.randn(B, 1, 1, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[..., :1] < 0, -1.0, 1.0)
    qe = q.expand_as(lrfs)
    u = qe[..., 1:]
    uv = torch.linalg.cross(u, points, dim=-1)
    uuv = torch.linalg.cross(u, uv, dim=-1)
    points2 = points + 2.0 * (qe[..., :1] * uv + uuv)
    lrfs2 = _ham(qe, lrfs)
    lrfs2 = lrfs2 * torch.where(lrfs2[..., :1] < 0, -1.0, 1.0)
    return torch.stack([points, points2], 1), torch.stack([lrfs, lrfs2], 1), torch.stack([act, act], 1)
