"""Evaluation-side quaternion ops of the reference's rotation test (test_rot_sia.py), batched on the
device.  Same names, argument meaning and results as the reference's
per-sample Python functions; the averaging runs on the same sm_100a kernel as the routing layer's
mean LRF (qec_lrf_mean_fwd: outer-product mean + top eigenvector + hemisphere fix)."""
import math

import torch

from . import functional as QF


def ave_pose(pose):
    """reference test_rot_sia.py:16-29: per sample, the top eigenvector of the mean outer product of the
    poses whose components do not sum to zero (the reference's mask, `torch.sum(pose[b], -1).nonzero()`),
    sign-fixed to w >= 0.  pose [B, n, 4] CUDA float32 -> [B, 4]."""
    B, n, _ = pose.shape
    pose = pose.float().contiguous()
    mask = (pose.sum(-1) != 0).to(pose.dtype)
    lrfs = (pose * mask.unsqueeze(-1)).view(B, 1, n, 1, 4)
    mean, _ = QF.lrf_mean_rotate(lrfs, mask.view(B, 1, n, 1), pose.new_zeros(B, 1, n, 3))
    return mean.view(B, 4)


def _fix(q):
    return q * torch.sign(q[..., :1])


def q_distance(pose1, pose2):
    """reference test_rot_sia.py:32-41: (2/pi) acos(min(|<fix(p1), fix(p2)>|, 1)); any leading batch shape"""
    d = (_fix(pose1) * _fix(pose2)).sum(-1).abs().clamp(max=1.0)
    return 2.0 * torch.acos(d) / math.pi


def relative_pose(pose1, pose2):
    """conj(pose1) (x) pose2, the relative rotation the reference evaluates (test_rot_sia.py:131-132)"""
    conj = pose1 * pose1.new_tensor([1.0, -1.0, -1.0, -1.0])
    w1, x1, y1, z1 = conj.unbind(-1)
    w2, x2, y2, z2 = pose2.unbind(-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
