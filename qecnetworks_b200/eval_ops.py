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


def q_distance(pose1, pose2):
    """reference test_rot_sia.py:32-41: (2/pi) acos(min(|<fix(p1), fix(p2)>|, 1)); any leading batch shape;
    one launch (qec_quat_eval)"""
    return QF.quat_eval(pose1.float(), pose2.float(), want_rel=False)[0]


def relative_pose(pose1, pose2):
    """conj(pose1) (x) pose2, the relative rotation the reference evaluates (test_rot_sia.py:131-132);
    one launch (qec_quat_eval)"""
    return QF.quat_eval(pose1.float(), pose2.float(), want_dist=False)[1]


def relative_rotation_error(pose_out, labels, rot_gt):
    """The rotation-error metric of the reference's Siamese test (test_rot_sia.py:94-123) on a batch, without the
    per-sample Python loop: pose_out [2,B,C,4] (QecSiaNet output), labels [B] (the class whose capsule is read:
    the prediction in the reference's test, the ground truth during training), rot_gt [B,4] = the rotation that
    took branch 0 to branch 1.  Returns d [B] = q_distance(pose_0, conj(rot_gt) (x) pose_1): 0 when the network's
    pose of branch 1 is exactly the rotated pose of branch 0."""
    S, B, C, _ = pose_out.shape
    idx = labels.view(1, B, 1, 1).expand(2, B, 1, 4).to(pose_out.device)
    sel = pose_out.float().gather(2, idx).squeeze(2)          # [2, B, 4]
    back = relative_pose(rot_gt.to(pose_out.device).float(), sel[1])  # conj(rot_gt) (x) pose_1
    return q_distance(sel[0].contiguous(), back)
