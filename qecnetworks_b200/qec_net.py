"""QecNet / QecSiaNet -- the two callers of the hot path (reference models/qec_net.py,
models/qec_sia_net.py), kept as thin PyTorch glue over the B200 QecModule so that
the reference's training scripts can switch imports and nothing else:
same constructors, forward signatures, state_dict keys (pool1.*, pool2.*) and losses."""
import torch
import torch.nn.functional as F

from . import functional as QF
from .qec_module import QecModule


def _two_pools(net, points, lrfs, activation):
    # pool1: I = 1 (qec_net.py:33-37); pool2 over the 256 patch centres (:39-43)
    pose1, act1 = net.pool1(points, lrfs.unsqueeze(-2), activation.unsqueeze(-1))
    centres = points[:, :, 0, :].unsqueeze(1)
    pose2, act2 = net.pool2(centres, pose1.unsqueeze(1), F.relu(act1.unsqueeze(1)))
    return pose2.squeeze(), act2.squeeze()


def _spread(x, target, margin=0.1):
    """reference qec_net.py:53-59 (x [B,C]) / qec_sia_net.py:69-77 (x [2,B,C], the sum over the two branches):
    ONE launch for the loss and its gradient (qec_class_pose_loss) instead of ~25 ATen launches per branch.
    CUDA only, like the rest of the path (a CPU tensor raises in the wrapper)."""
    xs = x.unsqueeze(0) if x.dim() == 2 else x
    return QF.class_pose_loss(xs, target.reshape(-1), None, margin=margin)[0]


def _pose_diff(pose):
    """reference qec_net.py:61-64 / qec_sia_net.py:79-82 on already selected poses [2, B, 4]: one launch
    (qec_class_pose_loss with the pose term only) instead of ~10 ATen launches each way"""
    pose = pose.float().reshape(2, -1, 1, 4)  # the reference's mean runs over every leading dimension
    labels = torch.zeros(pose.shape[1], device=pose.device, dtype=torch.int64)
    return QF.class_pose_loss(None, labels, pose, w_pose=1.0)[0]


class _QecBase(torch.nn.Module):
    def __init__(self, num_points, inter_out_channels, num_of_class, num_iterations):
        super().__init__()
        self.num_points = num_points
        self.inter_out_channels = inter_out_channels
        self.num_iterations = num_iterations
        self.num_of_class = num_of_class
        self.num_pool1_patches = 256
        self.num_pool2_patches = 1
        self.num_neighbours2 = 9
        self.num_neighbours3 = 256
        self.pool1 = QecModule(1, inter_out_channels, num_iterations, self.num_neighbours2, self.num_pool1_patches)
        self.pool2 = QecModule(
            inter_out_channels, num_of_class, num_iterations, self.num_neighbours3, self.num_pool2_patches
        )

    def pose_diff_loss(self, pose):
        return _pose_diff(pose)

    def class_pose_loss(self, x, pose_out, target, pose_weight=0.1, margin=0.1):
        """The whole loss of train_cls_sia.py:70-79 in one launch: spread loss of the branch(es) + pose_weight x
        pose_diff_loss of the poses of the ground-truth class (the reference's per-sample selection loop
        `pose_out_act[:, i] = pose_out[:, i, target_[i, 0]]` included).  x [2,B,C] / pose_out [2,B,C,4] as returned
        by QecSiaNet.forward (or x [B,C] with pose_out None for QecNet).  -> (total, spread, pose_weight * pose_diff)
        like the reference's (classify_loss + pose_loss, classify_loss, pose_loss)."""
        t = target[:, 0] if target.dim() > 1 else target
        xs = x if x.dim() == 3 else x.unsqueeze(0)
        out = QF.class_pose_loss(xs, t.reshape(-1), pose_out, margin=margin, w_pose=pose_weight)
        return out[0], out[1].detach(), (pose_weight * out[2]).detach()


class QecNet(_QecBase):
    def forward(self, points4pool1, lrfs4pool1, activation4pool1, pointspool1index=None):
        return _two_pools(self, points4pool1, lrfs4pool1, activation4pool1)

    def spread_loss(self, x, target, epoch=0):
        return _spread(x, target.reshape(-1))


class QecSiaNet(_QecBase):
    def forward(self, points4pool1_, lrfs4pool1_, activation4pool1_, pointspool1index_=None):
        outs = [_two_pools(self, points4pool1_[:, i], lrfs4pool1_[:, i], activation4pool1_[:, i]) for i in range(2)]
        return torch.stack([o[0] for o in outs], 0), torch.stack([o[1] for o in outs], 0)

    def spread_loss(self, x, target, epoch=0):
        return _spread(x, target[:, 0])
