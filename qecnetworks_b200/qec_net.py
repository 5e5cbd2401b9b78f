"""QecNet / QecSiaNet -- the two callers of the hot path (reference models/qec_net.py,
models/qec_sia_net.py), kept as thin PyTorch glue over the B200 QecModule so that
the reference's training scripts can switch imports and nothing else:
same constructors, forward signatures, state_dict keys (pool1.*, pool2.*) and losses."""
import math

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
    """reference qec_net.py:53-59; one launch pair for the loss and its gradient (qec_spread_loss) instead
    of ~25 ATen launches.  CUDA only, like the rest of the path (a CPU tensor raises in the wrapper)."""
    return QF.spread_loss(x, target.view(-1), margin)


def _pose_diff(pose):
    dot = (pose[0] * pose[1]).sum(dim=-1)
    return (2.0 * torch.acos(torch.clamp(dot.abs(), max=0.9999)) / math.pi).mean()


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
        t = target[:, 0].reshape(-1)
        return _spread(x[0], t) + _spread(x[1], t)
