"""Data-parallel plumbing: one process per GPU, replicas of the (0.9 M-float) parameter
set, and ONE flat-bucket all-reduce of the gradients per step (NCCL over NVLink on the
B200 box, gloo in the CPU tests).  Replaces the reference's nn.DataParallel
(train_cls.py:27: per-step replicate / scatter / gather / reduce to GPU 0).

The hot path shards over the batch with no data-path collective (every capsule
depends on one sample only); the gradient all-reduce is the only
exchange step."""
import datetime
import os

import torch
import torch.distributed as dist


# a mismatched collective must fail in minutes, not hang a GPU box for the default 10
_TIMEOUT = datetime.timedelta(seconds=int(os.environ.get("QEC_DIST_TIMEOUT_S", "120")))


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank).  Single process if WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local), timeout=_TIMEOUT)
        else:
            dist.init_process_group(backend, timeout=_TIMEOUT)
    return rank, world, local


class FlatGradBucket:
    """All parameter gradients as views into one contiguous FP32 buffer, so the step's
    only collective is a single all-reduce of ~3.7 MB (QecNet, 40 classes)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        """sum over ranks / world size: the loss is a mean over the GLOBAL batch"""
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(world)

    def nbytes(self):
        return self.flat.numel() * 4


def broadcast_parameters(module, src=0):
    """make every replica start from rank `src`'s parameters"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src)


def shard_batch(global_batch, rank, world):
    """contiguous batch slice of this rank (weak scaling uses a fixed per-rank batch instead)"""
    per = global_batch // world
    return rank * per, (rank + 1) * per
