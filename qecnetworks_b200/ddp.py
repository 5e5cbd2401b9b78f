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
    only exchange is an all-reduce of ~3.7 MB (QecNet, 40 classes).

    `early`: parameters whose gradients are complete well before the end of the backward pass (pool2 of
    QEC-Net: 90 % of the bytes, done when 40 % of the backward is still to run).  They are laid out first
    in the buffer and their slice is all-reduced ASYNCHRONOUSLY as soon as the last of them has been
    accumulated (post-accumulate-grad hooks), so the transfer overlaps the rest of the backward instead
    of being exposed at the end of the step; all_reduce_mean() then reduces the small remainder and
    joins.  Every rank launches the same two collectives in the same order.  Protocol per step:
    zero() -> forward -> backward -> all_reduce_mean(); a backward that is NOT followed by
    all_reduce_mean() must run with `overlap = False`."""

    def __init__(self, params, early=None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        early_ids = {id(p) for p in (early or [])}
        self.early = [p for p in params if id(p) in early_ids]
        self.params = self.early + [p for p in params if id(p) not in early_ids]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n_early = sum(p.numel() for p in self.early)
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._seen = 0
        self._work = None
        self._group = None
        self.overlap = True  # set to False for a backward pass that must not touch the network (no all_reduce_mean after it)
        for p in self.early:
            p.register_post_accumulate_grad_hook(self._on_early_grad)

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1

    def _on_early_grad(self, _param):
        self._seen += 1
        if self._seen == len(self.early):
            self._seen = 0
            if self.overlap and self._distributed() and self._work is None:
                self._work = dist.all_reduce(self.flat[:self.n_early], op=dist.ReduceOp.SUM, group=self._group,
                                             async_op=True)

    def zero(self):
        if self._work is not None:  # a reduction nobody joined: finish it before the buffer is reused
            self._work.wait()
            self._work = None
        self.flat.zero_()
        self._seen = 0

    def all_reduce_mean(self, group=None):
        """sum over ranks / world size: the loss is a mean over the GLOBAL batch"""
        self._group = group
        if self._distributed():
            world = dist.get_world_size(group)
            if self._work is not None:  # the early slice is already in flight
                if self.n_early < self.flat.numel():
                    dist.all_reduce(self.flat[self.n_early:], op=dist.ReduceOp.SUM, group=group)
                self._work.wait()
                self._work = None
            else:
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
