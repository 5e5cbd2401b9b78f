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
    joins.

    Collective order is the same on every rank by construction: with `early` parameters a step always
    issues exactly two all-reduces on `group`, the early slice first and the remainder second.  If the
    hooks of a rank did not all fire during its backward (an early parameter without a gradient in that
    step), all_reduce_mean() issues the early slice itself before the remainder -- late, but in the same
    order, so the ranks cannot mismatch.

    The gradients must stay views of the flat buffer: `optimizer.zero_grad()` of torch optimizers defaults
    to set_to_none=True, which drops the views (autograd would then allocate fresh `.grad` tensors and the
    all-reduce would see a stale buffer).  Use `bucket.zero()` (or `zero_grad(set_to_none=False)`);
    all_reduce_mean() checks the aliasing and raises otherwise.

    Protocol per step: zero() -> forward -> backward -> all_reduce_mean(); a backward that is NOT
    followed by all_reduce_mean() must run with `overlap = False`."""

    def __init__(self, params, early=None, group=None, direct=False):
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
        self._views = []
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self._views.append((p, off))
            off += p.numel()
        # direct: the backward kernels write parameter gradients straight into this buffer (no AccumulateGrad add per
        # parameter: 12 launches of the QEC-Net step).  Each parameter gets a "home"; release() drops the .grad views
        # so that the next backward hands the homes to the kernels and autograd adopts what they return as .grad.
        # Valid when every parameter enters the graph once per backward (a second use falls back to autograd's own
        # sum, which leaves the buffer: all_reduce_mean() / check_views() raise) and the buffer is zero when the
        # backward starts (FlatAdam leaves it so; otherwise zero()).
        self.direct = bool(direct)
        if self.direct:
            from .functional import GradHome

            for p, off in self._views:
                p._qec_grad_home = GradHome(self.flat, off)
            self.release()
        self._seen = 0
        self._work = None
        self._group = group  # fixed at construction: the hooks and all_reduce_mean() use the same one
        # set to False for a backward pass that must not touch the network (no all_reduce_mean after it);
        # QEC_DDP_OVERLAP=0 starts with it off (A/B measurements of the overlap)
        self.overlap = os.environ.get("QEC_DDP_OVERLAP", "1") != "0"
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

    def check_views(self):
        """every parameter's .grad is still the view of the flat buffer it was given"""
        base = self.flat.data_ptr()
        for p, off in self._views:
            g = p.grad
            if g is None or g.data_ptr() != base + 4 * off or g.numel() != p.numel():
                raise RuntimeError(
                    "FlatGradBucket: a parameter's .grad no longer aliases the flat gradient buffer (was "
                    "optimizer.zero_grad() called with set_to_none=True?): use bucket.zero() or "
                    "zero_grad(set_to_none=False), otherwise the all-reduce would exchange a stale buffer")

    def release(self):
        """direct mode: forget the .grad tensors (they alias the buffer, nothing is freed) and re-arm the homes"""
        for p, _ in self._views:
            p.grad = None
            home = getattr(p, "_qec_grad_home", None)
            if home is not None:
                home.taken = False

    def zero(self):
        if self._work is not None:  # a reduction nobody joined: finish it before the buffer is reused
            self._work.wait()
            self._work = None
        self.flat.zero_()
        self._seen = 0
        if self.direct:
            self.release()

    def all_reduce_mean(self, group=None, average=True):
        """sum over ranks / world size: the loss is a mean over the GLOBAL batch.  With average=False the
        division is left to the consumer (FlatAdam reads the gradient times 1/world); returns the factor
        that is still owed (1.0 if nothing is).  `group` is accepted for compatibility and must be the
        group given at construction."""
        if group is not None and group is not self._group:
            raise ValueError("FlatGradBucket: pass the process group at construction (the overlap hooks use it)")
        self.check_views()
        group = self._group
        owed = 1.0
        self._seen = 0
        if self._distributed():
            world = dist.get_world_size(group)
            if self.early and self.n_early < self.flat.numel():
                # always two collectives, in this order, on every rank
                if self._work is None:  # the hooks did not all fire on this rank (or overlap is off)
                    self._work = dist.all_reduce(self.flat[:self.n_early], op=dist.ReduceOp.SUM, group=group,
                                                 async_op=True)
                dist.all_reduce(self.flat[self.n_early:], op=dist.ReduceOp.SUM, group=group)
                self._work.wait()
                self._work = None
            else:
                if self._work is not None:  # every parameter is "early": the one collective is in flight
                    self._work.wait()
                    self._work = None
                else:
                    dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(world)
            else:
                owed = 1.0 / world
        return owed

    def nbytes(self):
        return self.flat.numel() * 4


class FlatAdam:
    """torch.optim.Adam (reference train_cls.py:44-51) on the bucket's flat layout: the parameters become views
    of one contiguous buffer in the same order as their gradients, and a step is ONE launch of
    qec_adam_flat over it (torch's multi-tensor fused Adam: 47 us of the 2.7 ms step for QEC-Net's 0.92 M
    parameters).  The launch also applies the 1/world factor of the gradient all-reduce and leaves the
    gradient buffer zeroed for the next step, so `bucket.zero()` is only needed before the first one.
    CUDA-graph capturable (the step count lives on the device).  CUDA only, like the rest of the path."""

    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, zero_grad=True):
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("betas must be in [0, 1)")
        self.bucket = bucket
        self.lr, self.betas, self.eps, self.weight_decay, self.zero_grad = float(lr), betas, float(eps), float(weight_decay), bool(zero_grad)
        g = bucket.flat
        self.flat = torch.empty_like(g)
        off = 0
        with torch.no_grad():
            for p in bucket.params:
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + n].view_as(p)
                off += n
        self.exp_avg = torch.zeros_like(g)
        self.exp_avg_sq = torch.zeros_like(g)
        self.state = torch.zeros(2, device=g.device, dtype=torch.int32)  # {steps taken, scratch of the kernel}

    def step(self, grad_scale=1.0):
        from . import _lib

        if not self.flat.is_cuda:
            raise _lib.QecLibraryError("FlatAdam.step: CUDA tensors only (no CPU fallback)")
        _lib.launch("qec_adam_flat", self.flat, self.bucket.flat, self.exp_avg, self.exp_avg_sq,
                    self.flat.numel(), self.lr, float(self.betas[0]), float(self.betas[1]), self.eps, self.weight_decay,
                    float(grad_scale), 1 if self.zero_grad else 0, self.state)
        if self.bucket.direct and self.zero_grad:
            self.bucket.release()  # the buffer is zero again: the next backward writes into it directly

    def steps_taken(self):
        return int(self.state[0].item())


def broadcast_parameters(module, src=0):
    """make every replica start from rank `src`'s parameters"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src)


def shard_batch(global_batch, rank, world):
    """contiguous batch slice of this rank (weak scaling uses a fixed per-rank batch instead)"""
    per = global_batch // world
    return rank * per, (rank + 1) * per
