"""torch.autograd.Function wrappers over the C-ABI kernels (host-side plumbing only:
PyTorch supplies device memory, the current stream and the autograd graph).

Tensors follow the reference layouts (models/qec_module.py):
  lrfs [B,P,K,I,4], activation [B,P,K,I], points [B,P,K,3],
  t_raw [B*P*K, 4*I*O] (feature order q*I*O + i*O + o), votes [B,P,O,K*I,4].
"""
import torch

from . import _lib
from ._lib import launch


def _check(t, name, shape=None):
    if not t.is_cuda:
        raise _lib.QecLibraryError(
            "%s must be a CUDA tensor: the QEC routing kernels have no CPU implementation" % name
        )
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t.contiguous()


class GradHome:
    """Where a parameter's gradient lives inside a flat gradient buffer (ddp.FlatGradBucket(direct=True)).
    `taken`: the home was handed to a backward node since the last release (a parameter used twice in one graph
    gets it once; the second contribution is a fresh tensor and autograd adds the two)."""

    __slots__ = ("flat", "offset", "taken")

    def __init__(self, flat, offset):
        self.flat, self.offset, self.taken = flat, int(offset), False


def _home_out(param, numel=None):
    """A fresh tensor aliasing `param`'s home in the flat gradient buffer, or None.  Given out only while the
    parameter has no .grad (autograd then adopts the returned tensor as .grad without a copy or an add) and once per
    release.  The home is all zeros at that point: the buffer is zeroed between steps (FlatAdam / bucket.zero())."""
    home = getattr(param, "_qec_grad_home", None)
    if home is None or home.taken or param.grad is not None or home.flat.device != param.device:
        return None
    n = param.numel() if numel is None else numel
    if home.offset + n > home.flat.numel():
        return None
    home.taken = True
    out = torch.empty(0, device=home.flat.device, dtype=torch.float32)
    shape = tuple(param.shape) if numel is None else (n,)
    out.set_(home.flat.untyped_storage(), home.flat.storage_offset() + home.offset, shape)
    return out


def _home_pair_out(a, b):
    """alpha, beta -> one 2-float tensor over their (adjacent) homes, or None"""
    ha, hb = getattr(a, "_qec_grad_home", None), getattr(b, "_qec_grad_home", None)
    if ha is None or hb is None or ha.flat is not hb.flat or hb.offset != ha.offset + 1 or hb.taken or b.grad is not None:
        return None
    out = _home_out(a, 2)
    if out is not None:
        hb.taken = True
    return out


def _zeros_split(dev, *shapes):
    """several zero tensors carved from ONE zero-filled buffer (one fill launch instead of one per tensor); a shape
    of None yields None.  Offsets are rounded up to 4 floats so that every piece stays 16-byte aligned."""
    sizes = [0 if sh is None else int(torch.Size(sh).numel()) for sh in shapes]
    padded = [(n + 3) // 4 * 4 for n in sizes]
    buf = torch.zeros(sum(padded), device=dev, dtype=torch.float32)
    out, off = [], 0
    for sh, n, pn in zip(shapes, sizes, padded):
        out.append(None if sh is None else buf[off:off + n].view(sh))
        off += pn
    return out


class LrfMeanRotate(torch.autograd.Function):
    """(lrfs, activation, points) -> (mean_lrf [B,P,I,4], xrot [B,P,K,I,3]).
    reference: QecModule.mean_lrf_per_patch + the qrotv at the top of get_votes
    (qec_module.py:101-120)."""

    @staticmethod
    def forward(ctx, lrfs, activation, points):
        B, P, K, I, _ = lrfs.shape
        lrfs = _check(lrfs, "lrfs", (B, P, K, I, 4))
        activation = _check(activation, "activation", (B, P, K, I))
        points = _check(points, "points", (B, P, K, 3))
        mean = torch.empty(B, P, I, 4, device=lrfs.device, dtype=torch.float32)
        xrot = torch.empty(B, P, K, I, 3, device=lrfs.device, dtype=torch.float32)
        launch("qec_lrf_mean_fwd", lrfs, activation, points, mean, xrot, B * P, K, I)
        ctx.save_for_backward(lrfs, points, mean)
        return mean, xrot

    @staticmethod
    def backward(ctx, g_mean, g_xrot):
        lrfs, points, mean = ctx.saved_tensors
        B, P, K, I, _ = lrfs.shape
        need_lrfs, _, need_points = ctx.needs_input_grad
        if g_xrot is None:
            g_xrot = torch.zeros(B, P, K, I, 3, device=lrfs.device, dtype=torch.float32)
        g_xrot = g_xrot.contiguous()
        g_mean = None if g_mean is None else g_mean.contiguous()
        g_lrfs = torch.empty_like(lrfs)
        g_points = torch.zeros_like(points) if need_points else None
        launch("qec_lrf_mean_bwd", lrfs, points, mean, g_xrot, g_mean, g_lrfs,
             g_points, B * P, K, I)
        return (g_lrfs if need_lrfs else None), None, g_points


class Votes(torch.autograd.Function):
    """(lrfs [B,P,K,I,4], t_raw [B*P*K, 4*I*O]) -> votes [B,P,O,K*I,4].
    reference: QecModule.get_votes after the Linears (qec_module.py:126-147)."""

    @staticmethod
    def forward(ctx, lrfs, t_raw):
        B, P, K, I, _ = lrfs.shape
        O = t_raw.shape[-1] // (4 * I)
        lrfs = _check(lrfs, "lrfs", (B, P, K, I, 4))
        t_raw = _check(t_raw, "t_raw", (B * P * K, 4 * I * O))
        votes = torch.empty(B, P, O, K * I, 4, device=lrfs.device, dtype=torch.float32)
        launch("qec_votes_fwd", lrfs, t_raw, votes, B * P, K, I, O)
        ctx.save_for_backward(lrfs, t_raw)
        return votes

    @staticmethod
    def backward(ctx, g_votes):
        lrfs, t_raw = ctx.saved_tensors
        B, P, K, I, _ = lrfs.shape
        O = t_raw.shape[-1] // (4 * I)
        need_lrfs = ctx.needs_input_grad[0]
        g_votes = g_votes.contiguous()
        g_traw = torch.empty_like(t_raw)
        g_lrfs = torch.zeros_like(lrfs) if need_lrfs else None
        launch("qec_votes_bwd", lrfs, t_raw, g_votes, g_traw, g_lrfs, B * P, K, I, O)
        return g_lrfs, g_traw


class Routing(torch.autograd.Function):
    """(votes [B,P,O,J,4], activation [B,P,J], alpha [1], beta [1], T) -> (pose [B,P,O,4],
    agreement [B,P,O]).  reference: the loop and epilogue of QecModule.forward
    (qec_module.py:172-194)."""

    @staticmethod
    def forward(ctx, votes, activation, alpha, beta, num_iterations):
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zero tensor autograd has to fill
        B, P, O, J, _ = votes.shape
        votes = _check(votes, "votes", (B, P, O, J, 4))
        activation = _check(activation, "activation", (B, P, J))
        alpha = _check(alpha, "alpha", (1,))
        beta = _check(beta, "beta", (1,))
        T = int(num_iterations)
        dev = votes.device
        pose = torch.empty(B, P, O, 4, device=dev, dtype=torch.float32)
        agreement = torch.empty(B, P, O, device=dev, dtype=torch.float32)
        poses_all = torch.empty(T, B, P, O, 4, device=dev, dtype=torch.float32)
        stats = torch.empty(B, P, O, 2, device=dev, dtype=torch.float32)
        launch("qec_routing_fwd", votes, activation, alpha, beta, pose, agreement,
             poses_all, stats, B * P, O, J, T)
        ctx.save_for_backward(votes, activation, alpha, beta, poses_all, stats)
        ctx.T = T
        return pose, agreement

    @staticmethod
    def backward(ctx, g_pose, g_agreement):
        votes, activation, alpha, beta, poses_all, stats = ctx.saved_tensors
        B, P, O, J, _ = votes.shape
        dev = votes.device
        g_pose = torch.zeros(B, P, O, 4, device=dev) if g_pose is None else g_pose.contiguous()
        g_agreement = torch.zeros(B, P, O, device=dev) if g_agreement is None else g_agreement.contiguous()
        need_act = ctx.needs_input_grad[1]
        g_votes = torch.empty_like(votes)
        g_act = torch.zeros_like(activation) if need_act else None
        g_ab = torch.zeros(2, device=dev, dtype=torch.float32)
        launch("qec_routing_bwd", votes, activation, alpha, beta, poses_all, stats,
             g_pose, g_agreement, g_votes, g_act, g_ab, B * P, O, J, ctx.T)
        return g_votes, g_act, g_ab[0:1], g_ab[1:2], None


def set_fused_wide(mode):
    """0 | 1 | 2 | 3, see qec_routing_fused_set_wide in include/qec_b200.h; returns the previous mode"""
    return _lib.query("qec_routing_fused_set_wide", int(mode))


def set_fused_pp(mode):
    """bit 0: forward, bit 1: backward on the third-generation cluster kernels (qec_routing_fused_set_pp); returns
    the previous mode"""
    return _lib.query("qec_routing_fused_set_pp", int(mode))


def routing_fused_forward_raw(t_oqi, lrfs, activation, alpha, beta, T):
    """no autograd: -> pose, agreement, poses_all [T,B,P,O,4], stats [B,P,O,2] (the last two are what
    the backward needs besides the inputs)"""
    B, P, K, I, _ = lrfs.shape
    O = t_oqi.shape[-1] // (4 * I)
    lrfs = _check(lrfs, "lrfs", (B, P, K, I, 4))
    t_oqi = _check(t_oqi, "t_oqi", (B * P * K, 4 * I * O))
    activation = _check(activation, "activation", (B, P, K * I))
    alpha = _check(alpha, "alpha", (1,))
    beta = _check(beta, "beta", (1,))
    dev = lrfs.device
    pose = torch.empty(B, P, O, 4, device=dev, dtype=torch.float32)
    agreement = torch.empty(B, P, O, device=dev, dtype=torch.float32)
    poses_all = torch.empty(T, B, P, O, 4, device=dev, dtype=torch.float32)
    stats = torch.empty(B, P, O, 2, device=dev, dtype=torch.float32)
    launch("qec_routing_fused_fwd", t_oqi, lrfs, activation, alpha, beta, pose,
         agreement, poses_all, stats, B * P, K, I, O, T)
    return pose, agreement, poses_all, stats


def routing_fused_backward_raw(t_oqi, lrfs, activation, alpha, beta, poses_all, stats, g_pose, g_agreement, T,
                               need_lrfs, need_act, split=False, g_ab=None):
    """no autograd: -> g_toqi, g_lrfs | None, g_act | None, g_alpha_beta [2].  With split=True g_toqi is a
    pair (hi, lo) of [B*P*K, 4*I*O] tensors: hi (float32) = the gradient rounded to the TF32 grid, lo
    (bfloat16) = the remainder (qec_routing_fused_bwd_split; packed-pair geometries only)."""
    B, P, K, I, _ = lrfs.shape
    O = t_oqi.shape[-1] // (4 * I)
    dev = lrfs.device
    # everything the kernel accumulates into, and the upstream gradients autograd left out, from one zero fill
    # (g_ab may be given by the caller: a zeroed home in the flat gradient buffer)
    z_pose, z_agr, g_lrfs, g_act, z_ab = _zeros_split(
        dev, (B, P, O, 4) if g_pose is None else None, (B, P, O) if g_agreement is None else None,
        lrfs.shape if need_lrfs else None, activation.shape if need_act else None, (2,) if g_ab is None else None)
    g_pose = z_pose if g_pose is None else g_pose.contiguous()
    g_agreement = z_agr if g_agreement is None else g_agreement.contiguous()
    g_ab = z_ab if g_ab is None else g_ab
    if split:
        g_toqi = (torch.empty_like(t_oqi), torch.empty(t_oqi.shape, device=dev, dtype=torch.bfloat16))
        launch("qec_routing_fused_bwd_split", t_oqi, lrfs, activation, alpha, beta,
             poses_all, stats, g_pose, g_agreement, g_toqi[0], g_toqi[1], g_lrfs,
             g_act, g_ab, B * P, K, I, O, T)
    else:
        g_toqi = torch.empty_like(t_oqi)
        launch("qec_routing_fused_bwd", t_oqi, lrfs, activation, alpha, beta, poses_all,
             stats, g_pose, g_agreement, g_toqi, g_lrfs, g_act, g_ab, B * P, K, I,
             O, T)
    return g_toqi, g_lrfs, g_act, g_ab


class RoutingFused(torch.autograd.Function):
    """(t_oqi [B*P*K, 4*I*O] capsule-major, lrfs [B,P,K,I,4], activation [B,P,K*I], alpha, beta, T)
    -> (pose [B,P,O,4], agreement [B,P,O]): vote generation + routing with the votes resident in
    shared memory across a thread-block cluster (qec_module.py:126-147 + 172-194 in one kernel)."""

    @staticmethod
    def forward(ctx, t_oqi, lrfs, activation, alpha, beta, num_iterations):
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zero tensor autograd has to fill
        T = int(num_iterations)
        pose, agreement, poses_all, stats = routing_fused_forward_raw(t_oqi, lrfs, activation, alpha, beta, T)
        ctx.save_for_backward(t_oqi, lrfs, activation, alpha, beta, poses_all, stats)
        ctx.T = T
        return pose, agreement

    @staticmethod
    def backward(ctx, g_pose, g_agreement):
        t_oqi, lrfs, activation, alpha, beta, poses_all, stats = ctx.saved_tensors
        g_toqi, g_lrfs, g_act, g_ab = routing_fused_backward_raw(
            t_oqi, lrfs, activation, alpha, beta, poses_all, stats, g_pose, g_agreement, ctx.T,
            ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return g_toqi, g_lrfs, g_act, g_ab[0:1], g_ab[1:2], None


def split_operand(src, extra=None, extra_const=1.0, perm=None, add_col=True, blocks=(1, 1, 2), stride=None, width=None,
                  bf16_width=None):
    """[src | extra column] -> (out_f32 | None, out_bf16 | None): the 3xTF32 GEMM operand built in one launch
    (qec_split_operand): `blocks` (1 = TF32 hi part, 2 = lo part) at block stride `stride`, zero-padded to
    `width`; and/or the hi part in BF16 padded to `bf16_width`.  Row r comes from src[perm[r]]."""
    src = _check(src, "src", tuple(src.shape))
    R, C = src.shape
    Kc = C + (1 if add_col else 0)
    dev = src.device
    if extra is not None:
        extra = _check(extra, "extra", (R,))
    if perm is not None:
        if perm.dtype != torch.int64 or not perm.is_cuda:
            raise TypeError("perm must be a CUDA int64 tensor")
        perm = perm.contiguous()
    nblk = len(blocks) if width is not None else 0
    stride = Kc if stride is None else int(stride)
    out = torch.empty(R, int(width), device=dev, dtype=torch.float32) if width is not None else None
    outb = torch.empty(R, int(bf16_width), device=dev, dtype=torch.bfloat16) if bf16_width is not None else None
    blk = list(blocks) + [0, 0, 0]
    launch("qec_split_operand", src, extra, float(extra_const), perm, out, outb, R, C,
         1 if add_col else 0, stride, int(width or 0), int(bf16_width or 0), blk[0], blk[1], blk[2], nblk)
    return out, outb


def fold_chunks(a, b, perm, out_main, out_last, K):
    """out_main[perm[r], c] = sum_s a[s,r,c] + a[s,r,KP+c] + b[s,r,c] (c < out_main width); column K-1 goes to
    out_last when out_main is K-1 wide (qec_fold_chunks)"""
    S, R, KP2 = a.shape
    KP = KP2 // 2
    launch("qec_fold_chunks", a, b, perm, out_main, out_last, S, R, KP, K, out_main.shape[1])


class SpreadLoss(torch.autograd.Function):
    """spread loss of QEC-Net (reference models/qec_net.py:53-59) with its gradient in one launch pair"""

    @staticmethod
    def forward(ctx, x, target, margin):
        B, C = x.shape
        x = _check(x, "x", (B, C))
        target = target.to(device=x.device, dtype=torch.int64).contiguous().view(B)
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        g = torch.empty_like(x)
        scratch = torch.empty(B, device=x.device, dtype=torch.float32)
        err = torch.zeros(1, device=x.device, dtype=torch.int32)
        launch("qec_spread_loss", x, target, float(margin), loss, g, scratch, err, B, C)
        ctx.save_for_backward(g)
        ctx.err = err  # checked lazily by callers that can afford a sync (spread_loss(..., check=True))
        return loss.view(())

    @staticmethod
    def backward(ctx, g_loss):
        (g,) = ctx.saved_tensors
        return g * g_loss, None, None


def spread_loss(x, target, margin=0.1):
    return SpreadLoss.apply(x, target, margin)


def gen2_bwd_supported(M, N, Kc):
    return bool(_lib.query("qec_gen2_bwd_supported", int(M), int(N), int(Kc)))


def gen2_fwd_supported(M, N, Kc):
    return bool(_lib.query("qec_gen2_fwd_supported", int(M), int(N), int(Kc)))


def gen2_forward(h, W2, b2, perm=None):
    """t [M, N] = h W2[perm]^T + b2[perm]: quater_gen2 (reference qec_module.py:124) on the tcgen05 tensor cores at
    FP32 accuracy (qec_gen2_fwd), output columns in the order `perm` gives (capsule-major for the cluster kernels)"""
    M, Kc = h.shape
    N = W2.shape[0]
    h = _check(h, "h", (M, Kc))
    W2 = _check(W2, "W2", (N, Kc))
    b2 = _check(b2, "b2", (N,))
    if perm is not None and (perm.dtype != torch.int64 or not perm.is_cuda):
        raise TypeError("perm must be a CUDA int64 tensor")
    ws = torch.empty(int(_lib.query("qec_gen2_fwd_workspace", M, N, Kc)), device=h.device, dtype=torch.float32)
    t = torch.empty(M, N, device=h.device, dtype=torch.float32)
    launch("qec_gen2_fwd", h, W2, b2, None if perm is None else perm.contiguous(), t, ws, M, N, Kc)
    return t


def gen2_backward(g, h, W2, perm, need_h=True, need_w=True, out_W2=None, out_b2=None):
    """Backward of quater_gen2 (reference qec_module.py:124) from the FP32 transform gradient g [M, N] on the tcgen05
    tensor cores (qec_gen2_bwd: TMA-fed, 3-term TF32 split, accumulator in tensor memory):
    -> (g_h [M, Kc] | None, g_W2 [N, Kc] | None, g_b2 [N] | None)"""
    M, N = g.shape
    Kc = h.shape[1]
    g = _check(g, "g", (M, N))
    h = _check(h, "h", (M, Kc))
    W2 = _check(W2, "W2", (N, Kc))
    if perm.dtype != torch.int64 or not perm.is_cuda:
        raise TypeError("perm must be a CUDA int64 tensor")
    dev = g.device
    ws = torch.empty(int(_lib.query("qec_gen2_bwd_workspace", M, N, Kc)), device=dev, dtype=torch.float32)
    g_h = torch.empty(M, Kc, device=dev, dtype=torch.float32) if need_h else None
    g_W2 = (out_W2 if out_W2 is not None else torch.empty(N, Kc, device=dev, dtype=torch.float32)) if need_w else None
    g_b2 = (out_b2 if out_b2 is not None else torch.empty(N, device=dev, dtype=torch.float32)) if need_w else None
    launch("qec_gen2_bwd", g, h, W2, perm.contiguous(), g_h, g_W2, g_b2, ws, M, N, Kc, 0)
    return g_h, g_W2, g_b2


def linear_supported(C, O):
    return bool(_lib.query("qec_linear_supported", int(C), int(O)))


class LinearSmall(torch.autograd.Function):
    """y = x W^T + b for a narrow layer (O <= 64 out features): quater_gen1 of the pool2 regime (reference
    qec_module.py:123) on the package's own FP32 kernels -- one launch forward, two backward (qec_linear_fwd / _bwd)
    instead of the library's badly shaped SGEMMs + split-K reduction + column sum."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        M, C = x.shape
        O = weight.shape[0]
        x = _check(x, "x", (M, C))
        weight = _check(weight, "weight", (O, C))
        bias = _check(bias, "bias", (O,))
        y = torch.empty(M, O, device=x.device, dtype=torch.float32)
        launch("qec_linear_fwd", x, weight, bias, y, M, C, O)
        ctx.save_for_backward(x, weight)
        ctx.bias = bias  # (only to find its home in a flat gradient buffer)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        M, C = x.shape
        O = weight.shape[0]
        g = _check(g, "g", (M, O))
        need_x, need_w, need_b = ctx.needs_input_grad
        dev = x.device
        ws = torch.empty(int(_lib.query("qec_linear_bwd_workspace", M, C, O)), device=dev, dtype=torch.float32)
        g_x = torch.empty_like(x) if need_x else None
        g_w = (_home_out(weight) if need_w else None)
        if need_w and g_w is None:
            g_w = torch.empty_like(weight)
        g_b = (_home_out(ctx.bias) if need_b else None)
        if need_b and g_b is None:
            g_b = torch.empty(O, device=dev, dtype=torch.float32)
        launch("qec_linear_bwd", x, weight, g, g_x, g_w, g_b, ws, M, C, O)
        return g_x, g_w, g_b


def compose_supported(C):
    return bool(_lib.query("qec_compose_supported", int(C)))


class ComposeLinears(torch.autograd.Function):
    """(W1 [O,C], b1 [O], W2 [R,O], b2 [R]) -> (Wc = W2 W1 [R,C], bc = W2 b1 + b2 [R]): the two Linears of get_votes
    composed (reference qec_module.py:123-124: no non-linearity between them), one launch each way."""

    @staticmethod
    def forward(ctx, W1, b1, W2, b2):
        O, C = W1.shape
        R = W2.shape[0]
        W1 = _check(W1, "W1", (O, C))
        b1 = _check(b1, "b1", (O,))
        W2 = _check(W2, "W2", (R, O))
        b2 = _check(b2, "b2", (R,))
        Wc = torch.empty(R, C, device=W1.device, dtype=torch.float32)
        bc = torch.empty(R, device=W1.device, dtype=torch.float32)
        launch("qec_compose_fwd", W1, b1, W2, b2, Wc, bc, R, O, C)
        ctx.save_for_backward(W1, b1, W2)
        ctx.b2 = b2  # (only to find its home in a flat gradient buffer)
        return Wc, bc

    @staticmethod
    def backward(ctx, g_Wc, g_bc):
        W1, b1, W2 = ctx.saved_tensors
        O, C = W1.shape
        R = W2.shape[0]
        g_Wc = _check(g_Wc, "g_Wc", (R, C))
        g_bc = _check(g_bc, "g_bc", (R,))
        g_W1, g_b1, g_W2 = (_home_out(t) for t in (W1, b1, W2))
        g_W1 = torch.empty_like(W1) if g_W1 is None else g_W1
        g_b1 = torch.empty_like(b1) if g_b1 is None else g_b1
        g_W2 = torch.empty_like(W2) if g_W2 is None else g_W2
        g_b2 = _home_out(ctx.b2)  # dL/db2 = dL/dbc: copied into b2's home when it has one, else passed through
        launch("qec_compose_bwd", W1, b1, W2, g_Wc, g_bc, g_W1, g_b1, g_W2, g_b2, R, O, C)
        return g_W1, g_b1, g_W2, (g_bc if g_b2 is None else g_b2)


class ClassPoseLoss(torch.autograd.Function):
    """(x [S,B,C] | None, labels [B], pose [2,B,C,4] | None) -> loss [3] = {total, spread, pose_diff}: the spread
    loss of S branches (reference qec_net.py:53-59 / qec_sia_net.py:69-77) + w_pose x the pose-difference loss on
    the ground-truth class poses (qec_sia_net.py:79-82 after the selection loop of train_cls_sia.py:72-78), values
    and gradients in one launch (qec_class_pose_loss).  Only loss[0] (the total) carries gradient."""

    @staticmethod
    def forward(ctx, x, pose, labels, margin, w_pose):
        ref = x if x is not None else pose
        dev = ref.device
        if x is not None:
            S, B, C = x.shape
            x = _check(x, "x", (S, B, C))
        else:
            S = 0
            _, B, C, _ = pose.shape
        if pose is not None:
            pose = _check(pose, "pose", (2, B, C, 4))
        labels = labels.to(device=dev, dtype=torch.int64).contiguous().view(B)
        loss = torch.empty(3, device=dev, dtype=torch.float32)
        g_x = torch.empty_like(x) if x is not None else None
        g_pose = torch.empty_like(pose) if pose is not None else None
        err = torch.zeros(1, device=dev, dtype=torch.int32)
        launch("qec_class_pose_loss", x, labels, float(margin), pose, float(w_pose), loss, g_x, g_pose, err, S, B, C)
        ctx.save_for_backward(g_x, g_pose)
        ctx.err = err
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        g_x, g_pose = ctx.saved_tensors
        g = g_loss[0]
        return (None if g_x is None else g_x * g), (None if g_pose is None else g_pose * g), None, None, None


def class_pose_loss(x, labels, pose=None, margin=0.1, w_pose=0.1):
    """-> loss [3] = {total, spread, w-less pose_diff}; differentiate loss[0]"""
    return ClassPoseLoss.apply(x, pose, labels, margin, w_pose)


def quat_eval(pose1, pose2, want_dist=True, want_rel=True):
    """batched q_distance / relative pose of the reference's rotation test (test_rot_sia.py:32-41, 131-132):
    pose1, pose2 [..., 4] -> (dist [...] | None, rel [..., 4] | None)   (qec_quat_eval)"""
    shape = pose1.shape[:-1]
    a = _check(pose1.reshape(-1, 4), "pose1")
    b = _check(pose2.reshape(-1, 4), "pose2", tuple(a.shape))
    n = a.shape[0]
    dist = torch.empty(n, device=a.device, dtype=torch.float32) if want_dist else None
    rel = torch.empty(n, 4, device=a.device, dtype=torch.float32) if want_rel else None
    if n:
        launch("qec_quat_eval", a, b, dist, rel, n)
    return (None if dist is None else dist.view(shape)), (None if rel is None else rel.view(*shape, 4))


def routing_fused_packed_supported(K, I, O, T):
    return bool(_lib.query("qec_routing_fused_packed_supported", int(K), int(I), int(O), int(T)))


class RoutingSmall(torch.autograd.Function):
    """in_channels = 1: (xrot [B,P,K,1,3], lrfs [B,P,K,1,4], activation [B,P,K], Wc [4*O,3], bc [4*O],
    alpha, beta, T) -> (pose [B,P,O,4], agreement [B,P,O]) with the composed transform Wc x + bc
    evaluated in-kernel (qec_module.py:123-147 + 172-194).  Differentiable w.r.t. Wc, bc, alpha,
    beta only; the caller falls back to votes() + routing() when the inputs need gradients."""

    @staticmethod
    def forward(ctx, xrot, lrfs, activation, Wc, bc, alpha, beta, num_iterations, keep):
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zero tensor autograd has to fill
        # keep: a backward may follow (decided by the caller: inside forward() the grad mode is always off, and
        # needs_input_grad reflects requires_grad of the inputs even under no_grad)
        saved = bool(keep)
        if saved and routing_small_supported(activation.shape[2], 1, bc.shape[0] // 4, num_iterations) != 1:
            raise _lib.QecLibraryError("routing_small: K > 16 is forward-only (inference); use votes() + routing()")
        B, P, K = activation.shape
        O = bc.shape[0] // 4
        xrot = _check(xrot, "xrot", (B, P, K, 1, 3))
        lrfs = _check(lrfs, "lrfs", (B, P, K, 1, 4))
        activation = _check(activation, "activation", (B, P, K))
        Wc = _check(Wc, "Wc", (4 * O, 3))
        bc = _check(bc, "bc", (4 * O,))
        alpha = _check(alpha, "alpha", (1,))
        beta = _check(beta, "beta", (1,))
        T = int(num_iterations)
        dev = lrfs.device
        pose = torch.empty(B, P, O, 4, device=dev, dtype=torch.float32)
        agreement = torch.empty(B, P, O, device=dev, dtype=torch.float32)
        # inference: nothing is kept for a backward (T x 16 B per capsule)
        poses_all = torch.empty(T, B, P, O, 4, device=dev, dtype=torch.float32) if saved else None
        stats = torch.empty(B, P, O, 2, device=dev, dtype=torch.float32) if saved else None
        launch("qec_routing_small_fwd", xrot, lrfs, activation, Wc, bc, alpha, beta,
             pose, agreement, poses_all, stats, B * P, K, O, T)
        if saved:
            ctx.save_for_backward(xrot, lrfs, activation, Wc, bc, alpha, beta, poses_all, stats)
        ctx.T = T
        return pose, agreement

    @staticmethod
    def backward(ctx, g_pose, g_agreement):
        xrot, lrfs, activation, Wc, bc, alpha, beta, poses_all, stats = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            raise _lib.QecLibraryError("routing_small has no gradient w.r.t. its data inputs; use votes() + routing()")
        B, P, K = activation.shape
        O = bc.shape[0] // 4
        dev = lrfs.device
        g_ab = _home_pair_out(alpha, beta)  # straight into the flat gradient buffer when the parameters live there
        z_pose, z_agr, g_Wc, g_bc, z_ab = _zeros_split(
            dev, (B, P, O, 4) if g_pose is None else None, (B, P, O) if g_agreement is None else None, Wc.shape, bc.shape,
            (2,) if g_ab is None else None)
        g_pose = z_pose if g_pose is None else g_pose.contiguous()
        g_agreement = z_agr if g_agreement is None else g_agreement.contiguous()
        g_ab = z_ab if g_ab is None else g_ab
        launch("qec_routing_small_bwd", xrot, lrfs, activation, Wc, bc, alpha, beta,
             poses_all, stats, g_pose, g_agreement, g_Wc, g_bc, g_ab, B * P, K, O,
             ctx.T)
        return None, None, None, g_Wc, g_bc, g_ab[0:1], g_ab[1:2], None, None


def routing_small_supported(K, I, O, T):
    """0 = not served, 1 = forward + backward (K <= 16), 2 = forward only (16 < K <= 64)"""
    return int(_lib.query("qec_routing_small_supported", int(K), int(I), int(O), int(T)))


def routing_small(xrot, lrfs, activation, Wc, bc, alpha, beta, num_iterations):
    keep = torch.is_grad_enabled() and any(t.requires_grad for t in (Wc, bc, alpha, beta))
    return RoutingSmall.apply(xrot, lrfs, activation, Wc, bc, alpha, beta, num_iterations, keep)


def routing_fused_supported(K, I, O, T):
    return bool(_lib.query("qec_routing_fused_supported", int(K), int(I), int(O), int(T)))


def routing_fused(t_oqi, lrfs, activation, alpha, beta, num_iterations):
    return RoutingFused.apply(t_oqi, lrfs, activation, alpha, beta, num_iterations)


def lrf_mean_rotate(lrfs, activation, points):
    return LrfMeanRotate.apply(lrfs, activation, points)


def votes(lrfs, t_raw):
    return Votes.apply(lrfs, t_raw)


def routing(votes_, activation, alpha, beta, num_iterations):
    return Routing.apply(votes_, activation, alpha, beta, num_iterations)


def knn_indices(points, K, centre_idx=None, centres=None):
    """K nearest neighbours of P centres among the N points of each sample, on the device (qec_knn):
    points [B,N,3] float32; centres as point indices centre_idx [B,P] int64 or coordinates centres [B,P,3].
    -> idx [B,P,K] int64 sorted by (squared distance, point index), bit-exact against
    oracle.qec_oracle.knn_indices (-1 where N < K); err [1] int32 (1 = a centre index was out of range).

    The ids are plain point numbers 0..N-1.  neighbour_gather() follows the reference LOADER's id contract
    instead (id <= 0 = inactive slot, the last row of the sets reads as zeros), so do not feed these indices
    to it unchanged: knn_neighbourhoods() composes the two correctly."""
    B, N, _ = points.shape
    points = _check(points, "points", (B, N, 3))
    if (centre_idx is None) == (centres is None):
        raise ValueError("give exactly one of centre_idx / centres")
    if centre_idx is not None:
        if not centre_idx.is_cuda or centre_idx.dtype != torch.int64:
            raise TypeError("centre_idx must be a CUDA int64 tensor")
        centre_idx = centre_idx.contiguous()
        P = centre_idx.shape[1]
    else:
        P = centres.shape[1]
        centres = _check(centres, "centres", (B, P, 3))
    idx = torch.empty(B, P, int(K), device=points.device, dtype=torch.int64)
    err = torch.zeros(1, device=points.device, dtype=torch.int32)
    launch("qec_knn", points, centres, centre_idx, idx, err, B, N, P, int(K))
    return idx, err


def knn_neighbourhoods(points, lrfs, K, centre_idx):
    """kNN search + loader gather on the device: the K-neighbourhoods of the P centres as the routing layer
    consumes them.  points [B,N,3], lrfs [B,N,4], centre_idx [B,P] -> (points [B,P,K,3], lrfs [B,P,K,4],
    act [B,P,K], idx [B,P,K]).

    qec_knn numbers points 0..N-1; qec_gather mirrors the reference loader, where id <= 0 marks an inactive slot
    and the last row of the sets is zeroed (modelnet_with_lrf_sample_index_loader.py:139-160).  The
    composition therefore shifts the ids by one and pads the sets with a leading dummy row and a trailing
    zero row, so that point 0 stays active, point N-1 keeps its coordinates and a missing neighbour (-1, only
    when N < K) reads zeros with activation 0."""
    B, N, _ = points.shape
    idx, err = knn_indices(points, K, centre_idx=centre_idx)
    pad3 = points.new_zeros(B, 1, 3)
    pad4 = lrfs.new_zeros(B, 1, 4)
    ids = torch.where(idx >= 0, idx + 1, idx)
    gp, gl, ga, gerr = neighbour_gather(torch.cat([pad3, points, pad3], 1), torch.cat([pad4, lrfs, pad4], 1), ids)
    return gp, gl, ga, idx, err + gerr


def neighbour_gather(point_set, lrf_set, idx, wrong_ids=None, wrong_count=None):
    """Batched loader gather on the device (reference
    my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160).
    point_set [B,N,3], lrf_set [B,N,4], idx [B,P,K] int64, wrong_ids [B,W] int64.
    Returns points [B,P,K,3], lrfs [B,P,K,4], act [B,P,K].  Raises IndexError on an
    out-of-range index like the reference's tensor indexing would."""
    B, N, _ = point_set.shape
    _, P, K = idx.shape
    point_set = _check(point_set, "point_set", (B, N, 3))
    lrf_set = _check(lrf_set, "lrf_set", (B, N, 4))
    if not idx.is_cuda or idx.dtype != torch.int64:
        raise TypeError("idx must be a CUDA int64 tensor")
    idx = idx.contiguous()
    W = 0
    if wrong_ids is not None:
        wrong_ids = wrong_ids.to(device=idx.device, dtype=torch.int64).contiguous().view(B, -1)
        W = wrong_ids.shape[1]
        if W == 0:
            wrong_ids = None
    if wrong_count is not None:
        wrong_count = wrong_count.to(device=idx.device, dtype=torch.int32).contiguous()
    dev = idx.device
    points = torch.empty(B, P, K, 3, device=dev, dtype=torch.float32)
    lrfs = torch.empty(B, P, K, 4, device=dev, dtype=torch.float32)
    act = torch.empty(B, P, K, device=dev, dtype=torch.float32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    launch("qec_gather", point_set, lrf_set, idx, wrong_ids, wrong_count, points,
         lrfs, act, err, B, N, P, K, W)
    return points, lrfs, act, err
