"""QecModule -- drop-in for the reference routing layer (models/qec_module.py:11-194).

Same constructor, parameters, state_dict keys/shapes and forward signature; the
forward is the B200 kernel pipeline instead of the ATen op sequence:

    lrf_mean_rotate  (1 launch pair)   mean LRF per patch + point rotation
    quater_gen1/2    (cuBLAS SGEMM)    the two Linears stay dense library GEMMs
    votes            (1 launch)        normalise + fix + Hamilton + fix + transpose
    routing          (1 launch)        all T IRLS iterations + activation, fused

No host synchronisation anywhere (the reference syncs 2(1+T) times per forward:
nonzero() and cuSOLVER's info copy).
"""
import threading

import torch
from torch.nn import Linear, Module, Parameter

from . import functional as QF


# The TF32 switch of cuBLAS is process-global in PyTorch (torch.backends.cuda.matmul.allow_tf32) while
# the reference's multi-GPU mode, nn.DataParallel (train_cls.py:27), runs one Python thread per GPU
# through this module.  Every matmul this package issues therefore goes through _mm(): one lock, the
# switch set explicitly to what that product needs (on for the pre-split 3xTF32 products, OFF for the
# plain FP32 ones) and restored afterwards, so no product of ours can be launched with another
# thread's setting.  (The switch is read at launch time; the kernels themselves run asynchronously.)
_MM_LOCK = threading.Lock()


def _mm(tf32, fn, *args, **kw):
    with _MM_LOCK:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        try:
            return fn(*args, **kw)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev


class _LinearFP32(torch.autograd.Function):
    """F.linear in plain FP32 (quater_gen1, reference qec_module.py:123), forward and backward under _mm()"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return _mm(False, torch.addmm, bias, x, weight.t())

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        g_x = _mm(False, torch.mm, g, weight) if ctx.needs_input_grad[0] else None
        g_w = _mm(False, torch.mm, g.t(), x) if ctx.needs_input_grad[1] else None
        g_b = g.sum(0) if ctx.needs_input_grad[2] else None
        return g_x, g_w, g_b


class _ComposeFP32(torch.autograd.Function):
    """(W1, b1, W2, b2) -> (Wc = W2 W1, bc = W2 b1 + b2): the two Linears of get_votes composed (no
    non-linearity between them, reference qec_module.py:123-124); plain FP32 under _mm() both ways"""

    @staticmethod
    def forward(ctx, W1, b1, W2, b2):
        ctx.save_for_backward(W1, b1, W2)
        return _mm(False, torch.mm, W2, W1), _mm(False, torch.addmv, b2, W2, b1)

    @staticmethod
    def backward(ctx, g_Wc, g_bc):
        W1, b1, W2 = ctx.saved_tensors
        g_Wc, g_bc = g_Wc.contiguous(), g_bc.contiguous()
        g_W1 = _mm(False, torch.mm, W2.t(), g_Wc)
        g_b1 = _mm(False, torch.mv, W2.t(), g_bc)
        g_W2 = _mm(False, torch.addmm, torch.outer(g_bc, b1), g_Wc, W1.t())
        return g_W1, g_b1, g_W2, g_bc


def _tf32_round(x):
    """round-to-nearest onto the TF32 grid (10-bit mantissa); the result is exactly representable"""
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _split_tf32_matmul(xa, wa):
    """xa @ wa^T in FP32 accuracy on the TF32 tensor cores (cuBLAS).

    Both operands are split x = hi + lo with hi, lo on the TF32 grid (exact inputs for the tensor
    cores) and the three significant products are ONE GEMM over a concatenated contraction axis:
        [x_hi | x_hi | x_lo] @ [w_hi | w_lo | w_hi]^T = x_hi w_hi + x_hi w_lo + x_lo w_hi
    accumulated in FP32; the dropped lo*lo term and the rounding of lo are ~2^-22 relative.
    Measured on B200 for [8192,41] x [41,20480]: 122 us vs 466 us for the SIMT SGEMM (which is
    bound by its 1.4 TB/s output path, not by FLOPs), max error 9e-7 of the largest entry vs
    3e-7 for SGEMM."""
    K = xa.shape[1]
    pad = (-3 * K) % 16  # tensor-core kernels need an aligned contraction length
    xh = _tf32_round(xa)
    xl = _tf32_round(xa - xh)
    wh = _tf32_round(wa)
    wl = _tf32_round(wa - wh)
    A = torch.cat([xh, xh, xl, xa.new_zeros(xa.shape[0], pad)], dim=1)
    Bm = torch.cat([wh, wl, wh, wa.new_zeros(wa.shape[0], pad)], dim=1)
    return _mm(True, torch.matmul, A, Bm.t())


class _SplitTf32Matmul(torch.autograd.Function):
    """y = xa @ wa^T (_split_tf32_matmul); the backward is plain FP32 SGEMM."""

    @staticmethod
    def forward(ctx, xa, wa):
        ctx.save_for_backward(xa, wa)
        return _split_tf32_matmul(xa, wa)

    @staticmethod
    def backward(ctx, g):
        xa, wa = ctx.saved_tensors
        g_xa = _mm(False, torch.mm, g, wa) if ctx.needs_input_grad[0] else None
        g_wa = _mm(False, torch.mm, g.t(), xa) if ctx.needs_input_grad[1] else None
        return g_xa, g_wa


# Backward of quater_gen2 on the package's own tcgen05 kernel (1) or on cuBLAS over the pre-split gradient (0).
import os as _os

GEN2_BACKWARD_TCGEN05 = _os.environ.get("QEC_GEN2_TC", "1") != "0"
# Forward of quater_gen2 on the package's own tcgen05 kernel (1) or as a cuBLAS 3xTF32 GEMM (0, default).  Measured at
# BASELINE config 2 (671 MB of output): own kernel 158 us incl. its operand prep, cuBLAS 120 us + 14 us of operand
# launches -- the own kernel moves 336 KB through shared memory per 64 KB output tile (operand fill, three operand
# reads, staging for the TMA store) and is bound by that; cuBLAS pairs two SMs per tile.  Step 2.524 vs 2.500 ms.
GEN2_FORWARD_TCGEN05 = _os.environ.get("QEC_GEN2_FWD_TC", "0") != "0"
# quater_gen1 (O <= 64) and the composition of the two Linears on the package's own FP32 kernels (1) or on cuBLAS (0)
OWN_SMALL_LINEARS = _os.environ.get("QEC_OWN_LINEARS", "1") != "0"


def _chunks(n, length=1024):
    """number of chunks (a power of two dividing n) that makes them about `length` long"""
    s = 1
    while n % (2 * s) == 0 and n // (2 * s) >= length:
        s *= 2
    return s


class _TransformRoutingFused(torch.autograd.Function):
    """quater_gen2 (t = [h, 1] @ [W2, b2]^T with the rows of W2 in capsule-major order) + the cluster-resident
    routing kernels as ONE autograd node (reference qec_module.py:124 + 126-147 + 172-194), so that

    * the GEMM operands are built by one launch each (qec_split_operand: bias column, row permutation, TF32
      hi/lo split, padding) instead of ~17 ATen launches, and the forward is one 3xTF32 tensor-core GEMM
      [h_hi | h_hi | h_lo] @ [w_hi | w_lo | w_hi]^T (K = 3*41 -> 128): 122 us instead of 466 us for the SIMT
      SGEMM, max error 9e-7 of the largest entry;
    * the 671 MB gradient of the transforms travels from the routing backward to the two weight / input
      gradient GEMMs pre-split for the tensor cores (TF32 hi plane + BF16 remainder), never as a plain FP32
      tensor:  g @ w = g_hi (w_hi + w_lo) + g_lo w_hi  -- a TF32 bmm and a BF16 bmm per product, with the
      long contraction (20 480 resp. 8 192 terms) cut into chunks of ~1 000 whose partial products are
      summed in FP32 (the tensor-core accumulator alone loses ~5e-5 over such lengths).  Error 1e-6 .. 6e-6
      of the largest entry, the same as the SIMT SGEMM it replaces, at 40 % of its time
      (scripts/gemm_probe7.py: 430 us vs 1 090 us for both products);
    * qec_fold_chunks sums the partial products and scatters them straight into dL/dW2 (rows permuted
      back), dL/db2 and dL/dh."""

    @staticmethod
    def forward(ctx, h, W2, b2, perm, lrfs, activation, alpha, beta, num_iterations):
        ctx.set_materialize_grads(False)  # an unused output (the pose under a pure classification loss) arrives as None
        T = int(num_iterations)
        if GEN2_FORWARD_TCGEN05 and QF.gen2_fwd_supported(h.shape[0], W2.shape[0], h.shape[1]):
            t_oqi = QF.gen2_forward(h, W2, b2, perm)  # csrc/gen2_fwd.cu: TMA + tcgen05, 3-term TF32 split
        else:
            K = h.shape[1] + 1
            W3 = 3 * K + ((-3 * K) % 16)  # tensor-core kernels need an aligned contraction length
            A, _ = QF.split_operand(h, None, 1.0, None, True, (1, 1, 2), K, W3)
            Bm, _ = QF.split_operand(W2, b2, 0.0, perm, True, (1, 2, 1), K, W3)
            t_oqi = _mm(True, torch.matmul, A, Bm.t())
        pose, agreement, poses_all, stats = QF.routing_fused_forward_raw(t_oqi, lrfs, activation, alpha, beta, T)
        ctx.save_for_backward(h, W2, b2, perm, t_oqi, lrfs, activation, alpha, beta, poses_all, stats)
        ctx.T = T
        return pose, agreement

    @staticmethod
    def backward(ctx, g_pose, g_agreement):
        h, W2, b2, perm, t_oqi, lrfs, activation, alpha, beta, poses_all, stats = ctx.saved_tensors
        need_h, need_w = ctx.needs_input_grad[0], (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        if GEN2_BACKWARD_TCGEN05 and QF.gen2_bwd_supported(t_oqi.shape[0], t_oqi.shape[1], h.shape[1]):
            # the transform gradient leaves the routing backward as plain FP32 and both products of the Linear's
            # backward run on this package's own tcgen05 kernel (csrc/gen2_bwd.cu) instead of four cuBLAS bmm's
            # (parameter gradients go straight into the flat gradient buffer when the parameters have a home there)
            g_t, g_lrfs, g_act, g_ab = QF.routing_fused_backward_raw(
                t_oqi, lrfs, activation, alpha, beta, poses_all, stats, g_pose, g_agreement, ctx.T,
                ctx.needs_input_grad[4], ctx.needs_input_grad[5], split=False, g_ab=QF._home_pair_out(alpha, beta))
            g_h, g_W2, g_b2 = QF.gen2_backward(g_t, h, W2, perm, need_h, need_w,
                                               out_W2=QF._home_out(W2) if need_w else None,
                                               out_b2=QF._home_out(b2) if need_w else None)
            return g_h, g_W2, g_b2, None, g_lrfs, g_act, g_ab[0:1], g_ab[1:2], None
        (g_hi, g_lo), g_lrfs, g_act, g_ab = QF.routing_fused_backward_raw(
            t_oqi, lrfs, activation, alpha, beta, poses_all, stats, g_pose, g_agreement, ctx.T,
            ctx.needs_input_grad[4], ctx.needs_input_grad[5], split=True)
        M, N = g_hi.shape
        K = h.shape[1] + 1
        KP = K + ((-K) % 8)
        g_h = g_W2 = g_b2 = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:  # dL/d[W2, b2] = g^T @ [h, 1]
            S = _chunks(M)
            X0, X1 = QF.split_operand(h, None, 1.0, None, True, (1, 2), KP, 2 * KP, KP)
            a = _mm(True, torch.bmm, g_hi.view(S, M // S, N).transpose(1, 2), X0.view(S, M // S, 2 * KP))
            b = _mm(False, torch.bmm, g_lo.view(S, M // S, N).transpose(1, 2), X1.view(S, M // S, KP),
                    out_dtype=torch.float32)
            g_W2 = torch.empty_like(W2)
            g_b2 = torch.empty_like(b2)
            QF.fold_chunks(a, b, perm, g_W2, g_b2, K)
        if ctx.needs_input_grad[0]:  # dL/dh = g @ [W2, b2] (the column of the bias is dropped)
            S = _chunks(N)
            W0, W1 = QF.split_operand(W2, b2, 0.0, perm, True, (1, 2), KP, 2 * KP, KP)
            a = _mm(True, torch.bmm, g_hi.view(M, S, N // S).transpose(0, 1), W0.view(S, N // S, 2 * KP))
            b = _mm(False, torch.bmm, g_lo.view(M, S, N // S).transpose(0, 1), W1.view(S, N // S, KP),
                    out_dtype=torch.float32)
            g_h = torch.empty_like(h)
            QF.fold_chunks(a, b, None, g_h, None, K)
        return g_h, g_W2, g_b2, None, g_lrfs, g_act, g_ab[0:1], g_ab[1:2], None


def _augment(x, weight, bias, row_perm=None):
    """([x, 1], [W, b]) so that F.linear(x, W, b) == [x, 1] @ [W, b]^T: the bias rides inside the GEMM
    (K+1) instead of a separate epilogue pass over the 671 MB output, and its gradient falls out of the
    weight-gradient GEMM instead of a separate 671 MB reduction"""
    xa = torch.cat([x, x.new_ones(x.shape[0], 1)], dim=1)
    wa = torch.cat([weight, bias.unsqueeze(1)], dim=1)
    if row_perm is not None:
        wa = wa.index_select(0, row_perm)  # output features re-ordered for free inside the GEMM
    return xa, wa


def _linear_bias_in_gemm(x, weight, bias, row_perm=None):
    """y = [x, 1] @ [W, b]^T == F.linear(x, W, b) on the tensor cores at FP32 accuracy"""
    xa, wa = _augment(x, weight, bias, row_perm)
    return _SplitTf32Matmul.apply(xa, wa)


class QecModule(Module):
    def __init__(self, in_channels, out_channels, num_iterations=3, num_neighbours=8, num_patches=8):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_iterations = num_iterations
        self.num_neighbours = num_neighbours
        self.num_patches = num_patches
        self.bottle_neck = in_channels * out_channels

        self.quater_gen1 = Linear(3 * in_channels, out_channels, bias=True)
        self.quater_gen2 = Linear(out_channels, out_channels * in_channels * 4, bias=True)
        self.alpha = Parameter(torch.ones(1))
        self.beta = Parameter(torch.ones(1))

    def reset_parameters(self):
        self.alpha.data.fill_(1)
        self.beta.data.fill_(1)

    # Inference on a large batch (BASELINE config 5: B = 1024): the paths that materialise per-vote tensors
    # (transforms: 16 B per vote; transforms + votes: 32 B per vote) run over slices of the batch so that no
    # temporary exceeds this many bytes.  Every capsule depends on one sample only, so slicing changes nothing.
    INFERENCE_TEMP_BYTES = 2 << 30

    def forward(self, points, lrfs, activation):
        """points [B,P,K,3], lrfs [B,P,K,I,4], activation [B,P,K,I]
        -> pose [B,P,O,4] (O squeezed when 1, as the reference's pose.squeeze(-2)),
           agreement [B,P,O]."""
        B, P, K, I, _ = lrfs.shape
        O = self.out_channels
        if I != self.in_channels or K != self.num_neighbours or P != self.num_patches:
            raise ValueError(
                "QecModule(in=%d, K=%d, P=%d) got lrfs %s"
                % (self.in_channels, self.num_neighbours, self.num_patches, tuple(lrfs.shape))
            )
        if not torch.is_grad_enabled() and B > 1:
            small = QF.routing_small_supported(K, I, O, self.num_iterations) if I == 1 else 0
            per_sample = 0 if small else 32 * P * K * I * O  # (the cluster path needs 16: bounded by the same cap)
            if per_sample * B > self.INFERENCE_TEMP_BYTES:
                step = max(1, self.INFERENCE_TEMP_BYTES // per_sample)
                outs = [self._forward(points[b:b + step], lrfs[b:b + step], activation[b:b + step])
                        for b in range(0, B, step)]
                return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
        return self._forward(points, lrfs, activation)

    def _forward(self, points, lrfs, activation):
        B, P, K, I, _ = lrfs.shape
        O = self.out_channels
        points = points.float().contiguous()
        lrfs = lrfs.float().contiguous()
        activation = activation.float().contiguous()
        _, xrot = QF.lrf_mean_rotate(lrfs, activation, points)
        act_flat = activation.view(B, P, K * I)
        data_grad = torch.is_grad_enabled() and (points.requires_grad or lrfs.requires_grad or activation.requires_grad)
        # (the tensors themselves, not self.parameters(): the replicas nn.DataParallel builds hold their parameters as
        # plain non-leaf tensors and list none)
        own = (self.quater_gen1.weight, self.quater_gen1.bias, self.quater_gen2.weight, self.quater_gen2.bias,
               self.alpha, self.beta)
        param_grad = torch.is_grad_enabled() and any(t.requires_grad for t in own)
        small = QF.routing_small_supported(K, I, O, self.num_iterations) if I == 1 else 0
        if not data_grad and (small == 1 or (small == 2 and not param_grad)):
            # no non-linearity between the two Linears: compose them (rank 3 for I = 1) and let the
            # register-resident kernel evaluate Wc x + bc per vote -- no [B*P*K, 4*O] tensor, no GEMM
            W1, b1 = self.quater_gen1.weight, self.quater_gen1.bias
            W2, b2 = self.quater_gen2.weight, self.quater_gen2.bias
            compose = QF.ComposeLinears if (OWN_SMALL_LINEARS and QF.compose_supported(W1.shape[1])) else _ComposeFP32
            if param_grad:
                Wc, bc = compose.apply(W1, b1, W2, b2)
            else:
                # inference: the composition depends on the parameters only -- keep it until one of them changes
                # (tensor version counters / storage), two launches less per forward
                key = tuple((t.data_ptr(), t._version, t.device) for t in (W1, b1, W2, b2))
                cached = getattr(self, "_composed", None)
                if cached is None or cached[0] != key:
                    with torch.no_grad():
                        cached = (key, compose.apply(W1, b1, W2, b2))
                    self._composed = cached
                Wc, bc = cached[1]
            pose, agreement = QF.routing_small(xrot, lrfs, act_flat, Wc, bc, self.alpha, self.beta,
                                               self.num_iterations)
            return pose.squeeze(-2), agreement
        linear1 = QF.LinearSmall if (OWN_SMALL_LINEARS and QF.linear_supported(3 * I, O)) else _LinearFP32
        h = linear1.apply(xrot.view(B * P * K, 3 * I), self.quater_gen1.weight, self.quater_gen1.bias)
        if K * I >= self.FUSED_MIN_VOTES and QF.routing_fused_supported(K, I, O, self.num_iterations):
            # capsule-major transforms straight out of the GEMM; votes never touch HBM
            if QF.routing_fused_packed_supported(K, I, O, self.num_iterations):
                pose, agreement = _TransformRoutingFused.apply(
                    h, self.quater_gen2.weight, self.quater_gen2.bias, self._oqi_perm(I, O), lrfs, act_flat,
                    self.alpha, self.beta, self.num_iterations)
            else:
                t_oqi = _linear_bias_in_gemm(h, self.quater_gen2.weight, self.quater_gen2.bias,
                                             self._oqi_perm(I, O))
                pose, agreement = QF.routing_fused(t_oqi, lrfs, act_flat, self.alpha, self.beta,
                                                   self.num_iterations)
        else:
            t_raw = _linear_bias_in_gemm(h, self.quater_gen2.weight, self.quater_gen2.bias)
            votes = QF.votes(lrfs, t_raw)
            pose, agreement = QF.routing(votes, act_flat, self.alpha, self.beta, self.num_iterations)
        return pose.squeeze(-2), agreement

    FUSED_MIN_VOTES = 1024  # below this a whole CTA per capsule is wasteful: general path

    def _oqi_perm(self, I, O):
        """row permutation of quater_gen2 that turns the reference feature order q*I*O + i*O + o
        (qec_module.py:126) into the capsule-major order o*4I + q*I + i"""
        dev = self.quater_gen2.weight.device
        perm = getattr(self, "_perm_cache", None)
        if perm is None or perm.device != dev or perm.numel() != 4 * I * O:
            o = torch.arange(O, device=dev).view(O, 1, 1)
            q = torch.arange(4, device=dev).view(1, 4, 1)
            i = torch.arange(I, device=dev).view(1, 1, I)
            perm = (q * (I * O) + i * O + o).reshape(-1)
            self._perm_cache = perm
        return perm
