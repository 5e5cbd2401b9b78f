// routing_small.cu -- the pool1 regime (in_channels = 1, a handful of neighbours per patch,
// hundreds of output capsules per patch: 1 M capsules x 9 votes at BASELINE config 2) as ONE
// register-resident kernel per direction.
//
// Replaces, for I = 1, everything of QecModule.forward after the mean LRF
// (reference models/qec_module.py:123-147 get_votes incl. the two Linears, :172-194 routing).
// The two Linears have no non-linearity between them (qec_module.py:123-124), so for I = 1
//     t_raw[(bp,k), q*O + o] = Wc[q*O + o, :] . x[bp,k] + bc[q*O + o],
//     Wc = W2 W1  [4*O, 3],   bc = W2 b1 + b2  [4*O]
// is a rank-3 map: the host composes Wc / bc (0.5 M FLOP, autograd carries the chain back to
// W1, b1, W2, b2) and the kernel evaluates 12 FMA per vote instead of reading a 151 MB tensor
// that two 9.7 GFLOP SGEMMs produced.
//
// One thread owns one output channel o and keeps its 12 + 4 composed weights in registers
// for its whole life; it walks over patches bp, and for each (bp, o) capsule generates the K
// votes into registers, runs all T routing iterations there (weights b^t stateful, 32 Jacobi
// solves in lock-step per warp) and writes pose / agreement coalesced over o.  The K rotated
// points, LRFs and activations of the patch are shared by all o: staged once per patch in
// shared memory and read as broadcasts.  The backward recomputes the votes, re-runs the
// weight chain from the T saved poses (no Jacobi), applies the closed-form eigenvector
// adjoint, and accumulates dL/dWc, dL/dbc in 16 registers across all patches: no gradient
// tensor of the size of t_raw ever exists.
#include "routing_core.cuh"
#include "vote_math.cuh"

namespace qec {

constexpr int kST = 128;  // threads per CTA = output channels per CTA (64 / 32 for layers with that few channels: the
                          // stress sweep of BASELINE config 5 goes down to O = 32, where a 128-thread CTA idles 3 of 4 warps)

template <int KMAX>
struct PatchStage {  // per-patch inputs shared by all output channels
    float x[KMAX][3];
    float4 lrf[KMAX];
    float a[KMAX];
};

template <int KMAX, int ST>
__device__ __forceinline__ void stage_patch(PatchStage<KMAX>& s, const float* __restrict__ xrot,
                                            const float* __restrict__ lrfs, const float* __restrict__ act,
                                            long long bp, int K) {
    __syncthreads();  // previous patch fully consumed
#pragma unroll
    for (int k = threadIdx.x; k < KMAX; k += ST) {
        if (k < K) {
            const size_t e = (size_t)bp * K + k;
            s.x[k][0] = __ldg(xrot + e * 3);
            s.x[k][1] = __ldg(xrot + e * 3 + 1);
            s.x[k][2] = __ldg(xrot + e * 3 + 2);
            s.lrf[k] = __ldg(reinterpret_cast<const float4*>(lrfs) + e);
            s.a[k] = __ldg(act + e);
        } else {  // padding votes: zero LRF and zero activation contribute nothing anywhere
            s.x[k][0] = s.x[k][1] = s.x[k][2] = 0.f;
            s.lrf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            s.a[k] = 0.f;
        }
    }
    __syncthreads();
}

struct ComposedRow {  // this thread's rows of Wc / bc: w[q][d], b[q]
    float w[4][3];
    float b[4];
    __device__ __forceinline__ void load(const float* __restrict__ Wc, const float* __restrict__ bc, int o, int O) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t n = (size_t)q * O + o;
#pragma unroll
            for (int d = 0; d < 3; ++d) w[q][d] = __ldg(Wc + n * 3 + d);
            b[q] = __ldg(bc + n);
        }
    }
    __device__ __forceinline__ Quat apply(const float* x) const {
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = fmaf(w[q][0], x[0], fmaf(w[q][1], x[1], fmaf(w[q][2], x[2], b[q])));
        return qmake(t[0], t[1], t[2], t[3]);
    }
};

__device__ __forceinline__ Quat q_from(float4 v) { return qmake(v.x, v.y, v.z, v.w); }

template <int KMAX, int ST>
__global__ void __launch_bounds__(ST)
    routing_small_fwd_kernel(const float* __restrict__ xrot, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ Wc,
                             const float* __restrict__ bc, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, float* __restrict__ pose,
                             float* __restrict__ agreement, float* __restrict__ poses_all,
                             float* __restrict__ stats, int BP, int K, int O, int T) {
    __shared__ PatchStage<KMAX> st;
    const int o_raw = blockIdx.y * ST + threadIdx.x;
    const bool active = o_raw < O;
    const int o = active ? o_raw : O - 1;  // tail lanes shadow the last channel (warp-uniform solves)
    ComposedRow cr;
    cr.load(Wc, bc, o, O);
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t Nc = (size_t)BP * O;
    for (long long bp = blockIdx.x; bp < BP; bp += gridDim.x) {
        stage_patch<KMAX, ST>(st, xrot, lrfs, act, bp, K);
        Quat v[KMAX];
        float b[KMAX];
        float acc[13];  // M (10), sum |b|, sum of vote components, sum b
        zero_acc(acc);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            v[k] = make_vote(q_from(st.lrf[k]), cr.apply(st.x[k])).vote;
            b[k] = st.a[k];
            acc[10] += fabsf(b[k]);
            acc[12] += b[k];
            acc_rank1(acc, b[k], v[k]);
            acc[11] += (v[k].w + v[k].x) + (v[k].y + v[k].z);
        }
        const bool valid = acc[11] != 0.f;  // reference qec_module.py:72-73
        const size_t c = (size_t)bp * O + o;
        Quat p;
        for (int t = 0;; ++t) {
            const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));
            p = qfix(top_eigvec_weighted2(m, acc[12] != acc[10], WarpDone()));
            if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
            if (active && poses_all != nullptr) stq(poses_all + ((size_t)t * Nc + c) * 4, p);
            if (t == T - 1) break;
            zero_acc(acc);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                b[k] *= st.a[k] * one_minus_dist(qdot(p, v[k]));
                acc[10] += fabsf(b[k]);
                acc[12] += b[k];
                acc_rank1(acc, b[k], v[k]);
            }
        }
        float ssum = 0.f, n = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            ssum = fmaf(b[k], one_minus_dist(qdot(p, v[k])), ssum);
            n += (b[k] != 0.f) ? 1.f : 0.f;
        }
        if (active) {
            const float s = (n > 0.f) ? ssum / n : 0.f;
            stq(pose + c * 4, p);
            agreement[c] = (n > 0.f) ? sigmoidf(fmaf(alpha, s, beta - 1.f)) : 0.f;
            if (stats != nullptr) {
                stats[2 * c] = s;
                stats[2 * c + 1] = n;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Forward for 16 < K <= 64 (BASELINE config 5: the stress sweep runs K up to 64 at I = 1, inference only).
// Sixty-four votes do not fit the register file, and parking them in shared memory (20 B per vote and
// thread: 160 KB per CTA of 128 threads at K = 64) would leave one CTA of four warps per SM.  A vote is cheap
// to REGENERATE instead -- 12 FMA for the composed transform, a normalisation and one Hamilton product from
// the patch's staged LRFs / points, which every output channel shares -- so each sweep recomputes its votes
// and only the routing weights b^t live in shared memory ([K][128] floats, thread-contiguous: conflict-free).
// ~2x the arithmetic of a stored-vote sweep at 8 resident CTAs per SM instead of 1; nothing of the size of the
// vote tensor (34 GB at the largest corner of the sweep) or of t_raw exists.  Same results as the
// register-resident kernel (same per-vote code, same accumulation order over k).
// ------------------------------------------------------------------------------------
template <int KMAX, int ST>
__global__ void __launch_bounds__(ST)
    routing_small_regen_fwd_kernel(const float* __restrict__ xrot, const float* __restrict__ lrfs,
                                   const float* __restrict__ act, const float* __restrict__ Wc,
                                   const float* __restrict__ bc, const float* __restrict__ alpha_p,
                                   const float* __restrict__ beta_p, float* __restrict__ pose,
                                   float* __restrict__ agreement, float* __restrict__ poses_all,
                                   float* __restrict__ stats, int BP, int K, int O, int T) {
    __shared__ PatchStage<KMAX> st;
    extern __shared__ float sbw[];  // [K][kST] routing weights b^t of this thread's capsule
    float* bw = sbw + threadIdx.x;
    const int o_raw = blockIdx.y * ST + threadIdx.x;
    const bool active = o_raw < O;
    const int o = active ? o_raw : O - 1;
    ComposedRow cr;
    cr.load(Wc, bc, o, O);
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t Nc = (size_t)BP * O;
    for (long long bp = blockIdx.x; bp < BP; bp += gridDim.x) {
        stage_patch<KMAX, ST>(st, xrot, lrfs, act, bp, K);
        float acc[13];
        zero_acc(acc);
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const Quat v = make_vote(q_from(st.lrf[k]), cr.apply(st.x[k])).vote;
            const float b = st.a[k];
            bw[k * ST] = b;
            acc[10] += fabsf(b);
            acc[12] += b;
            acc_rank1(acc, b, v);
            acc[11] += (v.w + v.x) + (v.y + v.z);
        }
        const bool valid = acc[11] != 0.f;
        const size_t c = (size_t)bp * O + o;
        Quat p;
        for (int t = 0;; ++t) {
            const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));
            p = qfix(top_eigvec_weighted2(m, acc[12] != acc[10], WarpDone()));
            if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
            if (active && poses_all != nullptr) stq(poses_all + ((size_t)t * Nc + c) * 4, p);
            if (t == T - 1) break;
            zero_acc(acc);
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                const Quat v = make_vote(q_from(st.lrf[k]), cr.apply(st.x[k])).vote;
                const float b = bw[k * ST] * (st.a[k] * one_minus_dist(qdot(p, v)));
                bw[k * ST] = b;
                acc[10] += fabsf(b);
                acc[12] += b;
                acc_rank1(acc, b, v);
            }
        }
        float ssum = 0.f, n = 0.f;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            const Quat v = make_vote(q_from(st.lrf[k]), cr.apply(st.x[k])).vote;
            const float b = bw[k * ST];
            ssum = fmaf(b, one_minus_dist(qdot(p, v)), ssum);
            n += (b != 0.f) ? 1.f : 0.f;
        }
        if (active) {
            const float s = (n > 0.f) ? ssum / n : 0.f;
            stq(pose + c * 4, p);
            agreement[c] = (n > 0.f) ? sigmoidf(fmaf(alpha, s, beta - 1.f)) : 0.f;
            if (stats != nullptr) {
                stats[2 * c] = s;
                stats[2 * c + 1] = n;
            }
        }
    }
}

// backward for the case where only the parameters need gradients (pool1: the LRFs, points and
// 0/1 activations are data): dL/dWc [4O,3], dL/dbc [4O], dL/dalpha, dL/dbeta, all ACCUMULATED.
template <int KMAX, int ST>
__global__ void __launch_bounds__(ST, 512 / ST)
    routing_small_bwd_kernel(const float* __restrict__ xrot, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ Wc,
                             const float* __restrict__ bc, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, const float* __restrict__ poses_all,
                             const float* __restrict__ stats, const float* __restrict__ g_pose,
                             const float* __restrict__ g_agr, float* g_Wc, float* g_bc, float* g_alpha_beta, int BP,
                             int K, int O, int T) {
    __shared__ PatchStage<KMAX> st;
    const int o_raw = blockIdx.y * ST + threadIdx.x;
    const bool active = o_raw < O;
    const int o = active ? o_raw : O - 1;
    ComposedRow cr;
    cr.load(Wc, bc, o, O);
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t Nc = (size_t)BP * O;
    float gW[4][3], gB[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        gB[q] = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) gW[q][d] = 0.f;
    }
    float ga_tot = 0.f, gb_tot = 0.f;
    for (long long bp = blockIdx.x; bp < BP; bp += gridDim.x) {
        stage_patch<KMAX, ST>(st, xrot, lrfs, act, bp, K);
        const size_t c = (size_t)bp * O + o;
        const float s = __ldg(stats + 2 * c), n = __ldg(stats + 2 * c + 1);
        const bool has = n > 0.f;
        const float sg = sigmoidf(fmaf(alpha, s, beta - 1.f));
        const float gz = has ? __ldg(g_agr + c) * sg * (1.f - sg) : 0.f;
        if (active) {
            ga_tot += gz * s;
            gb_tot += gz;
        }
        const float gsn = has ? gz * alpha / n : 0.f;
        // votes and the weight chain b^{T-1} from the saved poses
        Quat v[KMAX];
        float b[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            v[k] = make_vote(q_from(st.lrf[k]), cr.apply(st.x[k])).vote;
            b[k] = st.a[k];
        }
        for (int t = 0; t < T - 1; ++t) {
            const Quat pt = ldq_nc(poses_all + ((size_t)t * Nc + c) * 4);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) b[k] *= st.a[k] * one_minus_dist(qdot(pt, v[k]));
        }
        const Quat p = ldq_nc(poses_all + ((size_t)(T - 1) * Nc + c) * 4);
        // level T-1: G_p, M, S
        float acc[15];
        zero_acc(acc);
        float x[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            x[k] = qdot(p, v[k]);
            const float gx = -gsn * b[k] * dist_grad(x[k]);
            acc[0] = fmaf(gx, v[k].w, acc[0]); acc[1] = fmaf(gx, v[k].x, acc[1]);
            acc[2] = fmaf(gx, v[k].y, acc[2]); acc[3] = fmaf(gx, v[k].z, acc[3]);
            acc_rank1(acc + 4, b[k], v[k]);
            acc[14] += fabsf(b[k]);
        }
        const float rs = 1.f / fmaxf(acc[14], 1e-12f);
        const Sym4 m = acc_to_sym(acc + 4, rs);
        const Quat gp = qadd(qmake(acc[0], acc[1], acc[2], acc[3]), ldq_nc(g_pose + c * 4));
        const float lam = qdot(p, sym_mul(m, p));
        Quat u = top_eigvec_adjoint(m, p, lam, gp);
        if (!(qdot(p, p) > 0.5f) || !active) u = qmake(0.f, 0.f, 0.f, 0.f);
        const float live = active ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const float uv = qdot(u, v[k]);
            const float gx = -gsn * b[k] * dist_grad(x[k]) * live;
            const float wh = b[k] * rs;
            Quat gv = qscale(p, fmaf(wh, uv, gx));
            gv = qfma(wh * x[k], u, gv);
            const Quat lrf = q_from(st.lrf[k]);
            const VoteParts vp = make_vote(lrf, cr.apply(st.x[k]));
            Quat gt, gl;
            vote_bwd(lrf, vp, gv, gt, gl);
            const float g4[4] = {gt.w, gt.x, gt.y, gt.z};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                gB[q] += g4[q];
#pragma unroll
                for (int d = 0; d < 3; ++d) gW[q][d] = fmaf(g4[q], st.x[k][d], gW[q][d]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t nrow = (size_t)q * O + o;
            atomicAdd(g_bc + nrow, gB[q]);
#pragma unroll
            for (int d = 0; d < 3; ++d) atomicAdd(g_Wc + nrow * 3 + d, gW[q][d]);
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        ga_tot += __shfl_xor_sync(0xffffffffu, ga_tot, sft);
        gb_tot += __shfl_xor_sync(0xffffffffu, gb_tot, sft);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(g_alpha_beta, ga_tot);
        atomicAdd(g_alpha_beta + 1, gb_tot);
    }
}

static int small_kmax(int K) { return K <= 9 ? 9 : (K <= 16 ? 16 : 0); }
static int small_threads(int O) { return O <= 32 ? 32 : (O <= 64 ? 64 : kST); }

// persistent grid: exactly as many CTAs as are co-resident (one wave), each walking over many patches.
// `cache`: CTAs per SM of this kernel instantiation, per device.
template <class Kern>
static dim3 small_grid(Kern kern, int threads, size_t smem, int BP, int O, PerDeviceInt& cache) {
    const int gy = (O + threads - 1) / threads;
    std::atomic<int>* slot = cache.slot();
    int per_sm = slot ? slot->load(std::memory_order_relaxed) : -1;
    if (per_sm < 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        if (slot) slot->store(per_sm, std::memory_order_relaxed);
    }
    int gx = (kNumSMs * per_sm) / gy;
    if (gx < 1) gx = 1;
    if (gx > BP) gx = BP;
    return dim3((unsigned)gx, (unsigned)gy);
}

template <int KMAX, int ST>
static void launch_small_fwd(const float* xrot, const float* lrfs, const float* act, const float* Wc, const float* bc,
                             const float* alpha, const float* beta, float* pose, float* agreement, float* poses_all,
                             float* stats, int BP, int K, int O, int T, cudaStream_t st) {
    static PerDeviceInt occ;
    const dim3 grid = small_grid(routing_small_fwd_kernel<KMAX, ST>, ST, 0, BP, O, occ);
    routing_small_fwd_kernel<KMAX, ST><<<grid, ST, 0, st>>>(xrot, lrfs, act, Wc, bc, alpha, beta, pose, agreement,
                                                            poses_all, stats, BP, K, O, T);
}
template <int ST>
static void launch_small_regen_fwd(const float* xrot, const float* lrfs, const float* act, const float* Wc,
                                   const float* bc, const float* alpha, const float* beta, float* pose, float* agreement,
                                   float* poses_all, float* stats, int BP, int K, int O, int T, cudaStream_t st) {
    static PerDeviceInt occ;  // sized for the largest K, so one cached value serves every K
    const dim3 grid = small_grid(routing_small_regen_fwd_kernel<64, ST>, ST, 64 * ST * sizeof(float), BP, O, occ);
    routing_small_regen_fwd_kernel<64, ST><<<grid, ST, (size_t)K * ST * sizeof(float), st>>>(
        xrot, lrfs, act, Wc, bc, alpha, beta, pose, agreement, poses_all, stats, BP, K, O, T);
}
template <int KMAX, int ST>
static void launch_small_bwd(const float* xrot, const float* lrfs, const float* act, const float* Wc, const float* bc,
                             const float* alpha, const float* beta, const float* poses_all, const float* stats,
                             const float* g_pose, const float* g_agreement, float* g_Wc, float* g_bc, float* g_ab, int BP,
                             int K, int O, int T, cudaStream_t st) {
    static PerDeviceInt occ;
    const dim3 grid = small_grid(routing_small_bwd_kernel<KMAX, ST>, ST, 0, BP, O, occ);
    routing_small_bwd_kernel<KMAX, ST><<<grid, ST, 0, st>>>(xrot, lrfs, act, Wc, bc, alpha, beta, poses_all, stats, g_pose,
                                                            g_agreement, g_Wc, g_bc, g_ab, BP, K, O, T);
}

}  // namespace qec

using namespace qec;

// I = 1 geometries served by the composed-weight kernels: 1 = forward and backward (K <= 16, votes in
// registers), 2 = forward only (16 < K <= 64, votes regenerated per sweep), 0 = not served
extern "C" int qec_routing_small_supported(int K, int I, int O, int T) {
    if (!(I == 1 && O > 0 && T >= 1 && T <= 10 && K >= 1)) return 0;
    return small_kmax(K) > 0 ? 1 : (K <= 64 ? 2 : 0);
}

extern "C" int qec_routing_small_fwd(const float* xrot, const float* lrfs, const float* act, const float* Wc,
                                     const float* bc, const float* alpha, const float* beta, float* pose,
                                     float* agreement, float* poses_all, float* stats, int BP, int K, int O, int T,
                                     void* stream) {
    QEC_CHECK_ARG(xrot, 1);
    QEC_CHECK_ARG(lrfs && act, 2);
    QEC_CHECK_ARG(Wc && bc, 4);
    QEC_CHECK_ARG(alpha && beta, 6);
    QEC_CHECK_ARG(pose && agreement, 8);  // poses_all / stats (what the backward needs) may be NULL for inference
    QEC_CHECK_ARG(BP > 0, 12);
    QEC_CHECK_ARG(qec_routing_small_supported(K, 1, O, T), 13);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int km = small_kmax(K), th = small_threads(O);
#define QEC_SMALL_FWD(ST_)                                                                                              \
    if (km == 0) launch_small_regen_fwd<ST_>(xrot, lrfs, act, Wc, bc, alpha, beta, pose, agreement, poses_all, stats, BP, K, O, T, st); \
    else if (km == 9) launch_small_fwd<9, ST_>(xrot, lrfs, act, Wc, bc, alpha, beta, pose, agreement, poses_all, stats, BP, K, O, T, st); \
    else launch_small_fwd<16, ST_>(xrot, lrfs, act, Wc, bc, alpha, beta, pose, agreement, poses_all, stats, BP, K, O, T, st);
    if (th == 32) { QEC_SMALL_FWD(32) } else if (th == 64) { QEC_SMALL_FWD(64) } else { QEC_SMALL_FWD(128) }
#undef QEC_SMALL_FWD
    return launch_status();
}

extern "C" int qec_routing_small_bwd(const float* xrot, const float* lrfs, const float* act, const float* Wc,
                                     const float* bc, const float* alpha, const float* beta, const float* poses_all,
                                     const float* stats, const float* g_pose, const float* g_agreement, float* g_Wc,
                                     float* g_bc, float* g_alpha_beta, int BP, int K, int O, int T, void* stream) {
    QEC_CHECK_ARG(xrot, 1);
    QEC_CHECK_ARG(lrfs && act, 2);
    QEC_CHECK_ARG(Wc && bc, 4);
    QEC_CHECK_ARG(alpha && beta && poses_all && stats, 6);
    QEC_CHECK_ARG(g_pose && g_agreement, 10);
    QEC_CHECK_ARG(g_Wc && g_bc && g_alpha_beta, 12);
    QEC_CHECK_ARG(BP > 0, 15);
    QEC_CHECK_ARG(qec_routing_small_supported(K, 1, O, T) == 1, 16);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int km = small_kmax(K), th = small_threads(O);
#define QEC_SMALL_BWD(ST_)                                                                                             \
    if (km == 9) launch_small_bwd<9, ST_>(xrot, lrfs, act, Wc, bc, alpha, beta, poses_all, stats, g_pose, g_agreement, g_Wc, g_bc, g_alpha_beta, BP, K, O, T, st); \
    else launch_small_bwd<16, ST_>(xrot, lrfs, act, Wc, bc, alpha, beta, poses_all, stats, g_pose, g_agreement, g_Wc, g_bc, g_alpha_beta, BP, K, O, T, st);
    if (th == 32) { QEC_SMALL_BWD(32) } else if (th == 64) { QEC_SMALL_BWD(64) } else { QEC_SMALL_BWD(128) }
#undef QEC_SMALL_BWD
    return launch_status();
}
