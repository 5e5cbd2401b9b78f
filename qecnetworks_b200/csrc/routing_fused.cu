// routing_fused.cu -- vote generation + the whole routing loop of one output capsule in
// ONE kernel with the capsule's votes RESIDENT ON CHIP (the pool2 regime: thousands of
// votes per capsule, e.g. 1 280 capsules x 32 768 votes at BASELINE config 2).
//
// Replaces get_votes (after the Linears) + the loop of QecModule.forward
// (reference models/qec_module.py:126-147, 172-194) without ever materialising the
// [B,P,O,K*I,4] vote tensor (671 MB at config 2) or its gradient.
//
// Layout.  The second Linear is issued with its weight rows permuted so that cuBLAS
// already writes the transform quaternions capsule-major ("OQI"):
//     t_oqi[(bp*K + k), o*4I + q*I + i]
// For capsule (bp,o) every neighbour k then owns one contiguous 4I-float chunk; with lanes
// running along i the four components are four coalesced 128-byte loads, the LRF
// lrfs[bp,k,i,:] is one coalesced float4 load and the activation one coalesced float load.
//
// Residency.  A thread-block cluster of CL CTAs (CL = 1, 2, 4, 8) owns one capsule at a
// time; CTA r holds votes [r*Jc, (r+1)*Jc), Jc <= 4096, as float4 in its shared memory
// together with the per-vote routing state, so every sweep after the first touches no
// global memory at all:
//   forward : vote + (activation, routing weight b)             24 B / vote,  2 CTAs / SM
//   backward: vote + <pose^t,v>, 1-d^t for all t + g_b + g_a    48 B / vote,  1 CTA  / SM
// Each sweep reduces <= 16 floats: a value-splitting warp butterfly (16 SHFL instead of 80),
// one shared-memory stage across the warps, then the CTA partial is published in shared
// memory, the cluster barriers, and every CTA sums all partials through distributed shared
// memory in the same fixed order -- all CTAs of the cluster then run the same Jacobi /
// Cholesky solve on identical inputs, so no pose broadcast across CTAs is needed.
// HBM traffic per capsule: the transforms once (forward), plus their gradient once (backward).
#include <cooperative_groups.h>

#include "routing_core.cuh"
#include "vote_math.cuh"

namespace cg = cooperative_groups;

namespace qec {

constexpr int kFT = 512;      // threads per CTA
constexpr int kFW = kFT / 32;
constexpr int kJcMax = 4096;  // resident votes per CTA
constexpr int kTB = 3;        // routing iterations the cached backward supports

// After this, lane L holds (in the return value) the warp-wide total of a[idx16(L)].
__device__ __forceinline__ int idx16(int lane) { return (lane >> 1) & 15; }
__device__ __forceinline__ float warp_reduce_scatter16(float (&a)[16], int lane) {
#pragma unroll
    for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < h) {
                const float send = up ? a[i] : a[i + h];
                const float keep = up ? a[i + h] : a[i];
                a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

template <int CL>
struct ClusterReducer {
    float* scratch;  // [16][kFW] warp partials
    float* part;     // [2][16] this CTA's published partial sums (double-buffered across sweeps)
    float* out;      // [16] solve result broadcast inside the CTA
    unsigned phase;
    __device__ __forceinline__ explicit ClusterReducer(float* s)
        : scratch(s), part(s + 16 * kFW), out(s + 16 * kFW + 32), phase(0) {}
    static constexpr int kSmemFloats = 16 * kFW + 48;

    // Sum a[0..N) over every thread of every CTA of the cluster.  On return the totals are
    // valid in all lanes of warp 0 of every CTA (identical bits in every CTA).
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N]) {
        static_assert(N <= 16, "accumulator too wide");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = (i < N) ? a[i] : 0.f;
        const float mine_v = warp_reduce_scatter16(w, lane);
        __syncthreads();  // scratch / out may still be read from the previous round
        if ((lane & 1) == 0) scratch[idx16(lane) * kFW + warp] = mine_v;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
            if (lane < 16) {
#pragma unroll
                for (int x = 0; x < kFW; ++x) s += scratch[lane * kFW + x];  // fixed order
            }
            if (CL > 1) {
                float* mine = part + (phase & 1u) * 16;
                if (lane < 16) mine[lane] = s;
                cg::this_cluster().sync();  // all threads of all CTAs (see below)
                s = 0.f;
                if (lane < 16) {
#pragma unroll
                    for (int r = 0; r < CL; ++r) s += cg::this_cluster().map_shared_rank(mine, r)[lane];
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = __shfl_sync(0xffffffffu, s, i);
        } else if (CL > 1) {
            cg::this_cluster().sync();
        }
        ++phase;
    }
    __device__ __forceinline__ bool solver() const { return threadIdx.x < 32; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = r[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = out[i];
    }
};

__device__ __forceinline__ float ldg_stream_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_add_q(float* p, Quat q) {  // REDG.E.ADD.F32x4
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(q.w), "f"(q.x), "f"(q.y), "f"(q.z)
                 : "memory");
}
__device__ __forceinline__ Quat lds_q(const float4* p) {
    const float4 v = *p;
    return qmake(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void sts_q(float4* p, Quat q) { *p = make_float4(q.w, q.x, q.y, q.z); }

// walks this thread's votes jl = tid, tid + kFT, ... of the CTA's slice, tracking (k, i) of
// j = j_begin + jl = k*I + i without a division per vote
struct VoteWalker {
    int k, i, dk, di, I;
    __device__ __forceinline__ VoteWalker(int j_begin, int I_) : I(I_) {
        const int j = j_begin + threadIdx.x;
        k = j / I;
        i = j - k * I;
        dk = kFT / I;
        di = kFT - dk * I;
    }
    __device__ __forceinline__ void next() {
        i += di;
        k += dk;
        if (i >= I) {
            i -= I;
            ++k;
        }
    }
};

__device__ __forceinline__ Quat solve_pose(const float* acc, bool valid) {
    const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));  // F.normalize(p=1) of the weights
    float lam;
    Quat p = qfix(top_eigvec(m, lam, WarpDone()));
    if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
    return p;
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int CL>
__global__ void __launch_bounds__(kFT, 2)
    routing_fused_fwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, float* __restrict__ pose,
                             float* __restrict__ agreement, float* __restrict__ poses_all,
                             float* __restrict__ stats, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sv = reinterpret_cast<float4*>(smem_raw);
    float2* sab = reinterpret_cast<float2*>(sv + kJcMax);  // (activation, routing weight b)
    ClusterReducer<CL> red(reinterpret_cast<float*>(sab + kJcMax));
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = (J + CL - 1) / CL;
    const int j_begin = min(J, rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t trow = (size_t)4 * I * O;
    const bool writer = (rank == 0) && (threadIdx.x == 0);

    for (long long c = cid; c < Nc; c += ncl) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        // ---- generate the votes (once) + sweep 0: M^0 = sum a v v^T, b^0 = a ----------------
        float acc[12];
        zero_acc(acc);
        {
            VoteWalker w(j_begin, I);
            const float* tbase = t_oqi + (size_t)o * 4 * I;
            for (int jl = threadIdx.x; jl < nloc; jl += kFT) {
                const size_t row = (size_t)bp * K + w.k;
                const float* tp = tbase + row * trow + w.i;
                const Quat traw = qmake(ldg_stream_f32(tp), ldg_stream_f32(tp + I), ldg_stream_f32(tp + 2 * I),
                                        ldg_stream_f32(tp + 3 * I));
                const Quat lrf = ldq_nc(lrfs + (row * I + w.i) * 4);
                const float a = __ldg(act + row * I + w.i);
                const Quat v = make_vote(lrf, traw).vote;
                sts_q(sv + jl, v);
                sab[jl] = make_float2(a, a);
                acc[10] += fabsf(a);
                acc_rank1(acc, a, v);
                acc[11] += (v.w + v.x) + (v.y + v.z);
                w.next();
            }
        }
        red.reduce(acc);
        bool valid = false;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (red.solver()) {
            valid = acc[11] != 0.f;  // reference qec_module.py:72-73
            const Quat p = solve_pose(acc, valid);
            r[0] = p.w; r[1] = p.x; r[2] = p.y; r[3] = p.z;
        }
        red.bcast(r);
        Quat p = qmake(r[0], r[1], r[2], r[3]);
        if (writer) stq(poses_all + (size_t)c * 4, p);
        // ---- sweeps 1..T-1: b <- b a (1 - d(pose, v)), M^t = sum b v v^T, all from shared memory
        for (int t = 1; t < T; ++t) {
            float a2[11];
            zero_acc(a2);
#pragma unroll 2
            for (int jl = threadIdx.x; jl < nloc; jl += kFT) {
                const Quat v = lds_q(sv + jl);
                const float2 ab = sab[jl];
                const float b = ab.y * (ab.x * one_minus_dist(qdot(p, v)));
                sab[jl].y = b;
                a2[10] += fabsf(b);
                acc_rank1(a2, b, v);
            }
            red.reduce(a2);
            if (red.solver()) {
                const Quat q = solve_pose(a2, valid);
                r[0] = q.w; r[1] = q.x; r[2] = q.y; r[3] = q.z;
            }
            red.bcast(r);
            p = qmake(r[0], r[1], r[2], r[3]);
            if (writer) stq(poses_all + ((size_t)t * Nc + c) * 4, p);
        }
        // ---- epilogue sweep: inlier score (reference qec_module.py:188-193) ---------------------
        float e[2] = {0.f, 0.f};
#pragma unroll 2
        for (int jl = threadIdx.x; jl < nloc; jl += kFT) {
            const Quat v = lds_q(sv + jl);
            const float b = sab[jl].y;
            e[0] = fmaf(b, one_minus_dist(qdot(p, v)), e[0]);
            e[1] += (b != 0.f) ? 1.f : 0.f;
        }
        red.reduce(e);
        if (rank == 0 && threadIdx.x == 0) {
            const float n = e[1];
            const float s = (n > 0.f) ? e[0] / n : 0.f;
            stq(pose + (size_t)c * 4, p);
            agreement[c] = (n > 0.f) ? sigmoidf(fmaf(alpha, s, beta - 1.f)) : 0.f;
            stats[2 * c] = s;
            stats[2 * c + 1] = n;
        }
        // the next capsule's reduce() starts with __syncthreads(): the resident votes are not
        // overwritten while a warp is still sweeping (each thread only re-reads its own slots)
    }
    if (CL > 1) cg::this_cluster().sync();  // no CTA may exit while a peer can still read its partials
}

// ------------------------------------------------------------------------------------
// backward (T <= kTB).  Levels t = T-1..0; sweep k completes level T-k (its adjoint vector u is
// known from the previous reduction) and accumulates level T-k-1.  Per-vote quantities of the
// forward chain (x^t = <pose^t, v>, 1 - d^t) are computed once in sweep 0 and cached in shared
// memory; g_b of the last completed level and the running activation gradient are carried in
// shared memory as well.  See routing_core.cuh for the algebra.
// ------------------------------------------------------------------------------------
struct BwdSmem {
    float4* sv;
    float* sx[kTB];
    float* som[kTB];
    float* sgb;
    float* sga;
};

struct BwdSweepArgs {
    BwdSmem sm;
    int nloc, j_begin, I, K;
    long long bp;
    size_t trow;
    const float* act;
    const float* lrfs;
    const float* tbase;
    float* gtbase;
    float* g_lrfs;     // nullable
    float* g_act_row;  // nullable: no activation gradient -> only the top level is processed
    Quat ut, pt;       // adjoint vector and pose of the level being completed
    float rst, gsn;
};

// One backward sweep over this CTA's resident votes: completes level TT (TOP: TT = T-1, the
// level whose votes are not detached -> writes dL/dt_raw, dL/dlrf) and accumulates level TT-1.
template <int TT, bool TOP>
__device__ __forceinline__ void bwd_sweep(const BwdSweepArgs& A, float (&acc)[15]) {
    const bool chain = A.g_act_row != nullptr;
    const bool more = chain && TT >= 1;
    const bool last = chain && TT == 0;
    const int I = A.I;
    VoteWalker w(A.j_begin, I);
    for (int jl = threadIdx.x; jl < A.nloc; jl += kFT) {
        const Quat v = lds_q(A.sm.sv + jl);
        const size_t row = (size_t)A.bp * A.K + w.k;
        const float a = __ldg(A.act + row * I + w.i);
        // b^TT and b^{TT-1} from the cached (1 - d) chain: b^{t+1} = b^t a (1 - d^t)
        float b = a, blow = a;
        if (TT >= 1) b = a * (a * A.sm.som[0][jl]);
        if (TT >= 2) {
            blow = b;
            b = b * (a * A.sm.som[1][jl]);
        }
        const float x_t = A.sm.sx[TT][jl];
        const float om_t = A.sm.som[TT][jl];
        const float uv = qdot(A.ut, v);
        float gb, ga;
        if (TOP) {
            gb = A.gsn * om_t;
            ga = 0.f;
            const float gx = -A.gsn * b * dist_grad(x_t);
            const float wh = b * A.rst;
            Quat gv = qscale(A.pt, fmaf(wh, uv, gx));  // p (gx + w <u,v>)
            gv = qfma(wh * x_t, A.ut, gv);             // + u w <p,v>
            const float* tp = A.tbase + row * A.trow + w.i;
            const Quat traw = qmake(__ldg(tp), __ldg(tp + I), __ldg(tp + 2 * I), __ldg(tp + 3 * I));
            const Quat lrf = ldq_nc(A.lrfs + (row * I + w.i) * 4);
            const VoteParts vp = make_vote(lrf, traw);
            Quat gt, gl;
            vote_bwd(lrf, vp, gv, gt, gl);
            float* gp = A.gtbase + row * A.trow + w.i;
            gp[0] = gt.w; gp[I] = gt.x; gp[2 * I] = gt.y; gp[3 * I] = gt.z;
            if (A.g_lrfs != nullptr) red_add_q(A.g_lrfs + (row * I + w.i) * 4, gl);  // sum over the O capsules
        } else {
            const float gprev = A.sm.sgb[jl];            // dL/db^{TT+1}
            ga = fmaf(gprev * b, om_t, A.sm.sga[jl]);    // b^{TT+1} = b^TT a (1 - d^TT)
            gb = gprev * a * om_t;
        }
        if (a != 0.f) gb = fmaf(uv * x_t, A.rst, gb);  // the M-path is gated by mask = [a != 0]
        if (more) {
            A.sm.sgb[jl] = gb;
            A.sm.sga[jl] = ga;
            // level TT-1: dL/dd_j = -g_b^TT b^{TT-1} a
            const float x_l = A.sm.sx[TT >= 1 ? TT - 1 : 0][jl];
            const float gx = -gb * blow * a * dist_grad(x_l);
            acc[0] = fmaf(gx, v.w, acc[0]); acc[1] = fmaf(gx, v.x, acc[1]);
            acc[2] = fmaf(gx, v.y, acc[2]); acc[3] = fmaf(gx, v.z, acc[3]);
            acc_rank1(acc + 4, blow, v);
            acc[14] += fabsf(blow);
        }
        if (last) atomicAdd(A.g_act_row + jl, ga + gb);  // + dL/db^0 (b^0 = a)
        w.next();
    }
}

template <int CL>
__global__ void __launch_bounds__(kFT, 1)
    routing_fused_bwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, const float* __restrict__ poses_all,
                             const float* __restrict__ stats, const float* __restrict__ g_pose,
                             const float* __restrict__ g_agr, float* __restrict__ g_toqi, float* g_lrfs, float* g_act,
                             float* g_alpha_beta, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem sm;
    sm.sv = reinterpret_cast<float4*>(smem_raw);
    float* f = reinterpret_cast<float*>(sm.sv + kJcMax);
#pragma unroll
    for (int t = 0; t < kTB; ++t) {
        sm.sx[t] = f + (2 * t) * kJcMax;
        sm.som[t] = f + (2 * t + 1) * kJcMax;
    }
    sm.sgb = f + 2 * kTB * kJcMax;
    sm.sga = sm.sgb + kJcMax;
    ClusterReducer<CL> red(sm.sga + kJcMax);
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = (J + CL - 1) / CL;
    const int j_begin = min(J, rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t trow = (size_t)4 * I * O;
    const bool need_chain = g_act != nullptr;
    float ga_tot = 0.f, gb_tot = 0.f;

    for (long long c = cid; c < Nc; c += ncl) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        const float* tbase = t_oqi + (size_t)o * 4 * I;
        float* gtbase = g_toqi + (size_t)o * 4 * I;
        // per-capsule scalars
        Quat ph[kTB], uh[kTB];
        float rS[kTB];
#pragma unroll
        for (int t = 0; t < kTB; ++t) {
            ph[t] = (t < T) ? ldq_nc(poses_all + ((size_t)t * Nc + c) * 4) : qmake(0.f, 0.f, 0.f, 0.f);
            uh[t] = qmake(0.f, 0.f, 0.f, 0.f);
            rS[t] = 0.f;
        }
        const float s = __ldg(stats + 2 * c), n = __ldg(stats + 2 * c + 1);
        const bool has = n > 0.f;
        const float sg = sigmoidf(fmaf(alpha, s, beta - 1.f));
        const float gz = has ? __ldg(g_agr + c) * sg * (1.f - sg) : 0.f;
        ga_tot += gz * s;
        gb_tot += gz;
        const float gsn = has ? gz * alpha / n : 0.f;
        const Quat gpose = ldq_nc(g_pose + (size_t)c * 4);

        // ---- sweep 0: generate votes, forward chain (cached), accumulate level T-1 -----------
        float acc[15];
        zero_acc(acc);
        {
            VoteWalker w(j_begin, I);
            for (int jl = threadIdx.x; jl < nloc; jl += kFT) {
                const size_t row = (size_t)bp * K + w.k;
                const float* tp = tbase + row * trow + w.i;
                const Quat traw = qmake(ldg_stream_f32(tp), ldg_stream_f32(tp + I), ldg_stream_f32(tp + 2 * I),
                                        ldg_stream_f32(tp + 3 * I));
                const Quat lrf = ldq_nc(lrfs + (row * I + w.i) * 4);
                const float a = __ldg(act + row * I + w.i);
                const Quat v = make_vote(lrf, traw).vote;
                sts_q(sm.sv + jl, v);
                float b = a, xl = 0.f, bl = 0.f;
#pragma unroll
                for (int t = 0; t < kTB; ++t) {
                    if (t < T) {
                        const float x = qdot(ph[t], v);
                        const float om = one_minus_dist(x);
                        sm.sx[t][jl] = x;
                        sm.som[t][jl] = om;
                        xl = x;
                        bl = b;  // b^t
                        b *= a * om;
                    }
                }
                // level T-1 (final distance): dL/dd_j = -gsn b^{T-1}_j
                const float gx = -gsn * bl * dist_grad(xl);
                acc[0] = fmaf(gx, v.w, acc[0]); acc[1] = fmaf(gx, v.x, acc[1]);
                acc[2] = fmaf(gx, v.y, acc[2]); acc[3] = fmaf(gx, v.z, acc[3]);
                acc_rank1(acc + 4, bl, v);
                acc[14] += fabsf(bl);
                w.next();
            }
        }
        // ---- sweeps k = 1..: complete level t = T-k, accumulate level t-1 -------------------
        const int nsweeps = need_chain ? T : 1;
        for (int k = 1; k <= nsweeps; ++k) {
            const int t = T - k;  // level being completed
            // solve u^t from the reduction of the previous sweep
            red.reduce(acc);
            float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (red.solver()) {
                Quat p = ph[0];
#pragma unroll
                for (int u = 1; u < kTB; ++u)
                    if (u == t) p = ph[u];
                const float rs = 1.f / fmaxf(acc[14], 1e-12f);
                const Sym4 m = acc_to_sym(acc + 4, rs);
                Quat gp = qmake(acc[0], acc[1], acc[2], acc[3]);
                if (t == T - 1) gp = qadd(gp, gpose);
                const float lam = qdot(p, sym_mul(m, p));
                Quat u = top_eigvec_adjoint(m, p, lam, gp);
                if (!(qdot(p, p) > 0.5f)) u = qmake(0.f, 0.f, 0.f, 0.f);  // invalid / sign-zeroed pose
                r[0] = u.w; r[1] = u.x; r[2] = u.y; r[3] = u.z; r[4] = rs;
            }
            red.bcast(r);
            BwdSweepArgs a;
            a.sm = sm; a.nloc = nloc; a.j_begin = j_begin; a.I = I; a.K = K; a.bp = bp; a.trow = trow;
            a.act = act; a.lrfs = lrfs; a.tbase = tbase; a.gtbase = gtbase; a.g_lrfs = g_lrfs;
            a.g_act_row = need_chain ? g_act + (size_t)bp * J + j_begin : nullptr;
            a.ut = qmake(r[0], r[1], r[2], r[3]); a.rst = r[4]; a.gsn = gsn;
            zero_acc(acc);
            const bool top = (k == 1);
            // the level index is a template parameter: the (1-d) chain, the cached loads and the
            // top-level vote-gradient code fold at compile time
            if (t == 2) { a.pt = ph[2]; bwd_sweep<2, true>(a, acc); }
            else if (t == 1) { a.pt = ph[1]; if (top) bwd_sweep<1, true>(a, acc); else bwd_sweep<1, false>(a, acc); }
            else { a.pt = ph[0]; if (top) bwd_sweep<0, true>(a, acc); else bwd_sweep<0, false>(a, acc); }
        }
    }
    if (rank == 0 && threadIdx.x == 0) {
        atomicAdd(g_alpha_beta, ga_tot);
        atomicAdd(g_alpha_beta + 1, gb_tot);
    }
    if (CL > 1) cg::this_cluster().sync();
}

constexpr size_t kRedBytes = sizeof(float) * (16 * kFW + 48);
constexpr size_t kFwdSmem = (sizeof(float4) + sizeof(float2)) * kJcMax + kRedBytes;
constexpr size_t kBwdSmem = (sizeof(float4) + sizeof(float) * (2 * kTB + 2)) * kJcMax + kRedBytes;

// persistent grid: as many clusters as can be co-resident, never more than there are capsules
template <class Kern>
static int fused_grid(Kern kern, int CL, size_t smem, long long Nc, cudaLaunchConfig_t& cfg,
                      cudaLaunchAttribute* attr, int& cached_clusters) {
    cfg.blockDim = dim3(kFT);
    cfg.dynamicSmemBytes = smem;
    cfg.numAttrs = 0;
    cfg.attrs = attr;
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.numAttrs = 1;
    }
    int clusters = cached_clusters;  // per kernel instantiation; every B200 gives the same answer
    if (clusters < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (CL > 1) {
            cfg.gridDim = dim3(CL * 2 * kNumSMs);  // upper bound for the occupancy query
            e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
            if (e != cudaSuccess) return (int)e;
        } else {
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFT, smem);
            if (e != cudaSuccess) return (int)e;
            clusters = per_sm * kNumSMs;
        }
        if (clusters < 1) return (int)cudaErrorLaunchOutOfResources;
        cached_clusters = clusters;
    }
    if (clusters > Nc) clusters = (int)Nc;
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    return 0;
}

static int pick_cluster(int J) {
    for (int cl = 1; cl <= 8; cl *= 2)
        if ((J + cl - 1) / cl <= kJcMax) return cl;
    return 0;
}

template <int CL>
static int launch_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                            const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                            long long Nc, int K, int I, int O, int T, cudaStream_t st) {
    auto kern = routing_fused_fwd_kernel<CL>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static int cached = -1;
    const int rc = fused_grid(kern, CL, kFwdSmem, Nc, cfg, attr, cached);
    if (rc) return rc;
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all,
                                             stats, Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

template <int CL>
static int launch_fused_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                            const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                            const float* g_agr, float* g_toqi, float* g_lrfs, float* g_act, float* g_ab, long long Nc,
                            int K, int I, int O, int T, cudaStream_t st) {
    auto kern = routing_fused_bwd_kernel<CL>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static int cached = -1;
    const int rc = fused_grid(kern, CL, kBwdSmem, Nc, cfg, attr, cached);
    if (rc) return rc;
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose,
                                             g_agr, g_toqi, g_lrfs, g_act, g_ab, Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

}  // namespace qec

using namespace qec;

#define QEC_DISPATCH_CL(FN, ...)               \
    switch (CL) {                              \
        case 1: return FN<1>(__VA_ARGS__);     \
        case 2: return FN<2>(__VA_ARGS__);     \
        case 4: return FN<4>(__VA_ARGS__);     \
        default: return FN<8>(__VA_ARGS__);    \
    }

// 1 if this geometry is served by the fused resident kernels, else 0 (callers then use
// qec_votes_* + qec_routing_*)
extern "C" int qec_routing_fused_supported(int K, int I, int O, int T) {
    if (K <= 0 || I <= 0 || O <= 0) return 0;
    const long long J = (long long)K * I;
    if (J > 8ll * kJcMax) return 0;
    return (T >= 1 && T <= kTB && pick_cluster((int)J) > 0) ? 1 : 0;
}

extern "C" int qec_routing_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                     const float* beta, float* pose, float* agreement, float* poses_all,
                                     float* stats, int BP, int K, int I, int O, int T, void* stream) {
    QEC_CHECK_ARG(t_oqi, 1);
    QEC_CHECK_ARG(lrfs, 2);
    QEC_CHECK_ARG(act && alpha && beta, 3);
    QEC_CHECK_ARG(pose && agreement && poses_all && stats, 6);
    QEC_CHECK_ARG(BP > 0, 10);
    QEC_CHECK_ARG(qec_routing_fused_supported(K, I, O, T), 11);
    const int CL = pick_cluster(K * I);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL(launch_fused_fwd, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I, O,
                    T, st)
}

extern "C" int qec_routing_fused_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                     const float* beta, const float* poses_all, const float* stats,
                                     const float* g_pose, const float* g_agreement, float* g_toqi, float* g_lrfs,
                                     float* g_act, float* g_alpha_beta, int BP, int K, int I, int O, int T,
                                     void* stream) {
    QEC_CHECK_ARG(t_oqi, 1);
    QEC_CHECK_ARG(lrfs, 2);
    QEC_CHECK_ARG(act && alpha && beta && poses_all && stats, 3);
    QEC_CHECK_ARG(g_pose && g_agreement, 8);
    QEC_CHECK_ARG(g_toqi && g_alpha_beta, 10);
    QEC_CHECK_ARG(BP > 0, 14);
    QEC_CHECK_ARG(qec_routing_fused_supported(K, I, O, T), 15);
    const int CL = pick_cluster(K * I);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL(launch_fused_bwd, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi,
                    g_lrfs, g_act, g_alpha_beta, Nc, K, I, O, T, st)
}
