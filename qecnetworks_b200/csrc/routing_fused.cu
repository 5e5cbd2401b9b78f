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

#include "routing_fused_shared.cuh"
#include "vote_math.cuh"

namespace cg = cooperative_groups;

namespace qec {

// walks this thread's votes jl = tid, tid + kFT, ... of the CTA's slice, tracking (k, i) of
// j = j_begin + jl = k*I + i without a division per vote
template <int NT>
struct VoteWalkerT {
    int k, i, dk, di, I;
    __device__ __forceinline__ VoteWalkerT(int j_begin, int I_, int skip = 0) : I(I_) {
        const int j = j_begin + threadIdx.x + skip;
        k = j / I;
        i = j - k * I;
        dk = NT / I;
        di = NT - dk * I;
    }
    __device__ __forceinline__ void next2() {  // advance by two strides
        next();
        next();
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

// ------------------------------------------------------------------------------------
// forward.  256 threads x 16 resident votes per thread, two CTAs per SM; every sweep handles
// two votes per loop iteration with the loads issued first (ILP 2).
// ------------------------------------------------------------------------------------
constexpr int kFwdT = 256;

template <int CL>
__global__ void __launch_bounds__(kFwdT, 2)
    routing_fused_fwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, float* __restrict__ pose,
                             float* __restrict__ agreement, float* __restrict__ poses_all,
                             float* __restrict__ stats, long long Nc, int K, int I, int O, int T) {
    constexpr int NT = kFwdT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* __restrict__ sv = reinterpret_cast<float4*>(smem_raw);
    float2* __restrict__ sab = reinterpret_cast<float2*>(sv + kJcMax);  // (activation, routing weight b)
    ClusterReducer<CL, NT> red(reinterpret_cast<float*>(sab + kJcMax));
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
        const float* tbase = t_oqi + (size_t)o * 4 * I;
        // ---- generate the votes (once) + sweep 0: M^0 = sum a v v^T, b^0 = a ----------------
        float acc[13];  // M (10), sum |a|, sum a, sum of all vote components
        zero_acc(acc);
        {
            struct Raw { Quat traw, lrf; float a; };
            auto load = [&](const VoteWalkerT<NT>& w) {
                Raw r;
                const size_t row = (size_t)bp * K + w.k;
                const float* tp = tbase + row * trow + w.i;
                r.traw = qmake(ldg_stream_f32(tp), ldg_stream_f32(tp + I), ldg_stream_f32(tp + 2 * I),
                               ldg_stream_f32(tp + 3 * I));
                r.lrf = ldq_nc(lrfs + (row * I + w.i) * 4);
                r.a = __ldg(act + row * I + w.i);
                return r;
            };
            auto use = [&](const Raw& r, int jl) {
                const Quat v = make_vote(r.lrf, r.traw).vote;
                sts_q(sv + jl, v);
                sab[jl] = make_float2(r.a, r.a);
                acc[10] += fabsf(r.a);
                acc[11] += r.a;
                acc_rank1(acc, r.a, v);
                acc[12] += (v.w + v.x) + (v.y + v.z);
            };
            VoteWalkerT<NT> w0(j_begin, I), w1(j_begin, I, NT);
            int jl = threadIdx.x;
            for (; jl + NT < nloc; jl += 2 * NT) {
                const Raw r0 = load(w0);
                const Raw r1 = load(w1);
                use(r0, jl);
                use(r1, jl + NT);
                w0.next2();
                w1.next2();
            }
            if (jl < nloc) use(load(w0), jl);
        }
        red.reduce(acc);
        bool valid = false;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (red.solver()) {
            valid = acc[12] != 0.f;  // reference qec_module.py:72-73
            const Quat p = solve_pose(acc, valid);
            r[0] = p.w; r[1] = p.x; r[2] = p.y; r[3] = p.z;
        }
        red.bcast(r);
        Quat p = qmake(r[0], r[1], r[2], r[3]);
        if (writer) stq(poses_all + (size_t)c * 4, p);
        // ---- sweeps 1..T-1: b <- b a (1 - d(pose, v)), M^t = sum b v v^T, all from shared memory
        for (int t = 1; t < T; ++t) {
            float a2[12];
            zero_acc(a2);
            auto step = [&](Quat v, float2 ab, int jl) {
                const float b = ab.y * (ab.x * one_minus_dist(qdot(p, v)));
                sab[jl].y = b;
                a2[10] += fabsf(b);
                a2[11] += b;
                acc_rank1(a2, b, v);
            };
            int jl = threadIdx.x;
            for (; jl + NT < nloc; jl += 2 * NT) {
                const Quat v0 = lds_q(sv + jl), v1 = lds_q(sv + jl + NT);
                const float2 ab0 = sab[jl], ab1 = sab[jl + NT];
                step(v0, ab0, jl);
                step(v1, ab1, jl + NT);
            }
            if (jl < nloc) step(lds_q(sv + jl), sab[jl], jl);
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
        {
            auto fin = [&](Quat v, float b) {
                e[0] = fmaf(b, one_minus_dist(qdot(p, v)), e[0]);
                e[1] += (b != 0.f) ? 1.f : 0.f;
            };
            int jl = threadIdx.x;
            for (; jl + NT < nloc; jl += 2 * NT) {
                const Quat v0 = lds_q(sv + jl), v1 = lds_q(sv + jl + NT);
                const float b0 = sab[jl].y, b1 = sab[jl + NT].y;
                fin(v0, b0);
                fin(v1, b1);
            }
            if (jl < nloc) fin(lds_q(sv + jl), sab[jl].y);
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
// backward (T <= kTB = 3).  Levels t = T-1..0; sweep k completes level TT = T-k (its adjoint
// vector u^TT is known from the previous reduction) and accumulates level TT-1.
//
// 256 threads x 16 resident votes per thread, two CTAs per SM (two capsules in flight per SM:
// one capsule's reduce / Cholesky latency hides behind the other's sweep).  Shared memory per
// vote: the vote (16 B) and the cached 1 - d^t of the levels below the top (8 B).  Everything
// else is recomputed per sweep from per-capsule registers (T poses, T adjoint vectors): the
// gradient w.r.t. b at level TT is rebuilt by walking down from the top level with two dot
// products per level, which is cheaper than keeping it in shared memory and lets the second
// CTA fit.  Two votes are processed per loop iteration with all loads issued first (ILP 2).
// See routing_core.cuh for the algebra.
// ------------------------------------------------------------------------------------
constexpr int kBT = 256;

struct BwdCaps {  // per-capsule state, registers
    Quat ph[kTB], uh[kTB];
    float rS[kTB];
    float gsn;
};

struct BwdGeom {
    const float4* sv;
    const float* som0;
    const float* som1;
    int nloc, I, K;
    long long bp;
    size_t trow;
    const float* act;
    const float* lrfs;
    const float* tbase;
    float* gtbase;
    float* g_lrfs;     // nullable
    float* g_act_row;  // nullable: no activation gradient -> only the top level is processed
};

// gradient w.r.t. b^TT (gb), running activation gradient (ga), and b^t for all levels
template <int T, int TT>
__device__ __forceinline__ void complete_levels(const BwdCaps& C, Quat v, float a, float om0, float om1, float& gb,
                                                float& ga, float& x_tt, float& uv_tt, float (&b)[kTB]) {
    b[0] = a;
    b[1] = a * (a * om0);
    b[2] = b[1] * (a * om1);
    gb = 0.f;
    ga = 0.f;
#pragma unroll
    for (int t = T - 1; t >= TT; --t) {
        const float x = qdot(C.ph[t], v);
        const float uv = qdot(C.uh[t], v);
        float direct;
        if (t == T - 1) {
            direct = C.gsn * one_minus_dist(x);
        } else {
            const float om = (t == 0) ? om0 : om1;  // cached 1 - d^t
            ga = fmaf(gb * b[t], om, ga);           // b^{t+1} = b^t a (1 - d^t)
            direct = gb * a * om;
        }
        gb = (a != 0.f) ? fmaf(uv * x, C.rS[t], direct) : direct;  // the M-path is gated by [a != 0]
        if (t == TT) {
            x_tt = x;
            uv_tt = uv;
        }
    }
}

struct VoteIn {  // everything a sweep loads for one vote
    Quat v;
    float a, om0, om1;
    Quat traw, lrf;  // only in the top sweep
};

template <int T, int TT>
__device__ __forceinline__ VoteIn bwd_load(const BwdGeom& G, int jl, const VoteWalkerT<kBT>& w) {
    VoteIn in;
    in.v = lds_q(G.sv + jl);
    const size_t row = (size_t)G.bp * G.K + w.k;
    in.a = __ldg(G.act + row * G.I + w.i);
    in.om0 = (T >= 2) ? G.som0[jl] : 1.f;
    in.om1 = (T >= 3) ? G.som1[jl] : 1.f;
    if (TT == T - 1) {
        const float* tp = G.tbase + row * G.trow + w.i;
        const int I = G.I;
        in.traw = qmake(__ldg(tp), __ldg(tp + I), __ldg(tp + 2 * I), __ldg(tp + 3 * I));
        in.lrf = ldq_nc(G.lrfs + (row * I + w.i) * 4);
    }
    return in;
}

template <int T, int TT>
__device__ __forceinline__ void bwd_vote(const BwdGeom& G, const BwdCaps& C, const VoteIn& in, int jl,
                                         const VoteWalkerT<kBT>& w, float (&acc)[15]) {
    const bool chain = G.g_act_row != nullptr;
    float gb, ga, x, uv, b[kTB];
    complete_levels<T, TT>(C, in.v, in.a, in.om0, in.om1, gb, ga, x, uv, b);
    if (TT == T - 1) {  // the level whose votes are not detached: dL/dvote -> dL/dt_raw, dL/dlrf
        const float gx = -C.gsn * b[TT] * dist_grad(x);
        const float wh = b[TT] * C.rS[TT];
        Quat gv = qscale(C.ph[TT], fmaf(wh, uv, gx));  // p (gx + w <u,v>)
        gv = qfma(wh * x, C.uh[TT], gv);               // + u w <p,v>
        const VoteParts vp = make_vote(in.lrf, in.traw);
        Quat gt, gl;
        vote_bwd(in.lrf, vp, gv, gt, gl);
        const size_t row = (size_t)G.bp * G.K + w.k;
        float* gp = G.gtbase + row * G.trow + w.i;
        const int I = G.I;
        gp[0] = gt.w; gp[I] = gt.x; gp[2 * I] = gt.y; gp[3 * I] = gt.z;
        if (G.g_lrfs != nullptr) red_add_q(G.g_lrfs + (row * I + w.i) * 4, gl);  // sum over the O capsules
    }
    if (TT >= 1 && chain) {  // level TT-1: dL/dd_j = -g_b^TT b^{TT-1} a
        const float xl = qdot(C.ph[TT >= 1 ? TT - 1 : 0], in.v);
        const float bl = b[TT >= 1 ? TT - 1 : 0];
        const float gx = -gb * bl * in.a * dist_grad(xl);
        acc[0] = fmaf(gx, in.v.w, acc[0]); acc[1] = fmaf(gx, in.v.x, acc[1]);
        acc[2] = fmaf(gx, in.v.y, acc[2]); acc[3] = fmaf(gx, in.v.z, acc[3]);
        acc_rank1(acc + 4, bl, in.v);
        acc[14] += fabsf(bl);
    }
    if (TT == 0 && chain) atomicAdd(G.g_act_row + jl, ga + gb);  // + dL/db^0 (b^0 = a)
}

template <int T, int TT>
__device__ __forceinline__ void bwd_sweep(const BwdGeom& G, const BwdCaps& C, int j_begin, float (&acc)[15]) {
    VoteWalkerT<kBT> w0(j_begin, G.I), w1(j_begin, G.I, kBT);
    int jl = threadIdx.x;
    for (; jl + kBT < G.nloc; jl += 2 * kBT) {  // two votes per iteration, loads first
        const VoteIn i0 = bwd_load<T, TT>(G, jl, w0);
        const VoteIn i1 = bwd_load<T, TT>(G, jl + kBT, w1);
        bwd_vote<T, TT>(G, C, i0, jl, w0, acc);
        bwd_vote<T, TT>(G, C, i1, jl + kBT, w1, acc);
        w0.next2();
        w1.next2();
    }
    if (jl < G.nloc) {
        const VoteIn i0 = bwd_load<T, TT>(G, jl, w0);
        bwd_vote<T, TT>(G, C, i0, jl, w0, acc);
    }
}

// sweep 0: generate the votes, cache 1 - d^t of the levels below the top, accumulate level T-1
template <int T>
__device__ __forceinline__ void bwd_sweep0(const BwdGeom& G, const BwdCaps& C, int j_begin, float4* sv, float* som0,
                                           float* som1, float (&acc)[15]) {
    VoteWalkerT<kBT> w(j_begin, G.I);
    const int I = G.I;
#pragma unroll 2
    for (int jl = threadIdx.x; jl < G.nloc; jl += kBT) {
        const size_t row = (size_t)G.bp * G.K + w.k;
        const float* tp = G.tbase + row * G.trow + w.i;
        const Quat traw = qmake(ldg_stream_f32(tp), ldg_stream_f32(tp + I), ldg_stream_f32(tp + 2 * I),
                                ldg_stream_f32(tp + 3 * I));
        const Quat lrf = ldq_nc(G.lrfs + (row * I + w.i) * 4);
        const float a = __ldg(G.act + row * I + w.i);
        const Quat v = make_vote(lrf, traw).vote;
        sts_q(sv + jl, v);
        float b = a;
        if (T >= 2) {
            const float om = one_minus_dist(qdot(C.ph[0], v));
            som0[jl] = om;
            b *= a * om;
        }
        if (T >= 3) {
            const float om = one_minus_dist(qdot(C.ph[1], v));
            som1[jl] = om;
            b *= a * om;
        }
        // level T-1 (final distance): dL/dd_j = -gsn b^{T-1}_j
        const float gx = -C.gsn * b * dist_grad(qdot(C.ph[T - 1], v));
        acc[0] = fmaf(gx, v.w, acc[0]); acc[1] = fmaf(gx, v.x, acc[1]);
        acc[2] = fmaf(gx, v.y, acc[2]); acc[3] = fmaf(gx, v.z, acc[3]);
        acc_rank1(acc + 4, b, v);
        acc[14] += fabsf(b);
        w.next();
    }
}

template <int CL>
__global__ void __launch_bounds__(kBT, 2)
    routing_fused_bwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, const float* __restrict__ poses_all,
                             const float* __restrict__ stats, const float* __restrict__ g_pose,
                             const float* __restrict__ g_agr, float* __restrict__ g_toqi, float* g_lrfs, float* g_act,
                             float* g_alpha_beta, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sv = reinterpret_cast<float4*>(smem_raw);
    float* som0 = reinterpret_cast<float*>(sv + kJcMax);
    float* som1 = som0 + kJcMax;
    ClusterReducer<CL, kBT> red(som1 + kJcMax);
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = (J + CL - 1) / CL;
    const int j_begin = min(J, rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const bool need_chain = g_act != nullptr;
    float ga_tot = 0.f, gb_tot = 0.f;

    for (long long c = cid; c < Nc; c += ncl) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        BwdGeom G;
        G.sv = sv; G.som0 = som0; G.som1 = som1; G.nloc = nloc; G.I = I; G.K = K; G.bp = bp;
        G.trow = (size_t)4 * I * O;
        G.act = act; G.lrfs = lrfs;
        G.tbase = t_oqi + (size_t)o * 4 * I;
        G.gtbase = g_toqi + (size_t)o * 4 * I;
        G.g_lrfs = g_lrfs;
        G.g_act_row = need_chain ? g_act + (size_t)bp * J + j_begin : nullptr;
        BwdCaps C;
#pragma unroll
        for (int t = 0; t < kTB; ++t) {
            C.ph[t] = (t < T) ? ldq_nc(poses_all + ((size_t)t * Nc + c) * 4) : qmake(0.f, 0.f, 0.f, 0.f);
            C.uh[t] = qmake(0.f, 0.f, 0.f, 0.f);
            C.rS[t] = 0.f;
        }
        const float s = __ldg(stats + 2 * c), n = __ldg(stats + 2 * c + 1);
        const bool has = n > 0.f;
        const float sg = sigmoidf(fmaf(alpha, s, beta - 1.f));
        const float gz = has ? __ldg(g_agr + c) * sg * (1.f - sg) : 0.f;
        ga_tot += gz * s;
        gb_tot += gz;
        C.gsn = has ? gz * alpha / n : 0.f;
        const Quat gpose = ldq_nc(g_pose + (size_t)c * 4);

        float acc[15];
        zero_acc(acc);
        if (T == 3) bwd_sweep0<3>(G, C, j_begin, sv, som0, som1, acc);
        else if (T == 2) bwd_sweep0<2>(G, C, j_begin, sv, som0, som1, acc);
        else bwd_sweep0<1>(G, C, j_begin, sv, som0, som1, acc);

        const int nsweeps = need_chain ? T : 1;
        for (int k = 1; k <= nsweeps; ++k) {
            const int t = T - k;  // level being completed
            red.reduce(acc);      // sums of the level accumulated by the previous sweep
            float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (red.solver()) {
                Quat p = C.ph[0];
#pragma unroll
                for (int u = 1; u < kTB; ++u)
                    if (u == t) p = C.ph[u];
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
#pragma unroll
            for (int u = 0; u < kTB; ++u)
                if (u == t) {
                    C.uh[u] = qmake(r[0], r[1], r[2], r[3]);
                    C.rS[u] = r[4];
                }
            zero_acc(acc);
            // total iterations and level are template parameters: the chain walk, the cached loads
            // and the top-level vote-gradient code fold at compile time
            switch (T * 4 + t) {
                case 3 * 4 + 2: bwd_sweep<3, 2>(G, C, j_begin, acc); break;
                case 3 * 4 + 1: bwd_sweep<3, 1>(G, C, j_begin, acc); break;
                case 3 * 4 + 0: bwd_sweep<3, 0>(G, C, j_begin, acc); break;
                case 2 * 4 + 1: bwd_sweep<2, 1>(G, C, j_begin, acc); break;
                case 2 * 4 + 0: bwd_sweep<2, 0>(G, C, j_begin, acc); break;
                default: bwd_sweep<1, 0>(G, C, j_begin, acc); break;
            }
        }
        // the next capsule's sweep 0 overwrites this thread's own slots only; the reducer's
        // scratch is protected by the barriers inside reduce()
    }
    if (rank == 0 && threadIdx.x == 0) {
        atomicAdd(g_alpha_beta, ga_tot);
        atomicAdd(g_alpha_beta + 1, gb_tot);
    }
    if (CL > 1) cg::this_cluster().sync();
}

constexpr size_t kRedBytes = sizeof(float) * (16 * (kFT / 32) + 48);  // sized for the wider CTA
constexpr size_t kFwdSmem = (sizeof(float4) + sizeof(float2)) * kJcMax + kRedBytes;
constexpr size_t kBwdSmem = (sizeof(float4) + sizeof(float) * (kTB - 1)) * kJcMax + kRedBytes;

template <int CL>
static int launch_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                            const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                            long long Nc, int K, int I, int O, int T, cudaStream_t st) {
    auto kern = routing_fused_fwd_kernel<CL>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static PerDeviceInt cached;
    const int rc = fused_grid(kern, CL, kFwdT, kFwdSmem, Nc, cfg, attr, cached);
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
    static PerDeviceInt cached;
    const int rc = fused_grid(kern, CL, kBT, kBwdSmem, Nc, cfg, attr, cached);
    if (rc) return rc;
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose,
                                             g_agr, g_toqi, g_lrfs, g_act, g_ab, Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

}  // namespace qec

using namespace qec;

// 1 if this geometry is served by the fused resident kernels, else 0 (callers then use
// qec_votes_* + qec_routing_*)
extern "C" int qec_routing_fused_supported(int K, int I, int O, int T) {
    if (K <= 0 || I <= 0 || O <= 0) return 0;
    const long long J = (long long)K * I;
    if (J > 8ll * kJcMax) return 0;
    return (T >= 1 && T <= kTB && pick_cluster((int)J) > 0) ? 1 : 0;
}

extern "C" int qec_routing_fused_fwd_scalar(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
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

extern "C" int qec_routing_fused_bwd_scalar(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
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
