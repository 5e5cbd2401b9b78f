// routing_pair.cuh -- building blocks of the packed-pair cluster-resident routing kernels that do not depend on
// the CTA shape: shared-memory pair layout, cp.async, cluster primitives (mapa / st.async / barrier.cluster), vote
// generation and its adjoint for two votes at once, the per-capsule pose solve, the level walk of the backward.
// Used by routing_fused2.cu (one capsule per cluster, CTA-wide barriers) and routing_fused3.cu (two capsules in
// flight per cluster, dedicated solver warp).
#pragma once

#include "f32x2.cuh"
#include "routing_core.cuh"
#include "routing_fused_shared.cuh"

namespace qec {

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ Quat2 lds_pair(const float4* sA, const float4* sB, int ps) {
    const float4 a = sA[ps], b = sB[ps];
    Quat2 v;
    v.w = pk(a.x, a.y); v.x = pk(a.z, a.w); v.y = pk(b.x, b.y); v.z = pk(b.z, b.w);
    return v;
}
__device__ __forceinline__ void sts_pair(float4* sA, float4* sB, int ps, const Quat2& v) {
    sA[ps] = make_float4(lo(v.w), hi(v.w), lo(v.x), hi(v.x));
    sB[ps] = make_float4(lo(v.y), hi(v.y), lo(v.z), hi(v.z));
}
__device__ __forceinline__ f2 lds_f2(const float2* p) {
    const float2 v = *p;
    return pk(v.x, v.y);
}
__device__ __forceinline__ void sts_f2(float2* p, f2 v) { *p = make_float2(lo(v), hi(v)); }
__device__ __forceinline__ f2 ldg_f2(const float* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return pk(v.x, v.y);
}


__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32(unsigned remote_addr, float v, unsigned remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}


// round-to-nearest onto the TF32 grid (10-bit mantissa), both halves; the result is exactly representable
__device__ __forceinline__ f2 tf32_round2(f2 v) {
    const unsigned a = (__float_as_uint(lo(v)) + 0x1000u) & 0xffffe000u, b = (__float_as_uint(hi(v)) + 0x1000u) & 0xffffe000u;
    return pk(__uint_as_float(a), __uint_as_float(b));
}

// (lo, hi) -> two BF16 (round to nearest even) in one word, lo at the lower address
__device__ __forceinline__ unsigned bf16x2(f2 v) {
    unsigned d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi(v)), "f"(lo(v)));
    return d;
}

// two votes at once (reference models/qec_module.py:129-145, scalar twin: make_vote in vote_math.cuh)
struct VotePair {
    Quat2 vote;
    Quat2 that;  // normalised transform before the hemisphere fix
    Quat2 t;     // fix(that)
    f2 rn, st, sv;
};
// sign(x) with sign(0) = 0 as clamp(x * 2^126, -1, 1): exact for every normal x (denormals are
// flushed to zero by -ftz, like in the comparisons of fsign); 1 FMUL2 + 4 FMNMX per pair
__device__ __forceinline__ f2 sign2(f2 x) {
    const f2 y = mul2(x, bc(8.507059173e37f));
    return pk(fminf(fmaxf(lo(y), -1.f), 1.f), fminf(fmaxf(hi(y), -1.f), 1.f));
}
__device__ __forceinline__ float rnorm_eps(float n2) { return (n2 > 1e-24f) ? frsqrt(n2) : 1e12f; }

template <bool kKeepParts>
__device__ __forceinline__ VotePair make_vote2(const Quat2& lrf, const Quat2& traw) {
    VotePair r;
    const f2 n2 = qdot2(traw, traw);
    r.rn = pk(rnorm_eps(lo(n2)), rnorm_eps(hi(n2)));  // x / max(|x|, eps)
    const f2 thw = mul2(traw.w, r.rn);
    if (kKeepParts) {
        r.st = sign2(thw);
        r.that = qscale2(traw, r.rn);
        r.t = qscale2(r.that, r.st);
        const Quat2 h = qham2(lrf, r.t);
        r.sv = sign2(h.w);
        r.vote = qscale2(h, r.sv);
    } else {
        // fix(l (x) fix(that)) == fix(l (x) that) unless that.w == 0 (then the vote is zeroed): a sign
        // flip of t flips the product exactly and the outer hemisphere fix absorbs it
        r.that.w = thw; r.that.x = mul2(traw.x, r.rn); r.that.y = mul2(traw.y, r.rn); r.that.z = mul2(traw.z, r.rn);
        const Quat2 h = qham2(lrf, r.that);
        const f2 s = sign2(h.w);
        r.sv = pk((lo(thw) != 0.f) ? lo(s) : 0.f, (hi(thw) != 0.f) ? hi(s) : 0.f);
        r.vote = qscale2(h, r.sv);
    }
    return r;
}


// acc: M (10), [10] = sum |b|, [11] = sum b.  All 32 lanes of the solving warp hold the same sums,
// so the convergence branch of the low-latency solver is uniform without a vote.
__device__ __forceinline__ Quat solve_pose2(const float* acc, bool valid) {
    const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));  // F.normalize(p=1) of the weights
    Quat p = qfix(top_eigvec_weighted2(m, acc[11] != acc[10], ThreadDone()));
    if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
    return p;
}


struct Bwd2Caps {  // per-capsule state, registers
    Quat ph[kTB], uh[kTB];
    float rS[kTB];
    float gsn;
};

// g = dL/dvote -> dL/dt_raw, dL/dlrf for a pair (scalar twin: vote_bwd in vote_math.cuh)
__device__ __forceinline__ void vote_bwd2(const Quat2& lrf, const VotePair& vp, const Quat2& g, Quat2& g_traw,
                                          Quat2& g_lrf) {
    const Quat2 gh = qscale2(g, vp.sv);
    g_lrf = qham2_conjr(gh, vp.t);                            // d(l (x) t)/dl = g (x) conj(t)
    const Quat2 gt = qscale2(qham2_conjl(lrf, gh), vp.st);    // d(l (x) t)/dt = conj(l) (x) g, through the fix
    // normalisation Jacobian (I - that that^T) / |t_raw|   (identity / eps if clamped)
    const f2 d = qdot2(vp.that, gt);
    const f2 proj = pk((lo(vp.rn) < 1e12f) ? lo(d) : 0.f, (hi(vp.rn) < 1e12f) ? hi(d) : 0.f);
    const f2 np = mul2(proj, bc(-1.f));
    g_traw.w = mul2(fma2(np, vp.that.w, gt.w), vp.rn);
    g_traw.x = mul2(fma2(np, vp.that.x, gt.x), vp.rn);
    g_traw.y = mul2(fma2(np, vp.that.y, gt.y), vp.rn);
    g_traw.z = mul2(fma2(np, vp.that.z, gt.z), vp.rn);
}

// gradient w.r.t. b^TT (gb), running activation gradient (ga), b^t of all levels, <p^TT,v>, <u^TT,v>
template <int T, int TT>
__device__ __forceinline__ void complete_levels2(const Bwd2Caps& C, const Quat2& v, f2 a, f2 gate, f2 om0, f2 om1,
                                                 f2& gb, f2& ga, f2& x_tt, f2& uv_tt, f2 (&b)[kTB]) {
    b[0] = a;
    b[1] = mul2(a, mul2(a, om0));
    b[2] = mul2(b[1], mul2(a, om1));
    gb = bc(0.f);
    ga = bc(0.f);
#pragma unroll
    for (int t = T - 1; t >= TT; --t) {
        const f2 x = qdot2(C.ph[t], v);
        const f2 uv = qdot2(C.uh[t], v);
        f2 direct;
        if (t == T - 1) {
            direct = mul2(bc(C.gsn), one_minus_dist2(x));
        } else {
            const f2 om = (t == 0) ? om0 : om1;        // cached 1 - d^t
            ga = fma2(mul2(gb, b[t]), om, ga);         // b^{t+1} = b^t a (1 - d^t)
            direct = mul2(mul2(gb, a), om);
        }
        // the M-path is gated by [a != 0]
        gb = fma2(mul2(mul2(uv, x), gate), bc(C.rS[t]), direct);
        if (t == TT) {
            x_tt = x;
            uv_tt = uv;
        }
    }
}


}  // namespace qec
