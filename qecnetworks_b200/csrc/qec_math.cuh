// qec_math.cuh -- per-capsule FP32 math of the QEC routing hot path.
//
// Everything here is register-resident straight-line code: quaternion algebra,
// the angular inlier distance, a cyclic Jacobi eigensolver for symmetric 4x4
// matrices (replaces cusolverDnSsyevjBatched, reference
// models/pytorch-cusolver/torch_cusolver.cpp:253-266) and the closed-form
// adjoint of the top eigenvector (replaces batch_symeig_backward, reference
// models/pytorch-autograd-solver/torch_autograd_solver.cpp:133-152).
//
// The functions are __host__ __device__ so that tests/host/ can compile this
// header with g++ and check the arithmetic against float64 numpy on the CPU-only
// build container.  The host build is test scaffolding; the product only ever
// runs the device code.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define QEC_HD __host__ __device__ __forceinline__
#else
#define QEC_HD inline
#endif

namespace qec {

// (w, x, y, z), same memory layout as float4 (reference quat_ops.py:30-48)
struct __attribute__((aligned(16))) Quat {
    float w, x, y, z;
};

QEC_HD float frsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);  // MUFU.RSQ under -ftz=true
#else
    return 1.0f / sqrtf(x);
#endif
}

QEC_HD float fsqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.SQRT, rel. err 2^-23
    return r;
#else
    return sqrtf(x);
#endif
}

QEC_HD float frcp(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}

QEC_HD float fsign(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

QEC_HD Quat qmake(float w, float x, float y, float z) {
    Quat q;
    q.w = w; q.x = x; q.y = y; q.z = z;
    return q;
}
QEC_HD Quat qscale(Quat q, float s) { return qmake(q.w * s, q.x * s, q.y * s, q.z * s); }
QEC_HD Quat qadd(Quat a, Quat b) { return qmake(a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z); }
QEC_HD Quat qfma(float s, Quat a, Quat b) {  // s*a + b
    return qmake(fmaf(s, a.w, b.w), fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z));
}
QEC_HD float qdot(Quat a, Quat b) { return fmaf(a.w, b.w, fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z))); }
QEC_HD Quat qconj(Quat q) { return qmake(q.w, -q.x, -q.y, -q.z); }

// Hamilton product q (x) r  (reference quat_ops.py:30-48)
QEC_HD Quat qham(Quat q, Quat r) {
    return qmake(q.w * r.w - q.x * r.x - q.y * r.y - q.z * r.z,
                 q.w * r.x + q.x * r.w + q.y * r.z - q.z * r.y,
                 q.w * r.y - q.x * r.z + q.y * r.w + q.z * r.x,
                 q.w * r.z + q.x * r.y - q.y * r.x + q.z * r.w);
}

// hemisphere fix: q * sign(q.w), sign(0) = 0 (reference qec_module.py:58-60,...)
QEC_HD Quat qfix(Quat q) { return qscale(q, fsign(q.w)); }

struct Vec3 {
    float x, y, z;
};
QEC_HD Vec3 vmake(float x, float y, float z) {
    Vec3 v;
    v.x = x; v.y = y; v.z = z;
    return v;
}
QEC_HD Vec3 vcross(Vec3 a, Vec3 b) {
    return vmake(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
QEC_HD float vdot(Vec3 a, Vec3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

// v + 2 (w (u x v) + u x (u x v)), u = q.xyz  (reference quat_ops.py:50-68)
QEC_HD Vec3 qrotv(Quat q, Vec3 v) {
    const Vec3 u = vmake(q.x, q.y, q.z);
    const Vec3 uv = vcross(u, v);
    const Vec3 uuv = vcross(u, uv);
    return vmake(v.x + 2.f * fmaf(q.w, uv.x, uuv.x), v.y + 2.f * fmaf(q.w, uv.y, uuv.y),
                 v.z + 2.f * fmaf(q.w, uv.z, uuv.z));
}

// adjoint of qrotv: given g = dL/dx for x = qrotv(q, v) returns dL/dq and dL/dv
QEC_HD void qrotv_bwd(Quat q, Vec3 v, Vec3 g, Quat& gq, Vec3& gv) {
    const Vec3 u = vmake(q.x, q.y, q.z);
    const Vec3 uv = vcross(u, v);
    const Vec3 vg = vcross(v, g);  // <g, u x v> = <u, v x g>
    const float up = vdot(u, v), ug = vdot(u, g), pg = vdot(v, g), uu = vdot(u, u);
    gq.w = 2.f * vdot(g, uv);
    gq.x = 2.f * (q.w * vg.x + up * g.x + ug * v.x - 2.f * pg * u.x);
    gq.y = 2.f * (q.w * vg.y + up * g.y + ug * v.y - 2.f * pg * u.y);
    gq.z = 2.f * (q.w * vg.z + up * g.z + ug * v.z - 2.f * pg * u.z);
    const Vec3 gu = vcross(g, u);  // <g, u x v> = <v, g x u>
    gv.x = g.x + 2.f * (q.w * gu.x + u.x * ug - g.x * uu);
    gv.y = g.y + 2.f * (q.w * gu.y + u.y * ug - g.y * uu);
    gv.z = g.z + 2.f * (q.w * gu.z + u.z * ug - g.z * uu);
}

// ---------------------------------------------------------------------------
// angular inlier distance (reference qec_module.py:156-160)
//   d(x) = (2/pi) acos(min(|x|, 0.9999)),  x = <pose, vote>
// acos(y) = sqrt(1-y) * P7(y) on [0,1] (Abramowitz & Stegun 4.4.46, |err| <= 2e-8),
// with 2/pi folded into the coefficients.  Returns 1 - d.
// ---------------------------------------------------------------------------
#define QEC_DCLAMP 0.9999f

QEC_HD float one_minus_dist(float x) {
    const float y = fminf(fabsf(x), QEC_DCLAMP);
    float p = -0.0008037268f;          // (2/pi) * -0.0012624911
    p = fmaf(p, y, 0.0042463112f);     // (2/pi) *  0.0066700901
    p = fmaf(p, y, -0.0108786386f);    // (2/pi) * -0.0170881256
    p = fmaf(p, y, 0.0196663823f);     // (2/pi) *  0.0308918810
    p = fmaf(p, y, -0.0319419544f);    // (2/pi) * -0.0501743046
    p = fmaf(p, y, 0.0566457827f);     // (2/pi) *  0.0889789874
    p = fmaf(p, y, -0.1366178402f);    // (2/pi) * -0.2145988016
    p = fmaf(p, y, 0.9999999861f);     // (2/pi) *  1.5707963050
    return fmaf(-fsqrt_fast(1.f - y), p, 1.f);
}

// The same polynomial in Estrin form: dependency depth 3 instead of 8 (one more multiply).  The cluster kernels
// (f32x2.cuh: one_minus_dist2) use it because a sweep is a chain dot -> distance -> weight -> accumulate per vote
// pair and the Horner recurrence alone was 8 dependent FFMA2 of it; rounding differs from the Horner form by
// ~1 ulp of the result (tests/test_host_math.py checks both against float64).
QEC_HD float one_minus_dist_estrin(float x) {
    const float y = fminf(fabsf(x), QEC_DCLAMP);
    const float y2 = y * y, y4 = y2 * y2;
    const float a = fmaf(-0.1366178402f, y, 0.9999999861f);
    const float b = fmaf(-0.0319419544f, y, 0.0566457827f);
    const float c = fmaf(-0.0108786386f, y, 0.0196663823f);
    const float d = fmaf(-0.0008037268f, y, 0.0042463112f);
    const float ab = fmaf(b, y2, a), cd = fmaf(d, y2, c);
    const float p = fmaf(cd, y4, ab);
    return fmaf(-fsqrt_fast(1.f - y), p, 1.f);
}

// d/dx of d(x): -(2/pi) sign(x) / sqrt(1 - x^2) where |x| <= 0.9999, else 0
// (torch: abs -> sign(x), clamp(max) passes the gradient where |x| <= max)
QEC_HD float dist_grad(float x) {
    const float ax = fabsf(x);
    const float g = -0.63661977236758f * frsqrt(fmaf(-x, x, 1.f));
    return (ax <= QEC_DCLAMP) ? g * fsign(x) : 0.f;
}

// ---------------------------------------------------------------------------
// symmetric 4x4 accumulator: the 10 unique entries of M = sum w v v^T
// ---------------------------------------------------------------------------
struct Sym4 {
    float a00, a01, a02, a03, a11, a12, a13, a22, a23, a33;
};
QEC_HD Sym4 sym_zero() {
    Sym4 m;
    m.a00 = m.a01 = m.a02 = m.a03 = m.a11 = m.a12 = m.a13 = m.a22 = m.a23 = m.a33 = 0.f;
    return m;
}
// M += w * v v^T   (4 MUL + 10 FMA)
QEC_HD void sym_rank1(Sym4& m, float w, Quat v) {
    const float w0 = w * v.w, w1 = w * v.x, w2 = w * v.y, w3 = w * v.z;
    m.a00 = fmaf(w0, v.w, m.a00); m.a01 = fmaf(w0, v.x, m.a01); m.a02 = fmaf(w0, v.y, m.a02);
    m.a03 = fmaf(w0, v.z, m.a03); m.a11 = fmaf(w1, v.x, m.a11); m.a12 = fmaf(w1, v.y, m.a12);
    m.a13 = fmaf(w1, v.z, m.a13); m.a22 = fmaf(w2, v.y, m.a22); m.a23 = fmaf(w2, v.z, m.a23);
    m.a33 = fmaf(w3, v.z, m.a33);
}
QEC_HD void sym_scale(Sym4& m, float s) {
    m.a00 *= s; m.a01 *= s; m.a02 *= s; m.a03 *= s; m.a11 *= s;
    m.a12 *= s; m.a13 *= s; m.a22 *= s; m.a23 *= s; m.a33 *= s;
}
QEC_HD Quat sym_mul(const Sym4& m, Quat v) {
    return qmake(fmaf(m.a00, v.w, fmaf(m.a01, v.x, fmaf(m.a02, v.y, m.a03 * v.z))),
                 fmaf(m.a01, v.w, fmaf(m.a11, v.x, fmaf(m.a12, v.y, m.a13 * v.z))),
                 fmaf(m.a02, v.w, fmaf(m.a12, v.x, fmaf(m.a22, v.y, m.a23 * v.z))),
                 fmaf(m.a03, v.w, fmaf(m.a13, v.x, fmaf(m.a23, v.y, m.a33 * v.z))));
}

// ---------------------------------------------------------------------------
// Cyclic Jacobi for a symmetric 4x4 held in registers.
// One rotation: 2 MUFU.RSQ + ~35 FP32 ops, branch-free; the sweep order
// (0,1)(2,3)(0,2)(1,3)(0,3)(1,2) pairs disjoint rotations so two angle
// computations are always independent (ILP 2).
// ---------------------------------------------------------------------------
#define QEC_JACOBI_MAX_SWEEPS 12

template <int P, int Q, int R, int S>
QEC_HD void jacobi_rot(float (&A)[4][4], float (&V)[4][4]) {
    // P < Q; R, S are the other two indices.  Only A[i][j] with i <= j is live.
    const float apq = A[P][Q];
    const float d = A[Q][Q] - A[P][P];
    const float o = apq + apq;
    const float h2 = fmaf(d, d, o * o);
    const bool live = h2 > 1e-36f && apq != 0.f;
    const float rh = frsqrt(live ? h2 : 1.f);
    const float cc = fmaf(0.5f * fabsf(d), rh, 0.5f);  // cos^2(phi) in [0.5, 1]
    const float rc = frsqrt(cc);
    const float c = live ? cc * rc : 1.f;
    const float s = live ? copysignf(0.5f, d) * o * rh * rc : 0.f;
    const float t = s * rc;  // tan(phi)
    A[P][P] = fmaf(-t, apq, A[P][P]);
    A[Q][Q] = fmaf(t, apq, A[Q][Q]);
    A[P][Q] = 0.f;
    {
        float& arp = (R < P) ? A[R][P] : A[P][R];
        float& arq = (R < Q) ? A[R][Q] : A[Q][R];
        const float x = arp, y = arq;
        arp = fmaf(c, x, -s * y);
        arq = fmaf(s, x, c * y);
    }
    {
        float& asp = (S < P) ? A[S][P] : A[P][S];
        float& asq = (S < Q) ? A[S][Q] : A[Q][S];
        const float x = asp, y = asq;
        asp = fmaf(c, x, -s * y);
        asq = fmaf(s, x, c * y);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x = V[i][P], y = V[i][Q];
        V[i][P] = fmaf(c, x, -s * y);
        V[i][Q] = fmaf(s, x, c * y);
    }
}

QEC_HD void jacobi_sweep(float (&A)[4][4], float (&V)[4][4]) {
    jacobi_rot<0, 1, 2, 3>(A, V);
    jacobi_rot<2, 3, 0, 1>(A, V);
    jacobi_rot<0, 2, 1, 3>(A, V);
    jacobi_rot<1, 3, 0, 2>(A, V);
    jacobi_rot<0, 3, 1, 2>(A, V);
    jacobi_rot<1, 2, 0, 3>(A, V);
}

QEC_HD bool jacobi_converged(const float (&A)[4][4]) {
    const float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[0][3]) + fabsf(A[1][2]) + fabsf(A[1][3]) +
                      fabsf(A[2][3]);
    const float dia = fabsf(A[0][0]) + fabsf(A[1][1]) + fabsf(A[2][2]) + fabsf(A[3][3]);
    return off <= 1e-9f * dia;
}

QEC_HD void sym_to_array(const Sym4& m, float (&A)[4][4]) {
    A[0][0] = m.a00; A[0][1] = m.a01; A[0][2] = m.a02; A[0][3] = m.a03;
    A[1][1] = m.a11; A[1][2] = m.a12; A[1][3] = m.a13;
    A[2][2] = m.a22; A[2][3] = m.a23; A[3][3] = m.a33;
    A[1][0] = A[2][0] = A[2][1] = A[3][0] = A[3][1] = A[3][2] = 0.f;
}
QEC_HD void eye4(float (&V)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.f : 0.f;
}

// Full decomposition: eigenvalues on the diagonal of A (unsorted), eigenvectors
// in the COLUMNS of V.  all_done(bool) lets the caller make the exit
// warp-uniform (e.g. __all_sync) so lanes do not diverge.  Returns sweeps used.
template <typename AllDone>
QEC_HD int jacobi4(float (&A)[4][4], float (&V)[4][4], AllDone all_done) {
    eye4(V);
    int sweeps = 0;
    for (; sweeps < QEC_JACOBI_MAX_SWEEPS; ++sweeps) {
        if (all_done(jacobi_converged(A))) break;
        jacobi_sweep(A, V);
    }
    return sweeps;
}

struct ThreadDone {
    QEC_HD bool operator()(bool d) const { return d; }
};
#if defined(__CUDACC__)
struct WarpDone {  // every lane of the warp must be converged (no divergence)
    __device__ __forceinline__ bool operator()(bool d) const {
#if defined(__CUDA_ARCH__)
        return __all_sync(0xffffffffu, d);
#else
        return d;
#endif
    }
};
#endif

// Top eigenpair of M (largest eigenvalue; LAST index among exact ties, which is
// what an ascending sort returns).  Output is the raw unit eigenvector, not yet
// sign-fixed.
template <typename AllDone>
QEC_HD Quat top_eigvec(const Sym4& m, float& lambda, AllDone all_done) {
    float A[4][4], V[4][4];
    sym_to_array(m, A);
    jacobi4(A, V, all_done);
    float best = A[0][0];
    Quat v = qmake(V[0][0], V[1][0], V[2][0], V[3][0]);
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        const bool take = A[j][j] >= best;
        best = take ? A[j][j] : best;
        v.w = take ? V[0][j] : v.w;
        v.x = take ? V[1][j] : v.x;
        v.y = take ? V[2][j] : v.y;
        v.z = take ? V[3][j] : v.z;
    }
    // the accumulated rotations use MUFU.RSQ (~2 ulp): renormalise once
    const float rn = frsqrt(qdot(v, v));
    lambda = best;
    return qscale(v, rn);
}

// ---------------------------------------------------------------------------
// Top eigenvector of a symmetric POSITIVE SEMI-DEFINITE 4x4 by repeated squaring.
// M_{k+1} = M_k^2 / trace(M_k^2): after k steps the spectrum is (lambda_i/lambda_3)^(2^k), i.e. the
// matrix collapses onto v3 v3^T doubly-exponentially.  One step is a symmetric 4x4 product
// (40 FMA, dependency depth 4) plus one reciprocal -- ~10x less serial latency than a Jacobi sweep
// sequence, which is what matters when one warp solves while a whole CTA waits.  Rounding at step
// k perturbs the eigenvector by ~eps / gap_k and the gaps grow geometrically, so the accuracy is
// ~2 eps / gap_0, the same as any backward-stable solver.  The loop stops (warp-uniformly) when
// 1 - trace(M_k^2)/trace(M_k)^2 < 2^-22; 26 steps cover relative gaps down to ~1e-6, below that the
// returned vector is a unit vector of the (numerically degenerate) top eigenspace.
// Returns the unit eigenvector (not sign-fixed); the zero matrix gives (0,0,0,1) like the
// ascending-sorted decomposition does.
// ---------------------------------------------------------------------------
#define QEC_POWER_MAX_STEPS 26

template <typename AllDone>
QEC_HD Quat top_eigvec_psd(const Sym4& m0, AllDone all_done) {
    float tr = m0.a00 + m0.a11 + m0.a22 + m0.a33;
    const bool zero = !(tr > 0.f);
    float s = zero ? 0.f : frcp(tr);
    float a00 = m0.a00 * s, a01 = m0.a01 * s, a02 = m0.a02 * s, a03 = m0.a03 * s, a11 = m0.a11 * s,
          a12 = m0.a12 * s, a13 = m0.a13 * s, a22 = m0.a22 * s, a23 = m0.a23 * s, a33 = m0.a33 * s;
    for (int k = 0; k < QEC_POWER_MAX_STEPS; ++k) {
        // B = A^2 (symmetric)
        const float b00 = fmaf(a00, a00, fmaf(a01, a01, fmaf(a02, a02, a03 * a03)));
        const float b01 = fmaf(a00, a01, fmaf(a01, a11, fmaf(a02, a12, a03 * a13)));
        const float b02 = fmaf(a00, a02, fmaf(a01, a12, fmaf(a02, a22, a03 * a23)));
        const float b03 = fmaf(a00, a03, fmaf(a01, a13, fmaf(a02, a23, a03 * a33)));
        const float b11 = fmaf(a01, a01, fmaf(a11, a11, fmaf(a12, a12, a13 * a13)));
        const float b12 = fmaf(a01, a02, fmaf(a11, a12, fmaf(a12, a22, a13 * a23)));
        const float b13 = fmaf(a01, a03, fmaf(a11, a13, fmaf(a12, a23, a13 * a33)));
        const float b22 = fmaf(a02, a02, fmaf(a12, a12, fmaf(a22, a22, a23 * a23)));
        const float b23 = fmaf(a02, a03, fmaf(a12, a13, fmaf(a22, a23, a23 * a33)));
        const float b33 = fmaf(a03, a03, fmaf(a13, a13, fmaf(a23, a23, a33 * a33)));
        const float t2 = (b00 + b11) + (b22 + b33);  // = sum lambda_i^2 of A (trace(A) = 1), in (1/4, 1]
        const bool done = zero || (1.f - t2 < 2.4e-7f);
        const float r = zero ? 0.f : frcp(t2);
        a00 = b00 * r; a01 = b01 * r; a02 = b02 * r; a03 = b03 * r; a11 = b11 * r;
        a12 = b12 * r; a13 = b13 * r; a22 = b22 * r; a23 = b23 * r; a33 = b33 * r;
        if (all_done(done)) break;
    }
    // A ~ v v^T: the column through the largest diagonal entry is the best-conditioned one
    Quat v = qmake(a00, a01, a02, a03);
    float best = a00;
    if (a11 > best) { best = a11; v = qmake(a01, a11, a12, a13); }
    if (a22 > best) { best = a22; v = qmake(a02, a12, a22, a23); }
    if (a33 > best) { best = a33; v = qmake(a03, a13, a23, a33); }
    const float n2 = qdot(v, v);
    return (zero || !(n2 > 0.f)) ? qmake(0.f, 0.f, 0.f, 1.f) : qscale(v, frsqrt(n2));
}

// 2^-floor(log2 x) for a normal positive x (x * result lies in [1, 2)): an exact power-of-two
// rescale costs one integer subtract instead of a reciprocal
QEC_HD float pow2_rescale(float x) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(0x7f000000 - (__float_as_int(x) & 0x7f800000));
#else
    int e;
    frexpf(x, &e);
    return ldexpf(1.f, 1 - e);
#endif
}

// Same iteration as top_eigvec_psd with the serial latency cut for the kernels where one warp
// solves while a whole cluster waits (routing_fused2.cu): two squarings per round, rescaled by an
// exact power of two (no reciprocal; between rescales the top eigenvalue stays within
// [2^-8, 2^4]), one convergence test per round on the scale-free ratio
//   sum_i lambda_i^2 >= (1 - 1e-5) (sum_i lambda_i)^2   of the once-squared matrix B;
// the vector is extracted from B^2, whose second-to-first eigenvalue ratio is then below 3e-11.
#define QEC_POWER2_MAX_ROUNDS 13

// Gershgorin lower bound of the smallest eigenvalue, clamped at 0.  Subtracting it from the
// diagonal of a PSD matrix keeps it PSD (the top eigenvector stays the dominant one) and turns
// the ratio lambda_2/lambda_1 into (lambda_2 - s)/(lambda_1 - s): when the spectrum is clustered --
// votes spread all over the sphere, the slow case of the squaring iteration -- the matrix is
// close to a multiple of the identity, the bound is tight and several squarings are saved.
QEC_HD float gershgorin_shift(float a00, float a01, float a02, float a03, float a11, float a12, float a13, float a22,
                              float a23, float a33) {
    const float x01 = fabsf(a01), x02 = fabsf(a02), x03 = fabsf(a03), x12 = fabsf(a12), x13 = fabsf(a13),
                x23 = fabsf(a23);
    const float g0 = a00 - ((x01 + x02) + x03), g1 = a11 - ((x01 + x12) + x13), g2 = a22 - ((x02 + x12) + x23),
                g3 = a33 - ((x03 + x13) + x23);
    return fmaxf(fminf(fminf(g0, g1), fminf(g2, g3)), 0.f);
}

template <typename AllDone>
QEC_HD Quat top_eigvec_psd2(const Sym4& m0, AllDone all_done) {
    const float tr = m0.a00 + m0.a11 + m0.a22 + m0.a33;
    const bool zero = !(tr > 0.f);
    float s = zero ? 0.f : pow2_rescale(tr);
    float a00 = m0.a00 * s, a01 = m0.a01 * s, a02 = m0.a02 * s, a03 = m0.a03 * s, a11 = m0.a11 * s,
          a12 = m0.a12 * s, a13 = m0.a13 * s, a22 = m0.a22 * s, a23 = m0.a23 * s, a33 = m0.a33 * s;
    for (int k = 0; k < QEC_POWER2_MAX_ROUNDS; ++k) {
        {
            const float sh = gershgorin_shift(a00, a01, a02, a03, a11, a12, a13, a22, a23, a33);
            a00 -= sh; a11 -= sh; a22 -= sh; a33 -= sh;
        }
        // B = A^2
        const float b00 = fmaf(a00, a00, fmaf(a01, a01, fmaf(a02, a02, a03 * a03)));
        const float b01 = fmaf(a00, a01, fmaf(a01, a11, fmaf(a02, a12, a03 * a13)));
        const float b02 = fmaf(a00, a02, fmaf(a01, a12, fmaf(a02, a22, a03 * a23)));
        const float b03 = fmaf(a00, a03, fmaf(a01, a13, fmaf(a02, a23, a03 * a33)));
        const float b11 = fmaf(a01, a01, fmaf(a11, a11, fmaf(a12, a12, a13 * a13)));
        const float b12 = fmaf(a01, a02, fmaf(a11, a12, fmaf(a12, a22, a13 * a23)));
        const float b13 = fmaf(a01, a03, fmaf(a11, a13, fmaf(a12, a23, a13 * a33)));
        const float b22 = fmaf(a02, a02, fmaf(a12, a12, fmaf(a22, a22, a23 * a23)));
        const float b23 = fmaf(a02, a03, fmaf(a12, a13, fmaf(a22, a23, a23 * a33)));
        const float b33 = fmaf(a03, a03, fmaf(a13, a13, fmaf(a23, a23, a33 * a33)));
        const float t2 = (b00 + b11) + (b22 + b33);
        // C = B^2
        const float c00 = fmaf(b00, b00, fmaf(b01, b01, fmaf(b02, b02, b03 * b03)));
        const float c01 = fmaf(b00, b01, fmaf(b01, b11, fmaf(b02, b12, b03 * b13)));
        const float c02 = fmaf(b00, b02, fmaf(b01, b12, fmaf(b02, b22, b03 * b23)));
        const float c03 = fmaf(b00, b03, fmaf(b01, b13, fmaf(b02, b23, b03 * b33)));
        const float c11 = fmaf(b01, b01, fmaf(b11, b11, fmaf(b12, b12, b13 * b13)));
        const float c12 = fmaf(b01, b02, fmaf(b11, b12, fmaf(b12, b22, b13 * b23)));
        const float c13 = fmaf(b01, b03, fmaf(b11, b13, fmaf(b12, b23, b13 * b33)));
        const float c22 = fmaf(b02, b02, fmaf(b12, b12, fmaf(b22, b22, b23 * b23)));
        const float c23 = fmaf(b02, b03, fmaf(b12, b13, fmaf(b22, b23, b23 * b33)));
        const float c33 = fmaf(b03, b03, fmaf(b13, b13, fmaf(b23, b23, b33 * b33)));
        const float t4 = (c00 + c11) + (c22 + c33);  // = sum lambda_B^2
        const bool done = zero || (fmaf(-0.99999f * t2, t2, t4) >= 0.f);
        const float r = zero ? 0.f : pow2_rescale(t4);
        a00 = c00 * r; a01 = c01 * r; a02 = c02 * r; a03 = c03 * r; a11 = c11 * r;
        a12 = c12 * r; a13 = c13 * r; a22 = c22 * r; a23 = c23 * r; a33 = c33 * r;
        if (all_done(done)) break;
    }
    Quat v = qmake(a00, a01, a02, a03);
    float best = a00;
    if (a11 > best) { best = a11; v = qmake(a01, a11, a12, a13); }
    if (a22 > best) { best = a22; v = qmake(a02, a12, a22, a23); }
    if (a33 > best) { best = a33; v = qmake(a03, a13, a23, a33); }
    const float n2 = qdot(v, v);
    return (zero || !(n2 > 0.f)) ? qmake(0.f, 0.f, 0.f, 1.f) : qscale(v, frsqrt(n2));
}

// Top eigenvector of M = sum_j w_j v_j v_j^T / sum_j |w_j|  (so every eigenvalue lies in [-1, 1]).
// Non-negative routing weights -- the reference's domain: 0/1 masks and sigmoid outputs -- make M
// positive semi-definite and the squaring iteration applies directly.  If a weight was negative
// the spectrum is shifted by +1 first, so that the LARGEST eigenvalue wins, not the largest
// magnitude (same eigenvectors, slower convergence; never taken on the reference's inputs).
template <typename AllDone>
QEC_HD Quat top_eigvec_weighted(Sym4 m, bool negative, AllDone all_done) {
    const float sh = negative ? 1.f : 0.f;
    m.a00 += sh; m.a11 += sh; m.a22 += sh; m.a33 += sh;
    return top_eigvec_psd(m, all_done);
}
template <typename AllDone>
QEC_HD Quat top_eigvec_weighted2(Sym4 m, bool negative, AllDone all_done) {
    const float sh = negative ? 1.f : 0.f;
    m.a00 += sh; m.a11 += sh; m.a22 += sh; m.a33 += sh;
    return top_eigvec_psd2(m, all_done);
}

// ---------------------------------------------------------------------------
// Adjoint of the top eigenvector.
// For p = fix(top eigenvector of M), lambda its eigenvalue and g = dL/dp:
//     u = (lambda I - M)^+ g        (u is orthogonal to p)
// and then dL/dM = u p^T.  Instead of the reference's full eigen-decomposition
// (V (F o V^T gV) V^T) the pseudo-inverse is applied by a 4x4 Cholesky solve of
//     C = lambda I - M + kappa p p^T,   C u = g - p <p, g>,
// which is SPD whenever the top eigenvalue is simple; kappa = trace(lambda I -
// M)/3 keeps the added direction at the scale of the other three.
// ---------------------------------------------------------------------------
QEC_HD Quat top_eigvec_adjoint(const Sym4& m, Quat p, float lambda, Quat g) {
    const float pg = qdot(p, g);
    const Quat r = qfma(-pg, p, g);
    const float kappa = fmaxf((4.f * lambda - (m.a00 + m.a11 + m.a22 + m.a33)) * (1.f / 3.f), 1e-30f);
    const float kw = kappa * p.w, kx = kappa * p.x, ky = kappa * p.y, kz = kappa * p.z;
    const float c00 = fmaf(kw, p.w, lambda - m.a00), c01 = fmaf(kw, p.x, -m.a01), c02 = fmaf(kw, p.y, -m.a02),
                c03 = fmaf(kw, p.z, -m.a03), c11 = fmaf(kx, p.x, lambda - m.a11), c12 = fmaf(kx, p.y, -m.a12),
                c13 = fmaf(kx, p.z, -m.a13), c22 = fmaf(ky, p.y, lambda - m.a22), c23 = fmaf(ky, p.z, -m.a23),
                c33 = fmaf(kz, p.z, lambda - m.a33);
    const float tiny = 1e-30f;
    // C = L L^T
    const float i0 = frsqrt(fmaxf(c00, tiny));
    const float l10 = c01 * i0, l20 = c02 * i0, l30 = c03 * i0;
    const float i1 = frsqrt(fmaxf(fmaf(-l10, l10, c11), tiny));
    const float l21 = fmaf(-l20, l10, c12) * i1, l31 = fmaf(-l30, l10, c13) * i1;
    const float i2 = frsqrt(fmaxf(fmaf(-l21, l21, fmaf(-l20, l20, c22)), tiny));
    const float l32 = fmaf(-l31, l21, fmaf(-l30, l20, c23)) * i2;
    const float i3 = frsqrt(fmaxf(fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, c33))), tiny));
    // L y = r   (diagonal of L is 1/i_k)
    const float y0 = r.w * i0;
    const float y1 = fmaf(-l10, y0, r.x) * i1;
    const float y2 = fmaf(-l21, y1, fmaf(-l20, y0, r.y)) * i2;
    const float y3 = fmaf(-l32, y2, fmaf(-l31, y1, fmaf(-l30, y0, r.z))) * i3;
    // L^T u = y
    const float u3 = y3 * i3;
    const float u2 = fmaf(-l32, u3, y2) * i2;
    const float u1 = fmaf(-l31, u3, fmaf(-l21, u2, y1)) * i1;
    const float u0 = fmaf(-l30, u3, fmaf(-l20, u2, fmaf(-l10, u1, y0))) * i0;
    Quat u = qmake(u0, u1, u2, u3);
    // remove the rounding-level component along p
    return qfma(-qdot(u, p), p, u);
}

QEC_HD float sigmoidf(float z) { return 1.f / (1.f + expf(-z)); }

}  // namespace qec
