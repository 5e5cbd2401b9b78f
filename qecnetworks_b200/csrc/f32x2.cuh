// f32x2.cuh -- packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2).
//
// The fused routing kernels are bound by instruction issue, not by the fma pipe or HBM
// (profiles/: issue slots ~45 % busy, fma pipe ~27 %, DRAM ~12 %).  sm_100 executes
// fma/mul/add on two packed FP32 values held in an aligned 64-bit register pair with one
// instruction, and ptxas folds a pair built from one scalar (bc(s)) into the `.F32`
// broadcast operand form, so "vote A in the low half, vote B in the high half" halves the
// FP32 instruction count of a sweep without any shuffling.  Each half is an ordinary IEEE
// fma.rn: results are bit-identical per vote to the scalar code in qec_math.cuh.
#pragma once

#include "qec_math.cuh"

namespace qec {

typedef unsigned long long f2;  // (lo, hi) packed floats

__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc(float s) { return pk(s, s); }
__device__ __forceinline__ float lo(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {  // a * b + c
    f2 d;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float hsum(f2 v) { return lo(v) + hi(v); }

// two quaternions, component-wise packed: w = (A.w, B.w), ...
struct Quat2 {
    f2 w, x, y, z;
};
__device__ __forceinline__ Quat lo(const Quat2& q) { return qmake(lo(q.w), lo(q.x), lo(q.y), lo(q.z)); }
__device__ __forceinline__ Quat hi(const Quat2& q) { return qmake(hi(q.w), hi(q.x), hi(q.y), hi(q.z)); }
__device__ __forceinline__ Quat2 pk(Quat a, Quat b) {
    Quat2 r;
    r.w = pk(a.w, b.w); r.x = pk(a.x, b.x); r.y = pk(a.y, b.y); r.z = pk(a.z, b.z);
    return r;
}
__device__ __forceinline__ Quat2 qscale2(const Quat2& q, f2 s) {
    Quat2 r;
    r.w = mul2(q.w, s); r.x = mul2(q.x, s); r.y = mul2(q.y, s); r.z = mul2(q.z, s);
    return r;
}
// <p, v> for one scalar quaternion p against both halves (same evaluation order as qdot)
__device__ __forceinline__ f2 qdot2(Quat p, const Quat2& v) {
    f2 d = mul2(bc(p.z), v.z);
    d = fma2(bc(p.y), v.y, d);
    d = fma2(bc(p.x), v.x, d);
    return fma2(bc(p.w), v.w, d);
}
__device__ __forceinline__ f2 qdot2(const Quat2& a, const Quat2& b) {
    f2 d = mul2(a.z, b.z);
    d = fma2(a.y, b.y, d);
    d = fma2(a.x, b.x, d);
    return fma2(a.w, b.w, d);
}
// s * a + b with a scalar quaternion a
__device__ __forceinline__ Quat2 qfma2(f2 s, Quat a, const Quat2& b) {
    Quat2 r;
    r.w = fma2(s, bc(a.w), b.w); r.x = fma2(s, bc(a.x), b.x); r.y = fma2(s, bc(a.y), b.y); r.z = fma2(s, bc(a.z), b.z);
    return r;
}
__device__ __forceinline__ Quat2 qscale2(Quat a, f2 s) {
    Quat2 r;
    r.w = mul2(s, bc(a.w)); r.x = mul2(s, bc(a.x)); r.y = mul2(s, bc(a.y)); r.z = mul2(s, bc(a.z));
    return r;
}

// Hamilton product of two packed pairs (same expression order as qham)
__device__ __forceinline__ Quat2 qham2(const Quat2& q, const Quat2& r) {
    const f2 m1 = bc(-1.f);
    const f2 nx = mul2(q.x, m1), ny = mul2(q.y, m1), nz = mul2(q.z, m1);
    Quat2 o;
    // ((q.w r.w - q.x r.x) - q.y r.y) - q.z r.z   evaluated as nested fma like the compiler contracts qham
    o.w = fma2(nz, r.z, fma2(ny, r.y, fma2(nx, r.x, mul2(q.w, r.w))));
    o.x = fma2(nz, r.y, fma2(q.y, r.z, fma2(q.x, r.w, mul2(q.w, r.x))));
    o.y = fma2(q.z, r.x, fma2(q.y, r.w, fma2(nx, r.z, mul2(q.w, r.y))));
    o.z = fma2(q.z, r.w, fma2(ny, r.x, fma2(q.x, r.y, mul2(q.w, r.z))));
    return o;
}

// q (x) conj(r) and conj(q) (x) r for packed pairs (the two adjoint Hamilton products of vote_bwd)
__device__ __forceinline__ Quat2 qham2_conjr(const Quat2& q, const Quat2& r) {
    const f2 m1 = bc(-1.f);
    const f2 nw = mul2(q.w, m1), nx = mul2(q.x, m1), ny = mul2(q.y, m1), nz = mul2(q.z, m1);
    Quat2 o;
    o.w = fma2(q.z, r.z, fma2(q.y, r.y, fma2(q.x, r.x, mul2(q.w, r.w))));
    o.x = fma2(q.z, r.y, fma2(ny, r.z, fma2(q.x, r.w, mul2(nw, r.x))));
    o.y = fma2(nz, r.x, fma2(q.y, r.w, fma2(q.x, r.z, mul2(nw, r.y))));
    o.z = fma2(q.z, r.w, fma2(q.y, r.x, fma2(nx, r.y, mul2(nw, r.z))));
    return o;
}
__device__ __forceinline__ Quat2 qham2_conjl(const Quat2& q, const Quat2& r) {
    const f2 m1 = bc(-1.f);
    const f2 nx = mul2(q.x, m1), ny = mul2(q.y, m1), nz = mul2(q.z, m1);
    Quat2 o;
    o.w = fma2(q.z, r.z, fma2(q.y, r.y, fma2(q.x, r.x, mul2(q.w, r.w))));
    o.x = fma2(q.z, r.y, fma2(ny, r.z, fma2(nx, r.w, mul2(q.w, r.x))));
    o.y = fma2(nz, r.x, fma2(ny, r.w, fma2(q.x, r.z, mul2(q.w, r.y))));
    o.z = fma2(nz, r.w, fma2(q.y, r.x, fma2(nx, r.y, mul2(q.w, r.z))));
    return o;
}

// 1 - d(x) for both halves.  Default (QEC_ESTRIN = 1): the Estrin form of the polynomial, dependency depth 3 instead
// of 8 -- bit-identical per half to one_minus_dist_estrin (qec_math.cuh).  QEC_ESTRIN = 0: the Horner form,
// bit-identical per half to one_minus_dist.  Negated coefficients so the final fma needs no negation.
#ifndef QEC_ESTRIN
#define QEC_ESTRIN 0
#endif
__device__ __forceinline__ f2 one_minus_dist2(f2 x) {
    const f2 y = pk(fminf(fabsf(lo(x)), QEC_DCLAMP), fminf(fabsf(hi(x)), QEC_DCLAMP));
#if QEC_ESTRIN
    const f2 y2 = mul2(y, y), y4 = mul2(y2, y2);
    const f2 a = fma2(bc(0.1366178402f), y, bc(-0.9999999861f));
    const f2 b = fma2(bc(0.0319419544f), y, bc(-0.0566457827f));
    const f2 c = fma2(bc(0.0108786386f), y, bc(-0.0196663823f));
    const f2 d = fma2(bc(0.0008037268f), y, bc(-0.0042463112f));
    const f2 ab = fma2(b, y2, a), cd = fma2(d, y2, c);
    const f2 pe = fma2(cd, y4, ab);
    const f2 ome = fma2(y, bc(-1.f), bc(1.f));
    const f2 se = pk(fsqrt_fast(lo(ome)), fsqrt_fast(hi(ome)));
    return fma2(se, pe, bc(1.f));
#endif
    f2 p = fma2(bc(0.0008037268f), y, bc(-0.0042463112f));
    p = fma2(p, y, bc(0.0108786386f));
    p = fma2(p, y, bc(-0.0196663823f));
    p = fma2(p, y, bc(0.0319419544f));
    p = fma2(p, y, bc(-0.0566457827f));
    p = fma2(p, y, bc(0.1366178402f));
    p = fma2(p, y, bc(-0.9999999861f));
    const f2 om = fma2(y, bc(-1.f), bc(1.f));
    const f2 s = pk(fsqrt_fast(lo(om)), fsqrt_fast(hi(om)));
    return fma2(s, p, bc(1.f));
}
// d/dx of d(x) for both halves: -(2/pi) sign(x) / sqrt(1 - x^2) where |x| <= 0.9999, else 0
// (scalar twin: dist_grad).  sign(x) = clamp(x * 2^126, -1, 1) is exact for normal x.
__device__ __forceinline__ f2 dist_grad2(f2 x) {
    const f2 om = fma2(mul2(x, bc(-1.f)), x, bc(1.f));
    const f2 g = mul2(pk(frsqrt(lo(om)), frsqrt(hi(om))), bc(-0.63661977236758f));
    const f2 y = mul2(x, bc(8.507059173e37f));
    const f2 r = mul2(g, pk(fminf(fmaxf(lo(y), -1.f), 1.f), fminf(fmaxf(hi(y), -1.f), 1.f)));
    // select, not multiply-by-zero: |x| = 1 exactly (a capsule with one live vote) makes g infinite
    return pk((fabsf(lo(x)) <= QEC_DCLAMP) ? lo(r) : 0.f, (fabsf(hi(x)) <= QEC_DCLAMP) ? hi(r) : 0.f);
}

// M += w v v^T for both halves: 4 FMUL2 + 10 FFMA2.  m[0..9] in acc_rank1 order.
__device__ __forceinline__ void acc_rank1_2(f2* m, f2 w, const Quat2& v) {
    const f2 w0 = mul2(w, v.w), w1 = mul2(w, v.x), w2 = mul2(w, v.y), w3 = mul2(w, v.z);
    m[0] = fma2(w0, v.w, m[0]); m[1] = fma2(w0, v.x, m[1]); m[2] = fma2(w0, v.y, m[2]);
    m[3] = fma2(w0, v.z, m[3]); m[4] = fma2(w1, v.x, m[4]); m[5] = fma2(w1, v.y, m[5]);
    m[6] = fma2(w1, v.z, m[6]); m[7] = fma2(w2, v.y, m[7]); m[8] = fma2(w2, v.z, m[8]);
    m[9] = fma2(w3, v.z, m[9]);
}

}  // namespace qec
