// routing_core.cuh -- the Weiszfeld/IRLS routing loop of one output capsule,
// forward and backward, written once and instantiated for every cooperation
// policy (thread / sub-warp group / CTA per capsule) and every vote source
// (materialised votes, or votes recomputed on the fly from the transform
// quaternions).
//
// Reference semantics: models/qec_module.py:167-194 (forward loop + activation
// epilogue), :64-98 (weighted quaternion mean), :156-160 (distance); SURVEY.md
// Appendix A (forward) and Appendix B (closed-form eigenvector adjoint).
//
// The sweeps are STATELESS per vote: the routing weight b^t of a vote is
// recomputed from the T saved poses (b^{t+1} = b^t a (1 - d(pose^t, v))), so no
// per-vote state is stored between sweeps -- the only per-capsule state is the
// T poses (and, in the backward, the T adjoint vectors u^t).  That is what makes
// the same code run with the votes in registers, in L1/L2, or streamed.
#pragma once

#include "qec_common.cuh"

namespace qec {

template <int N>
__device__ __forceinline__ void zero_acc(float (&a)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = 0.f;
}

__device__ __forceinline__ void acc_rank1(float* m, float w, Quat v) {
    const float w0 = w * v.w, w1 = w * v.x, w2 = w * v.y, w3 = w * v.z;
    m[0] = fmaf(w0, v.w, m[0]); m[1] = fmaf(w0, v.x, m[1]); m[2] = fmaf(w0, v.y, m[2]);
    m[3] = fmaf(w0, v.z, m[3]); m[4] = fmaf(w1, v.x, m[4]); m[5] = fmaf(w1, v.y, m[5]);
    m[6] = fmaf(w1, v.z, m[6]); m[7] = fmaf(w2, v.y, m[7]); m[8] = fmaf(w2, v.z, m[8]);
    m[9] = fmaf(w3, v.z, m[9]);
}
__device__ __forceinline__ Sym4 acc_to_sym(const float* m, float scale) {
    Sym4 s;
    s.a00 = m[0] * scale; s.a01 = m[1] * scale; s.a02 = m[2] * scale; s.a03 = m[3] * scale;
    s.a11 = m[4] * scale; s.a12 = m[5] * scale; s.a13 = m[6] * scale; s.a22 = m[7] * scale;
    s.a23 = m[8] * scale; s.a33 = m[9] * scale;
    return s;
}

// ---------------------------------------------------------------------------
// forward: T sweeps (M^t -> pose^t) + one epilogue sweep (inlier score).
//   ph[t]   <- pose^t (all lanes)
//   returns agreement; s_out / n_out are the mean inlier score and the active
//   vote count, saved for the backward.
// ---------------------------------------------------------------------------
template <int TMAX, class Coop, class Votes>
__device__ __forceinline__ float routing_forward_capsule(const Coop& coop, const Votes& votes,
                                                         const float* __restrict__ act_row, int J, int T, float alpha,
                                                         float beta, Quat (&ph)[TMAX], float& s_out, float& n_out) {
    bool valid = false;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        if (t < T) {
            float acc[13];  // M (10), S = sum |b|, sum of all vote components, sum b
            zero_acc(acc);
            for (int j = coop.first(); j < J; j += coop.stride()) {
                const Quat v = votes.get(j);
                const float a = __ldg(act_row + j);
                float b = a;  // b^0 = a; mask = [a != 0] is implied (b = 0 wherever a = 0)
#pragma unroll
                for (int u = 0; u < TMAX; ++u)
                    if (u < t) b *= a * one_minus_dist(qdot(ph[u], v));
                acc[10] += fabsf(b);
                acc[12] += b;
                acc_rank1(acc, b, v);
                if (t == 0) acc[11] += (v.w + v.x) + (v.y + v.z);
            }
            coop.reduce(acc);
            float r[4] = {0.f, 0.f, 0.f, 0.f};
            if (coop.solver()) {
                if (t == 0) valid = acc[11] != 0.f;  // reference qec_module.py:72-73
                const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));  // F.normalize(p=1)
                // sum b == sum |b| bit-exactly (same values, same order) unless a weight is negative
                Quat p = qfix(top_eigvec_weighted(m, acc[12] != acc[10], WarpDone()));
                if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
                r[0] = p.w; r[1] = p.x; r[2] = p.y; r[3] = p.z;
            }
            coop.bcast(r);
            ph[t] = qmake(r[0], r[1], r[2], r[3]);
        }
    }
    // epilogue sweep (reference qec_module.py:188-193)
    Quat pl = ph[0];
#pragma unroll
    for (int u = 1; u < TMAX; ++u)
        if (u == T - 1) pl = ph[u];
    float acc2[2] = {0.f, 0.f};
    for (int j = coop.first(); j < J; j += coop.stride()) {
        const Quat v = votes.get(j);
        const float a = __ldg(act_row + j);
        float b = a;
#pragma unroll
        for (int u = 0; u < TMAX; ++u)
            if (u < T - 1) b *= a * one_minus_dist(qdot(ph[u], v));
        acc2[0] = fmaf(b, one_minus_dist(qdot(pl, v)), acc2[0]);
        acc2[1] += (b != 0.f) ? 1.f : 0.f;
    }
    coop.reduce(acc2);
    float r[3] = {0.f, 0.f, 0.f};
    if (coop.solver()) {
        const float n = acc2[1];
        const float s = (n > 0.f) ? acc2[0] / n : 0.f;
        r[0] = (n > 0.f) ? sigmoidf(fmaf(alpha, s, beta - 1.f)) : 0.f;
        r[1] = s;
        r[2] = n;
    }
    coop.bcast(r);
    s_out = r[1];
    n_out = r[2];
    return r[0];
}

// ---------------------------------------------------------------------------
// backward.  Levels t = T-1 .. 0; sweep k accumulates level L = T-1-k:
//   G_p^L = sum_j g_d^L_j D'(x^L_j) v_j,  M^L = sum_j b^L_j v_j v_j^T,  S^L
// then u^L = (lambda I - M^L/S^L)^+ (G_p^L [+ g_pose if L = T-1]).  Once u^t is
// known the gradient w.r.t. b^t_j is complete:
//   g_b^t_j = direct + <u^t, v_j> <p^t, v_j> / S^t
// The vote gradient (only the last level sees non-detached votes) is written in
// sweep 1; the activation gradient in the last sweep.  If the activation needs
// no gradient (need_chain = false) only level T-1 is processed (2 sweeps).
//
//   Sink::put(j, g)     receives dL/dv_j
//   g_act_row           nullable; accumulated with atomicAdd (it is shared by
//                       the O capsules of a patch)
// ---------------------------------------------------------------------------
template <int TMAX, class Coop, class Votes, class Sink>
__device__ __forceinline__ void routing_backward_capsule(const Coop& coop, const Votes& votes, Sink& sink,
                                                         const float* __restrict__ act_row, float* g_act_row, int J,
                                                         int T, float alpha, float beta, const Quat (&ph)[TMAX],
                                                         float s, float n, Quat g_pose, float g_agr, bool need_chain,
                                                         float& g_alpha, float& g_beta) {
    const bool has = n > 0.f;
    const float sg = sigmoidf(fmaf(alpha, s, beta - 1.f));
    const float gz = has ? g_agr * sg * (1.f - sg) : 0.f;
    g_alpha = gz * s;
    g_beta = gz;
    const float gsn = has ? gz * alpha / n : 0.f;  // dL/d(sum_j b_j (1 - d_j))

    Quat uh[TMAX];
    float rS[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
        uh[t] = qmake(0.f, 0.f, 0.f, 0.f);
        rS[t] = 0.f;
    }
    const int nsweeps = need_chain ? T + 1 : 2;
    for (int k = 0; k < nsweeps; ++k) {
        const int L = T - 1 - k;
        const bool accumulate = (L >= 0) && (need_chain || k == 0);
        const bool write_gv = (k == 1);
        const bool write_ga = need_chain && (k == nsweeps - 1) && (g_act_row != nullptr);
        float acc[15];  // G_p (4), M (10), S
        zero_acc(acc);
        for (int j = coop.first(); j < J; j += coop.stride()) {
            const Quat v = votes.get(j);
            const float a = __ldg(act_row + j);
            float x[TMAX], om[TMAX], b[TMAX];
            {
                float bb = a;
#pragma unroll
                for (int t = 0; t < TMAX; ++t) {
                    b[t] = bb;
                    x[t] = qdot(ph[t], v);
                    om[t] = one_minus_dist(x[t]);
                    bb *= a * om[t];
                }
            }
            float gb = 0.f;  // dL/db^{t+1}_j while walking down
            float ga = 0.f;
#pragma unroll
            for (int t = TMAX - 1; t >= 0; --t) {
                if (t < T && t >= L) {
                    const bool top = (t == T - 1);
                    const float gd = top ? -gsn * b[t] : -gb * b[t] * a;  // dL/dd^t_j
                    if (!top) ga = fmaf(gb * b[t], om[t], ga);            // b^{t+1} = b^t a (1 - d^t)
                    const float direct = top ? gsn * om[t] : gb * a * om[t];
                    if (t > L) {
                        const float uv = qdot(uh[t], v);
                        // w = b * [a != 0]: the mask carries no gradient but gates the M-path
                        gb = (a != 0.f) ? fmaf(uv * x[t], rS[t], direct) : direct;
                        if (top && write_gv) {
                            const float gx = gd * dist_grad(x[t]);
                            const float wh = b[t] * rS[t];
                            Quat g = qscale(ph[t], fmaf(wh, uv, gx));  // p (gx + w <u,v>)
                            g = qfma(wh * x[t], uh[t], g);             // + u w <p,v>
                            sink.put(j, g);
                        }
                    } else if (accumulate) {  // t == L
                        const float gx = gd * dist_grad(x[t]);
                        acc[0] = fmaf(gx, v.w, acc[0]); acc[1] = fmaf(gx, v.x, acc[1]);
                        acc[2] = fmaf(gx, v.y, acc[2]); acc[3] = fmaf(gx, v.z, acc[3]);
                        acc_rank1(acc + 4, b[t], v);
                        acc[14] += fabsf(b[t]);
                    }
                }
            }
            if (write_ga) atomicAdd(g_act_row + j, ga + gb);  // + dL/db^0_j (b^0 = a)
        }
        if (accumulate) {
            coop.reduce(acc);
            float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (coop.solver()) {
                Quat p = ph[0];
#pragma unroll
                for (int t = 1; t < TMAX; ++t)
                    if (t == L) p = ph[t];
                const float rs = 1.f / fmaxf(acc[14], 1e-12f);
                const Sym4 m = acc_to_sym(acc + 4, rs);
                Quat gp = qmake(acc[0], acc[1], acc[2], acc[3]);
                if (L == T - 1) gp = qadd(gp, g_pose);
                const float lam = qdot(p, sym_mul(m, p));
                Quat u = top_eigvec_adjoint(m, p, lam, gp);
                if (!(qdot(p, p) > 0.5f)) u = qmake(0.f, 0.f, 0.f, 0.f);  // invalid / sign-zeroed pose
                r[0] = u.w; r[1] = u.x; r[2] = u.y; r[3] = u.z; r[4] = rs;
            }
            coop.bcast(r);
#pragma unroll
            for (int t = 0; t < TMAX; ++t)
                if (t == L) {
                    uh[t] = qmake(r[0], r[1], r[2], r[3]);
                    rS[t] = r[4];
                }
        }
    }
}

}  // namespace qec
