// vote_math.cuh -- one vote and its adjoint (reference models/qec_module.py:129-145):
//   t    = fix(normalise(t_raw))          F.normalize(p=2, eps=1e-12) then * sign(w)
//   vote = fix(lrf (x) t)                 quat_ops.qmul then * sign(w)
#pragma once

#include "qec_common.cuh"

namespace qec {

struct VoteParts {
    Quat that;   // normalised transform (before the hemisphere fix)
    float rn;    // 1 / max(|t_raw|, 1e-12)
    float st;    // sign(that.w)
    float sv;    // sign((lrf (x) t).w)
    Quat t;      // fix(that)
    Quat vote;
};

__device__ __forceinline__ VoteParts make_vote(Quat lrf, Quat traw) {
    VoteParts r;
    const float n2 = qdot(traw, traw);
    r.rn = (n2 > 1e-24f) ? frsqrt(n2) : 1e12f;  // x / max(|x|, eps)
    r.that = qscale(traw, r.rn);
    r.st = fsign(r.that.w);
    r.t = qscale(r.that, r.st);
    const Quat h = qham(lrf, r.t);
    r.sv = fsign(h.w);
    r.vote = qscale(h, r.sv);
    return r;
}

// g = dL/dvote -> dL/dt_raw, dL/dlrf
__device__ __forceinline__ void vote_bwd(Quat lrf, const VoteParts& vp, Quat g, Quat& g_traw, Quat& g_lrf) {
    const Quat gh = qscale(g, vp.sv);
    g_lrf = qham(gh, qconj(vp.t));                        // d(l (x) t)/dl = g (x) conj(t)
    const Quat gt = qscale(qham(qconj(lrf), gh), vp.st);  // d(l (x) t)/dt = conj(l) (x) g, through the fix
    // normalisation Jacobian (I - that that^T) / |t_raw|   (identity / eps if clamped)
    const float proj = (vp.rn < 1e12f) ? qdot(vp.that, gt) : 0.f;
    g_traw = qscale(qfma(-proj, vp.that, gt), vp.rn);
}

}  // namespace qec
