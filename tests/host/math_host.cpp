// Host build of qec_math.cuh for CPU-side unit tests (test scaffolding only).
#include "../../qecnetworks_b200/csrc/qec_math.cuh"
#include "../../qecnetworks_b200/csrc/adam_math.cuh"

using namespace qec;

static Sym4 load_sym(const float* m) {
    Sym4 s;
    s.a00 = m[0]; s.a01 = m[1]; s.a02 = m[2]; s.a03 = m[3]; s.a11 = m[4];
    s.a12 = m[5]; s.a13 = m[6]; s.a22 = m[7]; s.a23 = m[8]; s.a33 = m[9];
    return s;
}

extern "C" {

void h_top_eigvec(int n, const float* M10, float* v, float* lambda, int* sweeps) {
    for (int i = 0; i < n; ++i) {
        Sym4 s = load_sym(M10 + 10 * i);
        float A[4][4], V[4][4];
        sym_to_array(s, A);
        sweeps[i] = jacobi4(A, V, ThreadDone());
        float lam;
        Quat q = top_eigvec(s, lam, ThreadDone());
        v[4 * i + 0] = q.w; v[4 * i + 1] = q.x; v[4 * i + 2] = q.y; v[4 * i + 3] = q.z;
        lambda[i] = lam;
    }
}

void h_top_eigvec_psd(int n, const float* M10, float* v, int* steps) {
    for (int i = 0; i < n; ++i) {
        Sym4 s = load_sym(M10 + 10 * i);
        int cnt = 0;
        Quat q = top_eigvec_psd(s, [&cnt](bool d) { ++cnt; return d; });
        v[4 * i + 0] = q.w; v[4 * i + 1] = q.x; v[4 * i + 2] = q.y; v[4 * i + 3] = q.z;
        steps[i] = cnt;
    }
}

void h_top_eigvec_psd2(int n, const float* M10, float* v, int* rounds) {
    for (int i = 0; i < n; ++i) {
        Sym4 s = load_sym(M10 + 10 * i);
        int cnt = 0;
        Quat q = top_eigvec_psd2(s, [&cnt](bool d) { ++cnt; return d; });
        v[4 * i + 0] = q.w; v[4 * i + 1] = q.x; v[4 * i + 2] = q.y; v[4 * i + 3] = q.z;
        rounds[i] = cnt;
    }
}

void h_jacobi_full(int n, const float* M10, float* w, float* Vout) {
    for (int i = 0; i < n; ++i) {
        Sym4 s = load_sym(M10 + 10 * i);
        float A[4][4], V[4][4];
        sym_to_array(s, A);
        jacobi4(A, V, ThreadDone());
        for (int j = 0; j < 4; ++j) {
            w[4 * i + j] = A[j][j];
            for (int k = 0; k < 4; ++k) Vout[16 * i + 4 * j + k] = V[j][k];
        }
    }
}

void h_one_minus_dist_estrin(int n, const float* x, float* out) {
    for (int i = 0; i < n; ++i) out[i] = one_minus_dist_estrin(x[i]);
}

void h_one_minus_dist(int n, const float* x, float* out, float* grad) {
    for (int i = 0; i < n; ++i) {
        out[i] = one_minus_dist(x[i]);
        grad[i] = dist_grad(x[i]);
    }
}

void h_adjoint(int n, const float* M10, const float* p, const float* lambda, const float* g, float* u) {
    for (int i = 0; i < n; ++i) {
        Sym4 s = load_sym(M10 + 10 * i);
        Quat pp = qmake(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
        Quat gg = qmake(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
        Quat r = top_eigvec_adjoint(s, pp, lambda[i], gg);
        u[4 * i] = r.w; u[4 * i + 1] = r.x; u[4 * i + 2] = r.y; u[4 * i + 3] = r.z;
    }
}

void h_qrotv(int n, const float* q, const float* v, const float* g, float* x, float* gq, float* gv) {
    for (int i = 0; i < n; ++i) {
        Quat qq = qmake(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        Vec3 vv = vmake(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        Vec3 gg = vmake(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
        Vec3 r = qrotv(qq, vv);
        x[3 * i] = r.x; x[3 * i + 1] = r.y; x[3 * i + 2] = r.z;
        Quat a; Vec3 b;
        qrotv_bwd(qq, vv, gg, a, b);
        gq[4 * i] = a.w; gq[4 * i + 1] = a.x; gq[4 * i + 2] = a.y; gq[4 * i + 3] = a.z;
        gv[3 * i] = b.x; gv[3 * i + 1] = b.y; gv[3 * i + 2] = b.z;
    }
}

// `steps` Adam steps on n elements with the element update and bias corrections of qec_adam_flat; grads: [steps][n]
void h_adam_steps(int n, int steps, float* p, const float* grads, float* m, float* v, float lr, double beta1, double beta2,
                  float eps, float weight_decay, float grad_scale) {
    const AdamHyper h = adam_hyper(lr, beta1, beta2, eps, weight_decay, grad_scale, 0);
    for (int t = 1; t <= steps; ++t) {
        float c1, c2;
        adam_bias_corrections(h, t, c1, c2);
        for (int i = 0; i < n; ++i) {
            float g = grads[(size_t)(t - 1) * n + i];
            adam_one(p[i], g, m[i], v[i], h, c1, c2);
        }
    }
}
}
