// symeig.cu -- the L0 boundary: batched symmetric 4x4 eigensolve and its
// backward, as a drop-in for torch_autograd_solver.symeig on [n,4,4] input.
//
// forward  replaces cusolver_batch_eigh -> cusolverDnSsyevjBatched
//          (reference models/pytorch-cusolver/torch_cusolver.cpp:194-288 as called
//          from models/pytorch-autograd-solver/torch_autograd_solver/__init__.py:74-81):
//          eigenvalues ascending, eigenvectors as ROWS of V, upper triangle of the
//          row-major input.  No handle, no workspace malloc, no host sync: one
//          thread per matrix, Jacobi in registers, a device counter of
//          non-converged matrices instead of check_jacobi's host loop (:150-173).
// backward replaces batch_symeig_backward
//          (reference models/pytorch-autograd-solver/torch_autograd_solver.cpp:100-169):
//          gM = V (F o V^T gV) V^T + V diag(gw) V^T, F_ij = 1/(w_j - w_i).
#include "qec_common.cuh"

namespace qec {

template <class T>
__device__ __forceinline__ void cswap(bool c, T& a, T& b) {
    const T x = a, y = b;
    a = c ? y : x;
    b = c ? x : y;
}

__global__ void __launch_bounds__(128)
    symeig4_fwd_kernel(const float* __restrict__ M, float* __restrict__ w, float* __restrict__ V, int* status,
                       long long n) {
    const long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    const bool active = gid < n;
    const long long c = active ? gid : n - 1;
    const float* m = M + (size_t)c * 16;
    float A[4][4], E[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(m) + i);
        A[i][0] = r.x; A[i][1] = r.y; A[i][2] = r.z; A[i][3] = r.w;  // only j >= i is used (UPLO = 'U')
    }
    // scale to unit max-abs so the rotation thresholds are range-independent
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) mx = fmaxf(mx, fabsf(A[i][j]));
    const float sc = (mx > 0.f) ? 1.f / mx : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) A[i][j] *= sc;
    const int sweeps = jacobi4(A, E, WarpDone());
    const bool conv = jacobi_converged(A) || sweeps < QEC_JACOBI_MAX_SWEEPS;
    float ev[4] = {A[0][0] * mx, A[1][1] * mx, A[2][2] * mx, A[3][3] * mx};
    // ascending sorting network on (ev[j], column j of E)
#define QEC_CE(a, b)                                              \
    {                                                             \
        const bool sw = ev[a] > ev[b];                            \
        cswap(sw, ev[a], ev[b]);                                  \
        cswap(sw, E[0][a], E[0][b]); cswap(sw, E[1][a], E[1][b]); \
        cswap(sw, E[2][a], E[2][b]); cswap(sw, E[3][a], E[3][b]); \
    }
    QEC_CE(0, 1) QEC_CE(2, 3) QEC_CE(0, 2) QEC_CE(1, 3) QEC_CE(1, 2)
#undef QEC_CE
    if (!active) return;
    *reinterpret_cast<float4*>(w + (size_t)c * 4) = make_float4(ev[0], ev[1], ev[2], ev[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // row j of V = eigenvector j, renormalised
        const float rn = frsqrt(E[0][j] * E[0][j] + E[1][j] * E[1][j] + E[2][j] * E[2][j] + E[3][j] * E[3][j]);
        reinterpret_cast<float4*>(V + (size_t)c * 16)[j] =
            make_float4(E[0][j] * rn, E[1][j] * rn, E[2][j] * rn, E[3][j] * rn);
    }
    if (!conv && status != nullptr) atomicAdd(status, 1);
}

// fold = 1 reproduces the reference's final "triu(g) + triu(g^T, 1)" (upper=True branch)
__global__ void __launch_bounds__(128)
    symeig4_bwd_kernel(const float* __restrict__ gw, const float* __restrict__ gV, const float* __restrict__ w,
                       const float* __restrict__ V, float* __restrict__ gM, long long n, int fold) {
    const long long c = (long long)blockIdx.x * 128 + threadIdx.x;
    if (c >= n) return;
    float R[4][4], G[4][4], ev[4], gwv[4];  // R = V rows (R[e][c]), G = gV rows
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(V + (size_t)c * 16) + e);
        R[e][0] = r.x; R[e][1] = r.y; R[e][2] = r.z; R[e][3] = r.w;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gV != nullptr) g = __ldg(reinterpret_cast<const float4*>(gV + (size_t)c * 16) + e);
        G[e][0] = g.x; G[e][1] = g.y; G[e][2] = g.z; G[e][3] = g.w;
        ev[e] = __ldg(w + (size_t)c * 4 + e);
        gwv[e] = (gw != nullptr) ? __ldg(gw + (size_t)c * 4 + e) : 0.f;
    }
    // inner[i][j] = F_ij * <v_i, gv_j>,  F_ij = 1/(w_j - w_i) (0 on the diagonal), + diag(gw)
    float inner[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float d = R[i][0] * G[j][0] + R[i][1] * G[j][1] + R[i][2] * G[j][2] + R[i][3] * G[j][3];
            inner[i][j] = (i == j) ? gwv[i] : d / (ev[j] - ev[i]);
        }
    // gM = Vc inner Vc^T, Vc[a][i] = R[i][a]
    float tmp[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            tmp[a][j] = R[0][a] * inner[0][j] + R[1][a] * inner[1][j] + R[2][a] * inner[2][j] + R[3][a] * inner[3][j];
    float out[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            out[a][b] = tmp[a][0] * R[0][b] + tmp[a][1] * R[1][b] + tmp[a][2] * R[2][b] + tmp[a][3] * R[3][b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float r[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (fold)
                r[b] = (b > a) ? out[a][b] + out[b][a] : ((b == a) ? out[a][a] : 0.f);
            else
                r[b] = out[a][b];
        }
        reinterpret_cast<float4*>(gM + (size_t)c * 16)[a] = make_float4(r[0], r[1], r[2], r[3]);
    }
}

}  // namespace qec

using namespace qec;

extern "C" int qec_symeig4_fwd(const float* M, float* w, float* V, int* status, long long n, void* stream) {
    QEC_CHECK_ARG(M, 1);
    QEC_CHECK_ARG(w, 2);
    QEC_CHECK_ARG(V, 3);
    QEC_CHECK_ARG(n > 0, 5);
    symeig4_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(M, w, V, status,
                                                                                                  n);
    return launch_status();
}

extern "C" int qec_symeig4_bwd(const float* gw, const float* gV, const float* w, const float* V, float* gM,
                               long long n, int fold, void* stream) {
    QEC_CHECK_ARG(gw || gV, 1);
    QEC_CHECK_ARG(w, 3);
    QEC_CHECK_ARG(V, 4);
    QEC_CHECK_ARG(gM, 5);
    QEC_CHECK_ARG(n > 0, 6);
    symeig4_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(gw, gV, w, V, gM,
                                                                                                  n, fold);
    return launch_status();
}
