// lrf_mean.cu -- per-patch mean LRF and the multi-canonical point rotation
// (reference models/qec_module.py:101-110 mean_lrf_per_patch, :35-61
// averageQuaternions, :115-120 qrotv of the neighbour points by the conjugate
// mean), forward and backward.
//
//   Mbar[b,p,i] = (1/K) sum_k q q^T,  q = lrfs[b,p,k,i]     (unweighted)
//   mean[b,p,i] = [any act != 0] * fix(top eigenvector of Mbar)
//   xrot[b,p,k,i] = qrotv(conj(mean[b,p,i]), points[b,p,k])
//
// G lanes share one (b,p,i): G = 1 when K <= 32 (pool1: K = 9), a full warp
// otherwise (pool2: K = 256).  The eigen-solve is the same register-resident
// Jacobi as in the routing kernels; the backward uses the closed-form adjoint.
#include "routing_core.cuh"

namespace qec {

constexpr int kLT = 128;

template <int G>
__global__ void __launch_bounds__(kLT)
    lrf_mean_fwd_kernel(const float* __restrict__ lrfs, const float* __restrict__ act, float* __restrict__ mean,
                        long long Ncap, int K, int I, const float* __restrict__ points = nullptr,
                        float* __restrict__ xrot = nullptr) {
    const long long gid = ((long long)blockIdx.x * kLT + threadIdx.x) / G;
    const bool active = gid < Ncap;
    const long long c = active ? gid : Ncap - 1;
    const long long bp = c / I;
    const int i = (int)(c - bp * I);
    GroupCoop<G> coop;
    float acc[11];  // M (10), number of active neighbours
    zero_acc(acc);
    for (int k = coop.first(); k < K; k += coop.stride()) {
        const size_t e = ((size_t)bp * K + k) * I + i;
        const Quat q = ldq_nc(lrfs + e * 4);
        acc_rank1(acc, 1.f, q);
        acc[10] += (__ldg(act + e) != 0.f) ? 1.f : 0.f;
    }
    coop.reduce(acc);
    const Sym4 m = acc_to_sym(acc, 1.f / (float)K);
    Quat p = qfix(top_eigvec_psd(m, WarpDone()));  // (1/K) sum q q^T is positive semi-definite
    if (!(acc[10] > 0.f)) p = qmake(0.f, 0.f, 0.f, 0.f);
    if (active && coop.leader()) stq(mean + (size_t)c * 4, p);
    // G = 1 (a handful of neighbours: the pool1 regime): the thread that owns the mean also rotates its K points,
    // so the layer's prologue is one launch, not two (config 1 is latency: every launch is ~4 us of its ~20)
    if (G == 1 && xrot != nullptr && active) {
        const Quat cj = qconj(p);
        for (int k = 0; k < K; ++k) {
            const size_t bpk = (size_t)bp * K + k;
            const float* pp = points + bpk * 3;
            const Vec3 x = qrotv(cj, vmake(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)));
            float* o = xrot + (bpk * I + i) * 3;
            o[0] = x.x; o[1] = x.y; o[2] = x.z;
        }
    }
}

__global__ void __launch_bounds__(256)
    rotate_points_kernel(const float* __restrict__ mean, const float* __restrict__ points, float* __restrict__ xrot,
                         long long N, int K, int I) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;  // (bp, k, i)
    if (e >= N) return;
    const long long bpk = e / I;
    const int i = (int)(e - bpk * I);
    const long long bp = bpk / K;
    const Quat c = qconj(ldq_nc(mean + ((size_t)bp * I + i) * 4));
    const float* pp = points + (size_t)bpk * 3;
    const Vec3 x = qrotv(c, vmake(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)));
    float* o = xrot + (size_t)e * 3;
    o[0] = x.x; o[1] = x.y; o[2] = x.z;
}

// backward: g_xrot [BP,K,I,3] (+ optional g_mean [BP,I,4]) -> g_lrfs [BP,K,I,4],
// optional g_points [BP,K,3] (accumulated with atomicAdd over i: caller zeroes it)
template <int G>
__global__ void __launch_bounds__(kLT)
    lrf_mean_bwd_kernel(const float* __restrict__ lrfs, const float* __restrict__ points,
                        const float* __restrict__ mean, const float* __restrict__ g_xrot,
                        const float* __restrict__ g_mean, float* __restrict__ g_lrfs, float* g_points,
                        long long Ncap, int K, int I) {
    const long long gid = ((long long)blockIdx.x * kLT + threadIdx.x) / G;
    const bool active = gid < Ncap;
    const long long c = active ? gid : Ncap - 1;
    const long long bp = c / I;
    const int i = (int)(c - bp * I);
    GroupCoop<G> coop;
    const Quat m = ldq_nc(mean + (size_t)c * 4);
    const Quat cj = qconj(m);
    float acc[14];  // g_conj (4), M (10)
    zero_acc(acc);
    for (int k = coop.first(); k < K; k += coop.stride()) {
        const size_t bpk = (size_t)bp * K + k;
        const size_t e = bpk * I + i;
        const Quat q = ldq_nc(lrfs + e * 4);
        const float* pp = points + bpk * 3;
        const float* gp = g_xrot + e * 3;
        Quat gq;
        Vec3 gv;
        qrotv_bwd(cj, vmake(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)), vmake(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)),
                  gq, gv);
        acc[0] += gq.w; acc[1] += gq.x; acc[2] += gq.y; acc[3] += gq.z;
        acc_rank1(acc + 4, 1.f, q);
        if (g_points != nullptr && active) {
            atomicAdd(g_points + bpk * 3 + 0, gv.x);
            atomicAdd(g_points + bpk * 3 + 1, gv.y);
            atomicAdd(g_points + bpk * 3 + 2, gv.z);
        }
    }
    coop.reduce(acc);
    // c = conj(m): dL/dm = conj(dL/dc)
    Quat gm = qmake(acc[0], -acc[1], -acc[2], -acc[3]);
    if (g_mean != nullptr) gm = qadd(gm, ldq_nc(g_mean + (size_t)c * 4));
    const float rk = 1.f / (float)K;
    const Sym4 M = acc_to_sym(acc + 4, rk);
    const float lam = qdot(m, sym_mul(M, m));
    Quat u = top_eigvec_adjoint(M, m, lam, gm);
    if (!(qdot(m, m) > 0.5f)) u = qmake(0.f, 0.f, 0.f, 0.f);
    if (!active) return;
    for (int k = coop.first(); k < K; k += coop.stride()) {
        const size_t e = ((size_t)bp * K + k) * I + i;
        const Quat q = ldq_nc(lrfs + e * 4);
        // d/dq of (1/K) q^T (sym(u m^T)) q ... = (1/K) (u <m,q> + m <u,q>)
        const Quat g = qfma(rk * qdot(m, q), u, qscale(m, rk * qdot(u, q)));
        stq(g_lrfs + e * 4, g);
    }
}


// ------------------------------------------------------------------------------------
// Tile variants for wide layers (I >= 32, K > 32: pool2, K = 256 neighbours x I = 128 channels).
// The warp-per-(bp,i) kernels above walk k with lanes 2 KB apart: every load is its own sector.  Here a
// CTA owns 32 consecutive channels i of one bp: lane = channel (a warp reads 512 contiguous bytes of
// lrfs per k), warp = slice of k; the kTW partial sums per channel meet in shared memory, warp 0 solves
// the 32 channels, and (backward) every thread writes dL/dlrfs for the (k, i) it read.
// ------------------------------------------------------------------------------------
constexpr int kTW = 16;  // warps per CTA = slices of k

__global__ void __launch_bounds__(kTW * 32)
    lrf_mean_fwd_tile_kernel(const float* __restrict__ lrfs, const float* __restrict__ act, float* __restrict__ mean,
                             int K, int I, const float* __restrict__ points, float* __restrict__ xrot) {
    __shared__ float part[kTW][11][32];
    __shared__ float ms[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (I + 31) / 32;
    const long long bp = blockIdx.x / tiles;
    const int i = (int)(blockIdx.x - bp * tiles) * 32 + lane;
    const bool live = i < I;
    float acc[11];  // M (10), number of active neighbours
    zero_acc(acc);
    if (live) {
#pragma unroll 8  // 16 neighbours per warp at K = 256: two rounds of eight loads in flight
        for (int k = warp; k < K; k += kTW) {
            const size_t e = ((size_t)bp * K + k) * I + i;
            const Quat q = ldq_nc(lrfs + e * 4);
            acc_rank1(acc, 1.f, q);
            acc[10] += (__ldg(act + e) != 0.f) ? 1.f : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < 11; ++j) part[warp][j][lane] = acc[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < 11; ++j) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kTW; ++w) t += part[w][j][lane];  // fixed order
            acc[j] = t;
        }
        const Sym4 m = acc_to_sym(acc, 1.f / (float)K);
        Quat p = qfix(top_eigvec_psd(m, WarpDone()));  // (1/K) sum q q^T is positive semi-definite
        if (!(acc[10] > 0.f)) p = qmake(0.f, 0.f, 0.f, 0.f);
        if (live) stq(mean + ((size_t)bp * I + i) * 4, p);
        ms[0][lane] = p.w; ms[1][lane] = p.x; ms[2][lane] = p.y; ms[3][lane] = p.z;
    }
    if (xrot == nullptr) return;
    // the rotation of the patch's points into the mean frame (the qrotv at the top of get_votes) in the same launch:
    // every warp rotates its slice of k for the tile's 32 channels
    __syncthreads();
    if (!live) return;
    const Quat cj = qconj(qmake(ms[0][lane], ms[1][lane], ms[2][lane], ms[3][lane]));
#pragma unroll 4
    for (int k = warp; k < K; k += kTW) {
        const size_t bpk = (size_t)bp * K + k;
        const float* pp = points + bpk * 3;
        const Vec3 x = qrotv(cj, vmake(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)));
        float* o = xrot + (bpk * I + i) * 3;
        o[0] = x.x; o[1] = x.y; o[2] = x.z;
    }
}

__global__ void __launch_bounds__(kTW * 32)
    lrf_mean_bwd_tile_kernel(const float* __restrict__ lrfs, const float* __restrict__ points,
                             const float* __restrict__ mean, const float* __restrict__ g_xrot,
                             const float* __restrict__ g_mean, float* __restrict__ g_lrfs, float* g_points, int K,
                             int I) {
    __shared__ float part[kTW][14][32];
    __shared__ float us[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (I + 31) / 32;
    const long long bp = blockIdx.x / tiles;
    const int i = (int)(blockIdx.x - bp * tiles) * 32 + lane;
    const bool live = i < I;
    const Quat m = live ? ldq_nc(mean + ((size_t)bp * I + i) * 4) : qmake(0.f, 0.f, 0.f, 0.f);
    const Quat cj = qconj(m);
    float acc[14];  // g_conj (4), M (10)
    zero_acc(acc);
    if (live) {
#pragma unroll 4
        for (int k = warp; k < K; k += kTW) {
            const size_t bpk = (size_t)bp * K + k;
            const size_t e = bpk * I + i;
            const Quat q = ldq_nc(lrfs + e * 4);
            const float* pp = points + bpk * 3;
            const float* gp = g_xrot + e * 3;
            Quat gq;
            Vec3 gv;
            qrotv_bwd(cj, vmake(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)), vmake(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)),
                      gq, gv);
            acc[0] += gq.w; acc[1] += gq.x; acc[2] += gq.y; acc[3] += gq.z;
            acc_rank1(acc + 4, 1.f, q);
            if (g_points != nullptr) {
                atomicAdd(g_points + bpk * 3 + 0, gv.x);
                atomicAdd(g_points + bpk * 3 + 1, gv.y);
                atomicAdd(g_points + bpk * 3 + 2, gv.z);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 14; ++j) part[warp][j][lane] = acc[j];
    __syncthreads();
    const float rk = 1.f / (float)K;
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < 14; ++j) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kTW; ++w) t += part[w][j][lane];  // fixed order
            acc[j] = t;
        }
        // c = conj(m): dL/dm = conj(dL/dc)
        Quat gm = qmake(acc[0], -acc[1], -acc[2], -acc[3]);
        if (g_mean != nullptr && live) gm = qadd(gm, ldq_nc(g_mean + ((size_t)bp * I + i) * 4));
        const Sym4 M = acc_to_sym(acc + 4, rk);
        const float lam = qdot(m, sym_mul(M, m));
        Quat u = top_eigvec_adjoint(M, m, lam, gm);
        if (!(qdot(m, m) > 0.5f)) u = qmake(0.f, 0.f, 0.f, 0.f);
        us[0][lane] = u.w; us[1][lane] = u.x; us[2][lane] = u.y; us[3][lane] = u.z;
    }
    __syncthreads();
    if (!live) return;
    const Quat u = qmake(us[0][lane], us[1][lane], us[2][lane], us[3][lane]);
#pragma unroll 8
    for (int k = warp; k < K; k += kTW) {
        const size_t e = ((size_t)bp * K + k) * I + i;
        const Quat q = ldq_nc(lrfs + e * 4);
        // d/dq of (1/K) q^T (sym(u m^T)) q ... = (1/K) (u <m,q> + m <u,q>)
        const Quat g = qfma(rk * qdot(m, q), u, qscale(m, rk * qdot(u, q)));
        stq(g_lrfs + e * 4, g);
    }
}

}  // namespace qec

using namespace qec;

extern "C" int qec_lrf_mean_fwd(const float* lrfs, const float* act, const float* points, float* mean, float* xrot,
                                int BP, int K, int I, void* stream) {
    QEC_CHECK_ARG(lrfs, 1);
    QEC_CHECK_ARG(act, 2);
    QEC_CHECK_ARG(mean, 4);
    QEC_CHECK_ARG(BP > 0, 6);
    QEC_CHECK_ARG(K > 0, 7);
    QEC_CHECK_ARG(I > 0, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long Ncap = (long long)BP * I;
    if (xrot != nullptr) QEC_CHECK_ARG(points, 3);
    if (K <= 32) {
        const long long blocks = (Ncap + kLT - 1) / kLT;
        lrf_mean_fwd_kernel<1><<<(unsigned)blocks, kLT, 0, st>>>(lrfs, act, mean, Ncap, K, I, points, xrot);
        return launch_status();  // rotation done in the same launch
    } else if (I >= 32) {
        const long long blocks = (long long)BP * ((I + 31) / 32);
        lrf_mean_fwd_tile_kernel<<<(unsigned)blocks, kTW * 32, 0, st>>>(lrfs, act, mean, K, I, points, xrot);
        return launch_status();  // rotation done in the same launch
    } else {
        const long long blocks = (Ncap * 32 + kLT - 1) / kLT;
        lrf_mean_fwd_kernel<32><<<(unsigned)blocks, kLT, 0, st>>>(lrfs, act, mean, Ncap, K, I);
    }
    if (xrot != nullptr) {
        QEC_CHECK_ARG(points, 3);
        const long long N = (long long)BP * K * I;
        rotate_points_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(mean, points, xrot, N, K, I);
    }
    return launch_status();
}

extern "C" int qec_lrf_mean_bwd(const float* lrfs, const float* points, const float* mean, const float* g_xrot,
                                const float* g_mean, float* g_lrfs, float* g_points, int BP, int K, int I,
                                void* stream) {
    QEC_CHECK_ARG(lrfs, 1);
    QEC_CHECK_ARG(points, 2);
    QEC_CHECK_ARG(mean, 3);
    QEC_CHECK_ARG(g_xrot, 4);
    QEC_CHECK_ARG(g_lrfs, 6);
    QEC_CHECK_ARG(BP > 0, 8);
    QEC_CHECK_ARG(K > 0, 9);
    QEC_CHECK_ARG(I > 0, 10);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long Ncap = (long long)BP * I;
    if (K <= 32) {
        const long long blocks = (Ncap + kLT - 1) / kLT;
        lrf_mean_bwd_kernel<1><<<(unsigned)blocks, kLT, 0, st>>>(lrfs, points, mean, g_xrot, g_mean, g_lrfs,
                                                                 g_points, Ncap, K, I);
    } else if (I >= 32) {
        const long long blocks = (long long)BP * ((I + 31) / 32);
        lrf_mean_bwd_tile_kernel<<<(unsigned)blocks, kTW * 32, 0, st>>>(lrfs, points, mean, g_xrot, g_mean, g_lrfs,
                                                                          g_points, K, I);
    } else {
        const long long blocks = (Ncap * 32 + kLT - 1) / kLT;
        lrf_mean_bwd_kernel<32><<<(unsigned)blocks, kLT, 0, st>>>(lrfs, points, mean, g_xrot, g_mean, g_lrfs,
                                                                  g_points, Ncap, K, I);
    }
    return launch_status();
}
