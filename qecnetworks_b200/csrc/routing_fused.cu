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
// time; CTA r holds votes [r*Jc, (r+1)*Jc) as float4 in its shared memory (Jc <= 4096:
// 64 KB, two CTAs per SM so that one capsule's eigen-solve latency hides behind another's
// sweeps).  Each sweep reduces 10-15 floats inside the CTA (warp shuffles + shared memory),
// publishes the CTA partial in its own shared memory, cluster-barriers, and every CTA sums
// all partials through distributed shared memory in the same fixed order, so all CTAs of
// the cluster run the same Jacobi / Cholesky solve on identical inputs and no pose
// broadcast across CTAs is needed.  HBM traffic per capsule: the transforms once (forward),
// plus their gradient once (backward).
#include <cooperative_groups.h>

#include "routing_core.cuh"
#include "vote_math.cuh"

namespace cg = cooperative_groups;

namespace qec {

constexpr int kFT = 512;      // threads per CTA
constexpr int kJcMax = 4096;  // resident votes per CTA
constexpr int kCoopN = 16;

template <int CL>
struct ClusterCoop {
    static constexpr int kWarps = kFT / 32;
    float* scratch;  // [kCoopN * kWarps] intra-CTA staging
    float* part;     // [2][kCoopN] this CTA's published partial sums (double-buffered)
    float* out;      // [kCoopN] solve result broadcast inside the CTA
    mutable unsigned phase;
    __device__ __forceinline__ ClusterCoop(float* s) : scratch(s), part(s + kCoopN * kWarps),
                                                       out(s + kCoopN * kWarps + 2 * kCoopN), phase(0) {}
    static constexpr int kSmemFloats = kCoopN * kWarps + 3 * kCoopN;
    __device__ __forceinline__ int first() const { return threadIdx.x; }
    __device__ __forceinline__ int stride() const { return kFT; }
    // totals end up valid in every lane of warp 0 of EVERY CTA of the cluster
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N]) const {
        static_assert(N <= kCoopN, "accumulator too wide");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], s);
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) scratch[i * kWarps + warp] = a[i];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += scratch[i * kWarps + w];
                a[i] = s;
            }
        }
        if (CL > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            float* mine = part + (phase & 1u) * kCoopN;
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) mine[i] = a[i];
            }
            cluster.sync();  // release/acquire: every CTA's partial is visible cluster-wide
            if (warp == 0) {
                float s = 0.f;
                if (lane < N) {
#pragma unroll
                    for (int r = 0; r < CL; ++r) s += cluster.map_shared_rank(mine, r)[lane];  // DSMEM, fixed order
                }
#pragma unroll
                for (int i = 0; i < N; ++i) a[i] = __shfl_sync(0xffffffffu, s, i);
            }
            ++phase;
        }
    }
    __device__ __forceinline__ bool solver() const { return threadIdx.x < 32; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) const {
        static_assert(N <= kCoopN, "result too wide");
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = r[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = out[i];
    }
    __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
};

struct VotesShared {
    const float4* sv;
    __device__ __forceinline__ Quat get(int jl) const {
        const float4 v = sv[jl];
        return qmake(v.x, v.y, v.z, v.w);
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

struct CapsuleGeom {
    const float* t_oqi;  // [BP*K, 4*I*O]
    const float* lrfs;   // [BP*K*I, 4]
    long long bp;
    int o, K, I, O, j_begin;
    __device__ __forceinline__ void locate(int j, size_t& toff, size_t& loff) const {
        const int k = j / I, i = j - k * I;
        const size_t row = (size_t)bp * K + k;
        toff = row * (size_t)(4 * I * O) + (size_t)o * 4 * I + i;
        loff = row * I + i;
    }
};

// generate this CTA's votes into shared memory (each thread later re-reads only its own slots)
__device__ __forceinline__ void generate_votes(const CapsuleGeom& g, float4* sv, int nloc) {
    const int I = g.I;
    int j = g.j_begin + threadIdx.x;
    int k = j / I, i = j - k * I;
    const int dk = kFT / I, di = kFT - dk * I;
    for (int jl = threadIdx.x; jl < nloc; jl += kFT) {
        const size_t row = (size_t)g.bp * g.K + k;
        const float* tp = g.t_oqi + row * (size_t)(4 * I * g.O) + (size_t)g.o * 4 * I + i;
        const Quat traw = qmake(ldg_stream_f32(tp), ldg_stream_f32(tp + I), ldg_stream_f32(tp + 2 * I),
                                ldg_stream_f32(tp + 3 * I));
        const Quat lrf = ldq_nc(g.lrfs + (row * I + i) * 4);
        const Quat v = make_vote(lrf, traw).vote;
        sv[jl] = make_float4(v.w, v.x, v.y, v.z);
        i += di;
        k += dk;
        if (i >= I) {
            i -= I;
            ++k;
        }
    }
}

// receives dL/dvote for local vote jl: recompute the vote's parts from the (L2-resident)
// transform and LRF, push the gradient through fix / Hamilton / fix / normalise
struct SinkFused {
    CapsuleGeom g;
    float* g_toqi;
    float* g_lrfs;  // nullable
    __device__ __forceinline__ void put(int jl, Quat gv) {
        size_t toff, loff;
        g.locate(g.j_begin + jl, toff, loff);
        const int I = g.I;
        const float* tp = g.t_oqi + toff;
        const Quat traw = qmake(__ldg(tp), __ldg(tp + I), __ldg(tp + 2 * I), __ldg(tp + 3 * I));
        const Quat lrf = ldq_nc(g.lrfs + loff * 4);
        const VoteParts vp = make_vote(lrf, traw);
        Quat gt, gl;
        vote_bwd(lrf, vp, gv, gt, gl);
        float* gp = g_toqi + toff;
        gp[0] = gt.w;
        gp[I] = gt.x;
        gp[2 * I] = gt.y;
        gp[3 * I] = gt.z;
        if (g_lrfs != nullptr) red_add_q(g_lrfs + loff * 4, gl);  // sum over the O capsules of the patch
    }
};

template <int CL, int TMAX>
__global__ void __launch_bounds__(kFT, 2)
    routing_fused_fwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, float* __restrict__ pose,
                             float* __restrict__ agreement, float* __restrict__ poses_all,
                             float* __restrict__ stats, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sv = reinterpret_cast<float4*>(smem_raw);
    ClusterCoop<CL> coop(reinterpret_cast<float*>(sv + kJcMax));
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = (J + CL - 1) / CL;
    const int j_begin = min(J, rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    for (long long c = cid; c < Nc; c += ncl) {
        const long long bp = c / O;
        CapsuleGeom g{t_oqi, lrfs, bp, (int)(c - bp * O), K, I, O, j_begin};
        generate_votes(g, sv, nloc);
        VotesShared vs{sv};
        const float* act_row = act + (size_t)bp * J + j_begin;
        Quat ph[TMAX];
#pragma unroll
        for (int t = 0; t < TMAX; ++t) ph[t] = qmake(0.f, 0.f, 0.f, 0.f);
        float s, n;
        const float agr = routing_forward_capsule<TMAX>(coop, vs, act_row, nloc, T, alpha, beta, ph, s, n);
        if (rank == 0 && coop.leader()) {
#pragma unroll
            for (int t = 0; t < TMAX; ++t)
                if (t < T) {
                    stq(poses_all + ((size_t)t * Nc + c) * 4, ph[t]);
                    if (t == T - 1) stq(pose + (size_t)c * 4, ph[t]);
                }
            agreement[c] = agr;
            stats[2 * c] = s;
            stats[2 * c + 1] = n;
        }
        __syncthreads();  // the next capsule overwrites the resident votes
    }
    if (CL > 1) cg::this_cluster().sync();  // no CTA may exit while a peer can still read its partials
}

template <int CL, int TMAX>
__global__ void __launch_bounds__(kFT, 1)
    routing_fused_bwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                             const float* __restrict__ act, const float* __restrict__ alpha_p,
                             const float* __restrict__ beta_p, const float* __restrict__ poses_all,
                             const float* __restrict__ stats, const float* __restrict__ g_pose,
                             const float* __restrict__ g_agr, float* __restrict__ g_toqi, float* g_lrfs, float* g_act,
                             float* g_alpha_beta, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sv = reinterpret_cast<float4*>(smem_raw);
    ClusterCoop<CL> coop(reinterpret_cast<float*>(sv + kJcMax));
    int rank = 0;
    if (CL > 1) rank = (int)cg::this_cluster().block_rank();
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = (J + CL - 1) / CL;
    const int j_begin = min(J, rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    float ga_tot = 0.f, gb_tot = 0.f;
    for (long long c = cid; c < Nc; c += ncl) {
        const long long bp = c / O;
        CapsuleGeom g{t_oqi, lrfs, bp, (int)(c - bp * O), K, I, O, j_begin};
        generate_votes(g, sv, nloc);
        VotesShared vs{sv};
        SinkFused sink{g, g_toqi, g_lrfs};
        const float* act_row = act + (size_t)bp * J + j_begin;
        float* g_act_row = (g_act != nullptr) ? g_act + (size_t)bp * J + j_begin : nullptr;
        Quat ph[TMAX];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            ph[t] = (t < T) ? ldq_nc(poses_all + ((size_t)t * Nc + c) * 4) : qmake(0.f, 0.f, 0.f, 0.f);
        float ga, gb;
        routing_backward_capsule<TMAX>(coop, vs, sink, act_row, g_act_row, nloc, T, alpha, beta, ph,
                                       __ldg(stats + 2 * c), __ldg(stats + 2 * c + 1),
                                       ldq_nc(g_pose + (size_t)c * 4), __ldg(g_agr + c), g_act != nullptr, ga, gb);
        ga_tot += ga;
        gb_tot += gb;
        __syncthreads();
    }
    if (rank == 0 && coop.leader()) {
        atomicAdd(g_alpha_beta, ga_tot);
        atomicAdd(g_alpha_beta + 1, gb_tot);
    }
    if (CL > 1) cg::this_cluster().sync();
}

constexpr size_t kFusedSmem = sizeof(float4) * kJcMax + sizeof(float) * (kCoopN * (kFT / 32) + 3 * kCoopN);

// persistent grid: as many clusters as can be co-resident (the solve latency of one capsule
// hides behind the sweeps of the others), never more than there are capsules
template <class Kern>
static int fused_grid(Kern kern, int CL, long long Nc, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr,
                      int& cached_clusters) {
    cfg.blockDim = dim3(kFT);
    cfg.dynamicSmemBytes = kFusedSmem;
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
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
        if (e != cudaSuccess) return (int)e;
        if (CL > 1) {
            cfg.gridDim = dim3(CL * 2 * kNumSMs);  // upper bound for the occupancy query
            e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
            if (e != cudaSuccess) return (int)e;
        } else {
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFT, kFusedSmem);
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

template <int CL, int TMAX>
static int launch_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                            const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                            long long Nc, int K, int I, int O, int T, cudaStream_t st) {
    auto kern = routing_fused_fwd_kernel<CL, TMAX>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static int cached = -1;
    const int rc = fused_grid(kern, CL, Nc, cfg, attr, cached);
    if (rc) return rc;
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all,
                                             stats, Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

template <int CL, int TMAX>
static int launch_fused_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                            const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                            const float* g_agr, float* g_toqi, float* g_lrfs, float* g_act, float* g_ab, long long Nc,
                            int K, int I, int O, int T, cudaStream_t st) {
    auto kern = routing_fused_bwd_kernel<CL, TMAX>;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static int cached = -1;
    const int rc = fused_grid(kern, CL, Nc, cfg, attr, cached);
    if (rc) return rc;
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose,
                                             g_agr, g_toqi, g_lrfs, g_act, g_ab, Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

}  // namespace qec

using namespace qec;

#define QEC_DISPATCH_CL(FN, ...)                          \
    switch (CL) {                                         \
        case 1: return FN<1, 3>(__VA_ARGS__);             \
        case 2: return FN<2, 3>(__VA_ARGS__);             \
        case 4: return FN<4, 3>(__VA_ARGS__);             \
        default: return FN<8, 3>(__VA_ARGS__);            \
    }

// 1 if this geometry is served by the fused resident kernels, else 0 (callers then use
// qec_votes_* + qec_routing_*)
extern "C" int qec_routing_fused_supported(int K, int I, int O, int T) {
    if (K <= 0 || I <= 0 || O <= 0) return 0;
    const long long J = (long long)K * I;
    if (J > 8ll * kJcMax) return 0;
    return (T >= 1 && T <= 3 && pick_cluster((int)J) > 0) ? 1 : 0;
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
