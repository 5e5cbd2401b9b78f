// routing_fused_shared.cuh -- pieces shared by the two generations of cluster-resident routing
// kernels (routing_fused.cu: scalar, any I; routing_fused2.cu: packed vote pairs, even I):
// the cluster-wide reducer, the per-capsule pose solve, the persistent-grid sizing.
#pragma once

#include <cooperative_groups.h>

#include "routing_core.cuh"

namespace qec {
namespace cg = cooperative_groups;

constexpr int kFT = 512;      // threads per CTA
constexpr int kJcMax = 4096;  // resident votes per CTA
constexpr int kTB = 3;        // routing iterations the cached backward supports

// After this, lane L holds (in the return value) the warp-wide total of a[idx16(L)].
__device__ __forceinline__ int idx16(int lane) { return (lane >> 1) & 15; }
__device__ __forceinline__ float warp_reduce_scatter16(float (&a)[16], int lane) {
#pragma unroll
    for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < h) {
                const float send = up ? a[i] : a[i + h];
                const float keep = up ? a[i + h] : a[i];
                a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
    }
    return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

template <int CL, int NT = kFT>
struct ClusterReducer {
    static constexpr int kFW = NT / 32;
    float* scratch;  // [16][kFW] warp partials
    float* part;     // [2][16] this CTA's published partial sums (double-buffered across sweeps)
    float* out;      // [16] solve result broadcast inside the CTA
    unsigned phase;
    __device__ __forceinline__ explicit ClusterReducer(float* s)
        : scratch(s), part(s + 16 * kFW), out(s + 16 * kFW + 32), phase(0) {}
    static constexpr int kSmemFloats = 16 * kFW + 48;

    // Sum a[0..N) over every thread of every CTA of the cluster.  On return the totals are
    // valid in all lanes of warp 0 of every CTA (identical bits in every CTA).
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N]) {
        static_assert(N <= 16, "accumulator too wide");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = (i < N) ? a[i] : 0.f;
        const float mine_v = warp_reduce_scatter16(w, lane);
        __syncthreads();  // scratch / out may still be read from the previous round
        if ((lane & 1) == 0) scratch[idx16(lane) * kFW + warp] = mine_v;
        __syncthreads();
        if (warp == 0) {
            float s = 0.f;
            if (lane < 16) {
#pragma unroll
                for (int x = 0; x < kFW; ++x) s += scratch[lane * kFW + x];  // fixed order
            }
            if (CL > 1) {
                float* mine = part + (phase & 1u) * 16;
                if (lane < 16) mine[lane] = s;
                cg::this_cluster().sync();  // all threads of all CTAs (see below)
                s = 0.f;
                if (lane < 16) {
#pragma unroll
                    for (int r = 0; r < CL; ++r) s += cg::this_cluster().map_shared_rank(mine, r)[lane];
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = __shfl_sync(0xffffffffu, s, i);
        } else if (CL > 1) {
            cg::this_cluster().sync();
        }
        ++phase;
    }
    __device__ __forceinline__ bool solver() const { return threadIdx.x < 32; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = r[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = out[i];
    }
};

__device__ __forceinline__ float ldg_stream_f32(const float* p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));  // not volatile: schedulable
    return v;
}
__device__ __forceinline__ void red_add_q(float* p, Quat q) {  // REDG.E.ADD.F32x4
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(q.w), "f"(q.x), "f"(q.y), "f"(q.z)
                 : "memory");
}
__device__ __forceinline__ Quat lds_q(const float4* p) {
    const float4 v = *p;
    return qmake(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void sts_q(float4* p, Quat q) { *p = make_float4(q.w, q.x, q.y, q.z); }

// acc: M (10), [10] = sum |b|, [11] = sum b
__device__ __forceinline__ Quat solve_pose(const float* acc, bool valid) {
    const Sym4 m = acc_to_sym(acc, 1.f / fmaxf(acc[10], 1e-12f));  // F.normalize(p=1) of the weights
    // sum b == sum |b| bit-exactly (same values, same order) unless a weight is negative
    Quat p = qfix(top_eigvec_weighted(m, acc[11] != acc[10], WarpDone()));
    if (!valid) p = qmake(0.f, 0.f, 0.f, 0.f);
    return p;
}

// persistent grid: as many clusters as can be co-resident, never more than there are capsules
template <class Kern>
static int fused_grid(Kern kern, int CL, int threads, size_t smem, long long Nc, cudaLaunchConfig_t& cfg,
                      cudaLaunchAttribute* attr, PerDeviceInt& cache) {
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.numAttrs = 0;
    cfg.attrs = attr;
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.numAttrs = 1;
    }
    std::atomic<int>* slot = cache.slot();  // per kernel instantiation AND per device (the opt-in below is per device)
    int clusters = slot ? slot->load(std::memory_order_relaxed) : -1;
    if (clusters < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (CL > 1) {
            cfg.gridDim = dim3(CL * 4 * kNumSMs);  // upper bound for the occupancy query
            e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
            if (e != cudaSuccess) return (int)e;
        } else {
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
            if (e != cudaSuccess) return (int)e;
            clusters = per_sm * kNumSMs;
        }
        if (clusters < 1) return (int)cudaErrorLaunchOutOfResources;
        if (slot) slot->store(clusters, std::memory_order_relaxed);
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

}  // namespace qec

#define QEC_DISPATCH_CL(FN, ...)               \
    switch (CL) {                              \
        case 1: return FN<1>(__VA_ARGS__);     \
        case 2: return FN<2>(__VA_ARGS__);     \
        case 4: return FN<4>(__VA_ARGS__);     \
        default: return FN<8>(__VA_ARGS__);    \
    }

extern "C" int qec_routing_fused_fwd_scalar(const float* t_oqi, const float* lrfs, const float* act,
                                            const float* alpha, const float* beta, float* pose, float* agreement,
                                            float* poses_all, float* stats, int BP, int K, int I, int O, int T,
                                            void* stream);
extern "C" int qec_routing_fused_bwd_scalar(const float* t_oqi, const float* lrfs, const float* act,
                                            const float* alpha, const float* beta, const float* poses_all,
                                            const float* stats, const float* g_pose, const float* g_agreement,
                                            float* g_toqi, float* g_lrfs, float* g_act, float* g_alpha_beta, int BP,
                                            int K, int I, int O, int T, void* stream);
extern "C" int qec_routing_fused_supported(int K, int I, int O, int T);
