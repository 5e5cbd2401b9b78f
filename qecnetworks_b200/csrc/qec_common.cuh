// qec_common.cuh -- launch helpers, cache-hinted loads and the cooperation
// policies (how many lanes share one output capsule) used by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "qec_math.cuh"

namespace qec {

// B200: 148 SMs.  Grids of grid-stride kernels are sized in multiples of this.
constexpr int kNumSMs = 148;

#define QEC_CHECK_ARG(cond, idx) \
    do {                         \
        if (!(cond)) return -(idx); \
    } while (0)

// Launch configuration that is expensive to query (occupancy) or has to be set once per DEVICE
// (cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute): cached per device ordinal of the
// calling thread's current device, so that one process driving several GPUs -- the reference's
// nn.DataParallel, train_cls.py:27: one Python thread per GPU -- configures every one of them.
// Relaxed atomics: two threads racing on one device both compute and store the same value.
constexpr int kMaxDevices = 64;
struct PerDeviceInt {
    std::atomic<int> v[kMaxDevices];
    PerDeviceInt() {
        for (auto& x : v) x.store(-1, std::memory_order_relaxed);
    }
    // slot of the current device; nullptr (= do not cache) if the ordinal cannot be had or is out of range
    std::atomic<int>* slot() {
        int d = -1;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return nullptr;
        return &v[d];
    }
};

inline int launch_status() {
    const cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}

__device__ __forceinline__ Quat ldq(const float* p) {  // 16-byte aligned quaternion load
    const float4 v = *reinterpret_cast<const float4*>(p);
    return qmake(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ Quat ldq_nc(const float* p) {  // read-only path
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return qmake(v.x, v.y, v.z, v.w);
}
// streaming load: the vote tensors are read once per sweep and never fit in L1
__device__ __forceinline__ Quat ldq_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return qmake(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void stq(float* p, Quat q) {
    *reinterpret_cast<float4*>(p) = make_float4(q.w, q.x, q.y, q.z);
}
__device__ __forceinline__ void stq_stream(float* p, Quat q) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(q.w), "f"(q.x), "f"(q.y),
                 "f"(q.z)
                 : "memory");
}

// ---------------------------------------------------------------------------
// Cooperation policies.  A "capsule" is one output quaternion; its votes are
// swept by G lanes of a warp (GroupCoop<G>, G = 1..32) or by a whole CTA
// (BlockCoop<BLOCK>).  reduce<N>() sums an N-float accumulator over the lanes
// that share the capsule; afterwards solver() is true on the lanes that should
// run the per-capsule solve, and bcast() publishes its result to all of them.
// ---------------------------------------------------------------------------
template <int G>
struct GroupCoop {
    static constexpr int kLanes = G;
    static constexpr bool kBlock = false;
    int lane_in_group;
    __device__ __forceinline__ GroupCoop() : lane_in_group(threadIdx.x & (G - 1)) {}
    __device__ __forceinline__ int first() const { return lane_in_group; }
    __device__ __forceinline__ int stride() const { return G; }
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N]) const {
#pragma unroll
        for (int s = G >> 1; s > 0; s >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], s);
        }
    }
    __device__ __forceinline__ bool solver() const { return true; }  // all lanes redundantly: no divergence
    template <int N>
    __device__ __forceinline__ void bcast(float (&)[N]) const {}
    __device__ __forceinline__ bool leader() const { return lane_in_group == 0; }
};

template <int BLOCK>
struct BlockCoop {
    static constexpr int kLanes = BLOCK;
    static constexpr bool kBlock = true;
    static constexpr int kWarps = BLOCK / 32;
    static constexpr int kMaxN = 16;
    float* scratch;  // shared, kMaxN * kWarps + kMaxN floats
    __device__ __forceinline__ explicit BlockCoop(float* s) : scratch(s) {}
    __device__ __forceinline__ int first() const { return threadIdx.x; }
    __device__ __forceinline__ int stride() const { return BLOCK; }
    // totals end up valid in every lane of warp 0
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N]) const {
        static_assert(N <= kMaxN, "accumulator too wide");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], s);
        }
        __syncthreads();  // scratch may still be read from the previous bcast
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
                for (int w = 0; w < kWarps; ++w) s += scratch[i * kWarps + w];  // broadcast reads, fixed order
                a[i] = s;
            }
        }
    }
    __device__ __forceinline__ bool solver() const { return threadIdx.x < 32; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) const {
        static_assert(N <= kMaxN, "result too wide");
        float* out = scratch + kMaxN * kWarps;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = r[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = out[i];
    }
    __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
    static constexpr int kSmemFloats = kMaxN * kWarps + kMaxN;
};

}  // namespace qec
