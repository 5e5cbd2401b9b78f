// linear_small.cu -- the small FP32 contractions around the vote generator (reference models/qec_module.py:123-124):
//
//   qec_linear_fwd / _bwd    y = x W^T + b for a narrow layer (out features O <= 64): quater_gen1 of the pool2 regime,
//                            nn.Linear(3*I, O) over M = B*P*K rows.  The library SGEMMs are badly shaped for it (40 output
//                            columns): 15 + 4 us forward, 17 + 28 + 4 + 11 us backward (two SGEMMs, a split-K reduction
//                            and a column sum) at BASELINE config 2, against ~5 us of memory traffic.
//   qec_compose_fwd / _bwd   (W1, b1, W2, b2) -> Wc = W2 W1, bc = W2 b1 + b2 and its adjoint: the two Linears composed for
//                            the register-resident regime (I = 1: W1 is [O x 3]).  Was 2 + 4 library launches of 4-16 us.
//
// Plain FP32 FMA (no tensor cores: the contractions are 3 .. 384 long and the outputs narrow), deterministic (fixed
// summation order: partial sums per row block, folded by a second launch; no atomics).
#include <cstdint>

#include "qec_common.cuh"

namespace qec {
namespace lin {

constexpr int kRows = 64;      // rows of x per CTA
constexpr int kChunk = 128;    // in-features per shared-memory chunk
constexpr int kPitch = 132;    // floats per shared-memory row: 16-byte aligned, 4 consecutive rows on different banks
constexpr int kThreads = 256;

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

__device__ __forceinline__ void cp_async_zfill(float* dst_smem, const float* src, int bytes, int valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    if (bytes == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(valid) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(valid) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [r0, r0 + nr_pad) x columns [c0, c0 + kChunk) of a row-major [R x C] matrix -> dst[nr_pad][kPitch] by cp.async
// (fire and forget: every request of the chunk is in flight at once), zero-filled beyond R / C.  16-byte requests when
// the rows allow it (C % 4 == 0, aligned base), 4-byte requests otherwise.
__device__ __forceinline__ void load_chunk_async(float* dst, const float* __restrict__ src, int r0, int nr_pad, int R, int C,
                                                 int c0, int tid) {
    const bool vec = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        for (int e = tid; e < nr_pad * (kChunk / 4); e += kThreads) {
            const int r = e / (kChunk / 4), c4 = e % (kChunk / 4);
            const int gr = r0 + r, gc = c0 + 4 * c4;
            const bool ok = gr < R && gc < C;
            cp_async_zfill(dst + r * kPitch + 4 * c4, ok ? src + (size_t)gr * C + gc : src, 16, ok ? 16 : 0);
        }
    } else {
        for (int e = tid; e < nr_pad * kChunk; e += kThreads) {
            const int r = e / kChunk, c = e % kChunk;
            const int gr = r0 + r, gc = c0 + c;
            const bool ok = gr < R && gc < C;
            cp_async_zfill(dst + r * kPitch + c, ok ? src + (size_t)gr * C + gc : src, 4, ok ? 4 : 0);
        }
    }
}

// y[m][o] = b[o] + sum_c x[m][c] W[o][c].  The 256 threads are four groups of 64; group kg takes the float4 columns
// c4 = kg, kg + 4, ... of every chunk (a 4-way split of the contraction), and inside a group thread (rg = t / 8, og = t % 8)
// owns the rows rg + 8 i (i < 8) x the outputs og + 8 j: 8 + NJ shared-memory loads per 32 NJ FMA.  (The first version
// gave every thread 2 rows x NJ outputs of the whole contraction: 2 + NJ loads per 8 NJ FMA -- bound by shared-memory
// wavefronts, 18 us at [8192 x 384] -> 40.)  The four partial sums meet in shared memory at the end, in a fixed order.
template <int NJ>
__global__ void __launch_bounds__(kThreads)
    linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                      float* __restrict__ y, int M, int C, int O) {
    extern __shared__ __align__(16) float sm[];
    constexpr int kBuf = (8 * NJ + kRows) * kPitch;  // one stage: W chunk [8 NJ][kPitch] + x chunk [kRows][kPitch]
    const int tid = threadIdx.x, kg = tid >> 6, t64 = tid & 63, og = t64 & 7, rg = t64 >> 3;
    const int m0 = blockIdx.x * kRows;
    const int nchunks = (C + kChunk - 1) / kChunk;
    float acc[8][NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
    auto issue = [&](int k) {
        float* ws = sm + (k & 1) * kBuf;
        load_chunk_async(ws, W, 0, 8 * NJ, O, C, k * kChunk, tid);
        load_chunk_async(ws + 8 * NJ * kPitch, x, m0, kRows, M, C, k * kChunk, tid);
        cp_async_commit();
    };
    issue(0);
    for (int k = 0; k < nchunks; ++k) {
        if (k + 1 < nchunks) {
            issue(k + 1);          // the other stage was released by the barrier that ended chunk k - 1
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* ws = sm + (k & 1) * kBuf;
        const float* xs = ws + 8 * NJ * kPitch;
        const int n4 = (min(kChunk, C - k * kChunk) + 3) / 4;
        for (int c4 = kg; c4 < n4; c4 += 4) {
            float4 w[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) w[j] = *reinterpret_cast<const float4*>(ws + (og + 8 * j) * kPitch + 4 * c4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (rg + 8 * i) * kPitch + 4 * c4);
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[i][j] = dot4(xv, w[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    // the four groups' partial sums: red[(kg * 8 NJ + e) * 64 + t64], e = i * NJ + j (stride 64: conflict-free), added in
    // group order by the thread that writes the output
    float* red = sm;  // 4 * 8 NJ * 64 floats <= two stages (every stage was released by the last barrier)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) red[(kg * 8 * NJ + i * NJ + j) * 64 + t64] = acc[i][j];
    __syncthreads();
    for (int q = tid; q < 8 * NJ * 64; q += kThreads) {
        const int e = q >> 6, t = q & 63;
        const int i = e / NJ, j = e - i * NJ;
        const int m = m0 + (t >> 3) + 8 * i, o = (t & 7) + 8 * j;
        if (m >= M || o >= O) continue;
        const float s = ((red[q] + red[q + 8 * NJ * 64]) + red[q + 2 * 8 * NJ * 64]) + red[q + 3 * 8 * NJ * 64];
        y[(size_t)m * O + o] = s + __ldg(b + o);
    }
}

// One CTA = kRows rows: g_x[rows] = g W (written), and this row block's share of g_W = g^T x, g_b = column sums of g
// (partial sums: part[cta][O * C + O], folded by linear_fold_kernel).
template <int NJ>
__global__ void __launch_bounds__(kThreads)
    linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ g,
                      float* __restrict__ g_x, float* __restrict__ part, int M, int C, int O) {
    constexpr int OP = 8 * NJ;            // padded out features
    constexpr int GP = OP + 4;
    constexpr int kBuf = (OP + kRows) * kPitch;  // one stage: W chunk [OP][kPitch] + x chunk [kRows][kPitch]
    extern __shared__ __align__(16) float sm[];
    float* gs = sm + 2 * kBuf;            // [kRows][GP]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * kRows;
    const int nchunks = (C + kChunk - 1) / kChunk;
    float* mypart = part + (size_t)blockIdx.x * ((size_t)O * C + O);
    auto issue = [&](int k) {
        float* ws = sm + (k & 1) * kBuf;
        load_chunk_async(ws, W, 0, OP, O, C, k * kChunk, tid);
        load_chunk_async(ws + OP * kPitch, x, m0, kRows, M, C, k * kChunk, tid);
        cp_async_commit();
    };
    for (int e = tid; e < kRows * OP; e += kThreads) {  // the block's rows of g (with chunk 0: the first commit group)
        const int r = e / OP, o = e % OP;
        const bool ok = m0 + r < M && o < O;
        cp_async_zfill(gs + r * GP + o, ok ? g + (size_t)(m0 + r) * O + o : g, 4, ok ? 4 : 0);
    }
    issue(0);
    cp_async_wait<0>();
    __syncthreads();
    if (tid < O) {  // bias gradient: column sums over this block's rows
        float s = 0.f;
        for (int r = 0; r < kRows; ++r) s += gs[r * GP + tid];
        mypart[(size_t)O * C + tid] = s;
    }
    const int c4 = tid & 31, rgrp = tid >> 5;    // (1): rows rgrp + 8 i, columns 4 c4 .. 4 c4 + 3 of the chunk
    const int cw = tid & 127, oh = tid >> 7;     // (2): column cw of the chunk, out features [oh * OP / 2, +OP / 2)
    for (int k = 0; k < nchunks; ++k) {
        const int c0 = k * kChunk;
        if (k + 1 < nchunks) {
            issue(k + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* ws = sm + (k & 1) * kBuf;
        const float* xs = ws + OP * kPitch;
        if (g_x != nullptr && c0 + 4 * c4 < C) {  // (1) g_x = g W
            float4 acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int o4 = 0; o4 < OP / 4; ++o4) {
                float4 w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(ws + (4 * o4 + k) * kPitch + 4 * c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 gv = *reinterpret_cast<const float4*>(gs + (rgrp + 8 * i) * GP + 4 * o4);
                    acc[i].x = fmaf(gv.x, w[0].x, acc[i].x); acc[i].y = fmaf(gv.x, w[0].y, acc[i].y);
                    acc[i].z = fmaf(gv.x, w[0].z, acc[i].z); acc[i].w = fmaf(gv.x, w[0].w, acc[i].w);
                    acc[i].x = fmaf(gv.y, w[1].x, acc[i].x); acc[i].y = fmaf(gv.y, w[1].y, acc[i].y);
                    acc[i].z = fmaf(gv.y, w[1].z, acc[i].z); acc[i].w = fmaf(gv.y, w[1].w, acc[i].w);
                    acc[i].x = fmaf(gv.z, w[2].x, acc[i].x); acc[i].y = fmaf(gv.z, w[2].y, acc[i].y);
                    acc[i].z = fmaf(gv.z, w[2].z, acc[i].z); acc[i].w = fmaf(gv.z, w[2].w, acc[i].w);
                    acc[i].x = fmaf(gv.w, w[3].x, acc[i].x); acc[i].y = fmaf(gv.w, w[3].y, acc[i].y);
                    acc[i].z = fmaf(gv.w, w[3].z, acc[i].z); acc[i].w = fmaf(gv.w, w[3].w, acc[i].w);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + rgrp + 8 * i;
                if (m >= M) continue;
                float* p = g_x + (size_t)m * C + c0 + 4 * c4;
                const float v[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + 4 * c4 + k < C) p[k] = v[k];
            }
        }
        {  // (2) this block's share of g_W = g^T x
            float acc[OP / 2];
#pragma unroll
            for (int q = 0; q < OP / 2; ++q) acc[q] = 0.f;
#pragma unroll 4
            for (int r = 0; r < kRows; ++r) {
                const float xv = xs[r * kPitch + cw];
#pragma unroll
                for (int k4 = 0; k4 < OP / 8; ++k4) {
                    const float4 gv = *reinterpret_cast<const float4*>(gs + r * GP + oh * (OP / 2) + 4 * k4);
                    acc[4 * k4] = fmaf(gv.x, xv, acc[4 * k4]);
                    acc[4 * k4 + 1] = fmaf(gv.y, xv, acc[4 * k4 + 1]);
                    acc[4 * k4 + 2] = fmaf(gv.z, xv, acc[4 * k4 + 2]);
                    acc[4 * k4 + 3] = fmaf(gv.w, xv, acc[4 * k4 + 3]);
                }
            }
            if (c0 + cw < C) {
#pragma unroll
                for (int q = 0; q < OP / 2; ++q) {
                    const int o = oh * (OP / 2) + q;
                    if (o < O) mypart[(size_t)o * C + c0 + cw] = acc[q];
                }
            }
        }
        __syncthreads();  // stage k & 1 is refilled by the next iteration's issue
    }
}

// g_W[o][c], g_b[o] = sum over the row blocks.  Block = 32 outputs x 8 warps: warp w sums the row blocks w, w + 8, ...
// (coalesced over the outputs), the eight partial sums are added in warp order: deterministic.
__global__ void __launch_bounds__(256)
    linear_fold_kernel(const float* __restrict__ part, float* __restrict__ g_W, float* __restrict__ g_b, int nblk, int C, int O) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const int n = O * C + O;
    float s0 = 0.f, s1 = 0.f;
    if (e < n) {
        int k = warp;
        for (; k + 8 < nblk; k += 16) {
            s0 += __ldg(part + (size_t)k * n + e);
            s1 += __ldg(part + (size_t)(k + 8) * n + e);
        }
        if (k < nblk) s0 += __ldg(part + (size_t)k * n + e);
    }
    red[warp][lane] = s0 + s1;
    __syncthreads();
    if (warp == 0 && e < n) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][lane];
        if (e < O * C) {
            if (g_W != nullptr) g_W[e] = s;
        } else if (g_b != nullptr) {
            g_b[e - O * C] = s;
        }
    }
}

constexpr int kCMax = 16;  // widest composed transform (columns of W1) served

// Wc[r][c] = sum_o W2[r][o] W1[o][c],  bc[r] = sum_o W2[r][o] b1[o] + b2[r]: one warp per row r
__global__ void __launch_bounds__(256)
    compose_fwd_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                       const float* __restrict__ b2, float* __restrict__ Wc, float* __restrict__ bc, int R, int O, int C) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= R) return;
    float acc[kCMax + 1];
#pragma unroll
    for (int c = 0; c <= kCMax; ++c) acc[c] = 0.f;
    for (int o = lane; o < O; o += 32) {
        const float w = __ldg(W2 + (size_t)r * O + o);
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
            if (c < C) acc[c] = fmaf(w, __ldg(W1 + (size_t)o * C + c), acc[c]);
        acc[kCMax] = fmaf(w, __ldg(b1 + o), acc[kCMax]);
    }
#pragma unroll
    for (int c = 0; c <= kCMax; ++c) {
        if (c < C || c == kCMax) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], s);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
            if (c < C) Wc[(size_t)r * C + c] = acc[c];
        bc[r] = acc[kCMax] + __ldg(b2 + r);
    }
}

// blocks [0, nA):  g_W2[r][o] = sum_c g_Wc[r][c] W1[o][c] + g_bc[r] b1[o]           (element-wise over r, o)
// blocks [nA, ..): g_W1[o][c] = sum_r W2[r][o] g_Wc[r][c],  g_b1[o] = sum_r W2[r][o] g_bc[r]
//                  (32 columns o per block, 8 warps over r, fixed-order reduction through shared memory)
constexpr int kCbThreads = 512;
__global__ void __launch_bounds__(kCbThreads)
    compose_bwd_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                       const float* __restrict__ g_Wc, const float* __restrict__ g_bc, float* __restrict__ g_W1,
                       float* __restrict__ g_b1, float* __restrict__ g_W2, float* __restrict__ g_b2, int R, int O, int C,
                       int nA) {
    __shared__ float red[kCbThreads / 32][kCMax + 1][32];
    constexpr int NW = kCbThreads / 32;
    if ((int)blockIdx.x < nA) {
        const long long e = (long long)blockIdx.x * kCbThreads + threadIdx.x;
        if (e >= (long long)R * O) return;
        const int r = (int)(e / O), o = (int)(e % O);
        if (g_b2 != nullptr && o == 0) g_b2[r] = __ldg(g_bc + r);  // dL/db2 = dL/dbc
        float s = __ldg(g_bc + r) * __ldg(b1 + o);
#pragma unroll
        for (int c = 0; c < kCMax; ++c)
            if (c < C) s = fmaf(__ldg(g_Wc + (size_t)r * C + c), __ldg(W1 + (size_t)o * C + c), s);
        g_W2[e] = s;
        return;
    }
    const int o = ((int)blockIdx.x - nA) * 32 + (threadIdx.x & 31), warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[kCMax + 1];
#pragma unroll
    for (int c = 0; c <= kCMax; ++c) acc[c] = 0.f;
    if (o < O) {
        int r = warp;
        for (; r + 3 * NW < R; r += 4 * NW) {  // four rows per round: their loads are in flight together
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = __ldg(W2 + (size_t)(r + u * NW) * O + o);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int c = 0; c < kCMax; ++c)
                    if (c < C) acc[c] = fmaf(w[u], __ldg(g_Wc + (size_t)(r + u * NW) * C + c), acc[c]);
                acc[kCMax] = fmaf(w[u], __ldg(g_bc + r + u * NW), acc[kCMax]);
            }
        }
        for (; r < R; r += NW) {
            const float w = __ldg(W2 + (size_t)r * O + o);
#pragma unroll
            for (int c = 0; c < kCMax; ++c)
                if (c < C) acc[c] = fmaf(w, __ldg(g_Wc + (size_t)r * C + c), acc[c]);
            acc[kCMax] = fmaf(w, __ldg(g_bc + r), acc[kCMax]);
        }
    }
#pragma unroll
    for (int c = 0; c <= kCMax; ++c) red[warp][c][lane] = acc[c];
    __syncthreads();
    for (int e = threadIdx.x; e < (kCMax + 1) * 32; e += kCbThreads) {
        const int c = e >> 5, l = e & 31, oo = ((int)blockIdx.x - nA) * 32 + l;
        if (oo >= O || (c >= C && c != kCMax)) continue;
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < NW; ++w) s += red[w][c][l];
        if (c == kCMax) g_b1[oo] = s;
        else g_W1[(size_t)oo * C + c] = s;
    }
}

template <int NJ>
static int launch_fwd(const float* x, const float* W, const float* b, float* y, int M, int C, int O, cudaStream_t st) {
    const size_t smem = 2 * (size_t)(8 * NJ + kRows) * kPitch * sizeof(float);
    static PerDeviceInt configured;
    if (std::atomic<int>* slot = configured.slot(); !slot || slot->load(std::memory_order_relaxed) < 0) {
        cudaError_t e = cudaFuncSetAttribute(linear_fwd_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (slot) slot->store(1, std::memory_order_relaxed);
    }
    linear_fwd_kernel<NJ><<<(M + kRows - 1) / kRows, kThreads, smem, st>>>(x, W, b, y, M, C, O);
    return 0;
}

template <int NJ>
static int launch_bwd(const float* x, const float* W, const float* g, float* g_x, float* part, int M, int C, int O,
                      cudaStream_t st) {
    const size_t smem = (2 * (size_t)(8 * NJ + kRows) * kPitch + (size_t)kRows * (8 * NJ + 4)) * sizeof(float);
    static PerDeviceInt configured;
    if (std::atomic<int>* slot = configured.slot(); !slot || slot->load(std::memory_order_relaxed) < 0) {
        cudaError_t e = cudaFuncSetAttribute(linear_bwd_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (slot) slot->store(1, std::memory_order_relaxed);
    }
    linear_bwd_kernel<NJ><<<(M + kRows - 1) / kRows, kThreads, smem, st>>>(x, W, g, g_x, part, M, C, O);
    return 0;
}

}  // namespace lin
}  // namespace qec

using namespace qec;

// 1 if qec_linear_fwd / _bwd serve a layer with C in and O out features
extern "C" int qec_linear_supported(int C, int O) { return (C >= 1 && O >= 1 && O <= 64) ? 1 : 0; }

// floats of workspace qec_linear_bwd needs
extern "C" long long qec_linear_bwd_workspace(int M, int C, int O) {
    if (!qec_linear_supported(C, O) || M < 1) return 0;
    return (long long)((M + lin::kRows - 1) / lin::kRows) * ((long long)O * C + O);
}

// y [M, O] = x [M, C] W^T + b     (W [O, C], b [O]; reference nn.Linear, qec_module.py:123)
extern "C" int qec_linear_fwd(const float* x, const float* W, const float* b, float* y, int M, int C, int O, void* stream) {
    QEC_CHECK_ARG(x, 1);
    QEC_CHECK_ARG(W, 2);
    QEC_CHECK_ARG(b, 3);
    QEC_CHECK_ARG(y, 4);
    QEC_CHECK_ARG(M >= 1, 5);
    QEC_CHECK_ARG(qec_linear_supported(C, O), 6);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = 0;
    switch ((O + 7) / 8) {
        case 1: rc = lin::launch_fwd<1>(x, W, b, y, M, C, O, st); break;
        case 2: rc = lin::launch_fwd<2>(x, W, b, y, M, C, O, st); break;
        case 3: rc = lin::launch_fwd<3>(x, W, b, y, M, C, O, st); break;
        case 4: rc = lin::launch_fwd<4>(x, W, b, y, M, C, O, st); break;
        case 5: rc = lin::launch_fwd<5>(x, W, b, y, M, C, O, st); break;
        case 6: rc = lin::launch_fwd<6>(x, W, b, y, M, C, O, st); break;
        case 7: rc = lin::launch_fwd<7>(x, W, b, y, M, C, O, st); break;
        default: rc = lin::launch_fwd<8>(x, W, b, y, M, C, O, st); break;
    }
    if (rc) return rc;
    return launch_status();
}

// g [M, O] -> g_x [M, C] | NULL, g_W [O, C] | NULL, g_b [O] | NULL.  workspace: qec_linear_bwd_workspace floats.
extern "C" int qec_linear_bwd(const float* x, const float* W, const float* g, float* g_x, float* g_W, float* g_b,
                              float* workspace, int M, int C, int O, void* stream) {
    QEC_CHECK_ARG(x, 1);
    QEC_CHECK_ARG(W, 2);
    QEC_CHECK_ARG(g, 3);
    QEC_CHECK_ARG(workspace, 7);
    QEC_CHECK_ARG(M >= 1, 8);
    QEC_CHECK_ARG(qec_linear_supported(C, O), 9);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = 0;
    switch ((O + 7) / 8) {
        case 1: rc = lin::launch_bwd<1>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 2: rc = lin::launch_bwd<2>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 3: rc = lin::launch_bwd<3>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 4: rc = lin::launch_bwd<4>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 5: rc = lin::launch_bwd<5>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 6: rc = lin::launch_bwd<6>(x, W, g, g_x, workspace, M, C, O, st); break;
        case 7: rc = lin::launch_bwd<7>(x, W, g, g_x, workspace, M, C, O, st); break;
        default: rc = lin::launch_bwd<8>(x, W, g, g_x, workspace, M, C, O, st); break;
    }
    if (rc) return rc;
    if (g_W != nullptr || g_b != nullptr) {
        const int n = O * C + O;
        lin::linear_fold_kernel<<<(n + 31) / 32, 256, 0, st>>>(workspace, g_W, g_b, (M + lin::kRows - 1) / lin::kRows, C, O);
    }
    return launch_status();
}

// 1 if qec_compose_fwd / _bwd serve W1 with C columns
extern "C" int qec_compose_supported(int C) { return (C >= 1 && C <= lin::kCMax) ? 1 : 0; }

// Wc [R, C] = W2 [R, O] W1 [O, C],  bc [R] = W2 b1 + b2   (the two Linears of get_votes composed, qec_module.py:123-124)
extern "C" int qec_compose_fwd(const float* W1, const float* b1, const float* W2, const float* b2, float* Wc, float* bc,
                               int R, int O, int C, void* stream) {
    QEC_CHECK_ARG(W1, 1);
    QEC_CHECK_ARG(b1, 2);
    QEC_CHECK_ARG(W2, 3);
    QEC_CHECK_ARG(b2, 4);
    QEC_CHECK_ARG(Wc, 5);
    QEC_CHECK_ARG(bc, 6);
    QEC_CHECK_ARG(R >= 1 && O >= 1, 7);
    QEC_CHECK_ARG(qec_compose_supported(C), 9);
    lin::compose_fwd_kernel<<<(R + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(W1, b1, W2, b2, Wc, bc, R, O, C);
    return launch_status();
}

// adjoint: (g_Wc [R, C], g_bc [R]) -> g_W1 [O, C], g_b1 [O], g_W2 [R, O], g_b2 [R] | NULL (dL/db2 = g_bc: a copy, for
// callers that want it somewhere else, e.g. inside a flat gradient buffer)
extern "C" int qec_compose_bwd(const float* W1, const float* b1, const float* W2, const float* g_Wc, const float* g_bc,
                               float* g_W1, float* g_b1, float* g_W2, float* g_b2, int R, int O, int C, void* stream) {
    QEC_CHECK_ARG(W1, 1);
    QEC_CHECK_ARG(b1, 2);
    QEC_CHECK_ARG(W2, 3);
    QEC_CHECK_ARG(g_Wc, 4);
    QEC_CHECK_ARG(g_bc, 5);
    QEC_CHECK_ARG(g_W1, 6);
    QEC_CHECK_ARG(g_b1, 7);
    QEC_CHECK_ARG(g_W2, 8);
    QEC_CHECK_ARG(R >= 1 && O >= 1, 10);
    QEC_CHECK_ARG(qec_compose_supported(C), 12);
    const int nA = (int)(((long long)R * O + lin::kCbThreads - 1) / lin::kCbThreads), nB = (O + 31) / 32;
    lin::compose_bwd_kernel<<<nA + nB, lin::kCbThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        W1, b1, W2, g_Wc, g_bc, g_W1, g_b1, g_W2, g_b2, R, O, C, nA);
    return launch_status();
}
