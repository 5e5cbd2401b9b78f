// peak.cu -- FP32 FMA roof of the device, measured in the same run as the kernels it
// is the denominator for ("the builder must run an FFMA
// microbenchmark").  Each thread runs 16 independent FMA chains so the fma pipe is
// issue-bound, not latency-bound.
//
// qec_pipe_probe additionally measures how the packed FP32 instructions of sm_100
// (FFMA2 = fma.rn.f32x2) share the issue slot / fma pipe with scalar FFMA and with
// alu-pipe instructions (FMNMX): the fused routing kernels are issue-bound, so what a
// packed instruction costs decides their design.
#include "qec_common.cuh"

namespace qec {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[blockIdx.x * 256 + threadIdx.x] = s;  // never true: keeps the chains alive
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// NF2 packed FMA chains + NF scalar FMA chains + NA alternating min/max (FMNMX, alu pipe) chains per
// thread, all independent; one round = one instruction per chain.  Inline PTX per instruction and a
// non-unrolled outer loop of 8 rounds keep the instruction mix exactly as written.
template <int NF2, int NF, int NA>
__global__ void __launch_bounds__(256) pipe_probe_kernel(float* out, int iters, float a, float b) {
    u64 p[NF2 > 0 ? NF2 : 1];
    float x[NF > 0 ? NF : 1], y[NA > 0 ? NA : 1];
#pragma unroll
    for (int i = 0; i < NF2; ++i) p[i] = pk2((float)(threadIdx.x + i) * 1e-3f, (float)(threadIdx.x + i) * 2e-3f);
#pragma unroll
    for (int i = 0; i < NF; ++i) x[i] = (float)(threadIdx.x + i) * 3e-3f;
#pragma unroll
    for (int i = 0; i < NA; ++i) y[i] = (float)(threadIdx.x + i) * 4e-3f;
    const u64 aa = pk2(a, a), bb = pk2(b, b);
    const float c1 = a * 3.f, c2 = b * 5.f;
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < NF2) asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(aa), "l"(bb));
                if (i < NF) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                if (i < NA) {
                    if (r & 1) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c1));
                    else asm volatile("max.ftz.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(c2));
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NF2; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
#pragma unroll
    for (int i = 0; i < NF; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NA; ++i) s += y[i];
    if (s == 123.456f) out[blockIdx.x * 256 + threadIdx.x] = s;
}
}  // namespace qec

// Launches blocks*256 threads doing iters*16 FMAs each: 2*16*iters*256*blocks FLOP.
extern "C" int qec_ffma_peak(float* out, int blocks, int iters, void* stream) {
    QEC_CHECK_ARG(out, 1);
    QEC_CHECK_ARG(blocks > 0, 2);
    QEC_CHECK_ARG(iters > 0, 3);
    qec::ffma_peak_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, iters, 0.999f, 1e-3f);
    return qec::launch_status();
}

// mode 0: 16 FFMA | 1: 16 FFMA2 | 2: 8 FFMA + 8 FMNMX | 3: 8 FFMA2 + 8 FMNMX | 4: 8 FFMA2 + 8 FFMA | 5: 8 FFMA2
// | 6: 8 FFMA | 7: 16 FMNMX | 8: 16 FFMA + 8 FMNMX | 9: 8 FMNMX.  Each thread issues iters * (chains of the mode) instructions of each kind.
extern "C" int qec_pipe_probe(float* out, int blocks, int iters, int mode, void* stream) {
    QEC_CHECK_ARG(out, 1);
    QEC_CHECK_ARG(blocks > 0, 2);
    QEC_CHECK_ARG(iters > 0, 3);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float a = 0.999f, b = 1e-3f;
    switch (mode) {
        case 0: qec::pipe_probe_kernel<0, 16, 0><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 1: qec::pipe_probe_kernel<16, 0, 0><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 2: qec::pipe_probe_kernel<0, 8, 8><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 3: qec::pipe_probe_kernel<8, 0, 8><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 4: qec::pipe_probe_kernel<8, 8, 0><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 5: qec::pipe_probe_kernel<8, 0, 0><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 6: qec::pipe_probe_kernel<0, 8, 0><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 7: qec::pipe_probe_kernel<0, 0, 16><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 8: qec::pipe_probe_kernel<0, 16, 8><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        case 9: qec::pipe_probe_kernel<0, 0, 8><<<blocks, 256, 0, st>>>(out, iters, a, b); break;
        default: return -4;
    }
    return qec::launch_status();
}
