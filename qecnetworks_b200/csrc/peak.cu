// peak.cu -- FP32 FMA roof of the device, measured in the same run as the kernels it
// is the denominator for ("the builder must run an FFMA
// microbenchmark").  Each thread runs 16 independent FMA chains so the fma pipe is
// issue-bound, not latency-bound.
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
}  // namespace qec

// Launches blocks*256 threads doing iters*16 FMAs each: 2*16*iters*256*blocks FLOP.
extern "C" int qec_ffma_peak(float* out, int blocks, int iters, void* stream) {
    QEC_CHECK_ARG(out, 1);
    QEC_CHECK_ARG(blocks > 0, 2);
    QEC_CHECK_ARG(iters > 0, 3);
    qec::ffma_peak_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, iters, 0.999f, 1e-3f);
    return qec::launch_status();
}
