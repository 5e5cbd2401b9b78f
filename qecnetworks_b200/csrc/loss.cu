// loss.cu -- spread loss of QEC-Net (reference models/qec_net.py:53-59) and its gradient in one launch:
//   loss = mean_b sum_{c != y_b} relu(m - (x[b, y_b] - x[b, c]))^2
// The reference's ATen sequence is ~25 launches of a few microseconds each on a [B, classes] tensor
// (one_hot scatter, mul, sum, relu, pow, ...), forward and backward; at 3 ms per training step that
// is a visible slice of the step.  One warp per sample, deterministic (fixed-order block sum).
#include "qec_common.cuh"

namespace qec {

__global__ void __launch_bounds__(256)
    spread_loss_kernel(const float* __restrict__ x, const long long* __restrict__ labels, float margin,
                       float* __restrict__ per_sample, float* __restrict__ g_x, int B, int C, int* __restrict__ err) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const long long y = labels[warp];
    if (y < 0 || y >= C) {  // the reference's scatter_ raises here
        if (lane == 0) {
            atomicExch(err, 1);
            per_sample[warp] = 0.f;
        }
        for (int c = lane; c < C; c += 32) g_x[(size_t)warp * C + c] = 0.f;
        return;
    }
    const float* row = x + (size_t)warp * C;
    const float at = __ldg(row + y);
    const float scale = 2.f / (float)B;
    float loss = 0.f, gsum = 0.f;
    for (int c = lane; c < C; c += 32) {
        if (c == (int)y) continue;
        const float d = fmaxf(margin - (at - __ldg(row + c)), 0.f);
        loss = fmaf(d, d, loss);
        const float g = scale * d;
        g_x[(size_t)warp * C + c] = g;
        gsum += g;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        loss += __shfl_xor_sync(0xffffffffu, loss, s);
        gsum += __shfl_xor_sync(0xffffffffu, gsum, s);
    }
    if (lane == 0) {
        per_sample[warp] = loss;
        g_x[(size_t)warp * C + y] = -gsum;
    }
}

__global__ void __launch_bounds__(256) mean_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
    __shared__ float sh[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += v[i];  // fixed order per thread
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0] / (float)n;
}

}  // namespace qec

// x [B, C] agreements, labels [B] int64 -> loss [1] (mean over the batch), g_x [B, C] = d loss / d x.
// scratch [B] floats.  err: device int set to 1 if a label is outside [0, C) (the reference raises).
extern "C" int qec_spread_loss(const float* x, const long long* labels, float margin, float* loss, float* g_x,
                               float* scratch, int* err, int B, int C, void* stream) {
    QEC_CHECK_ARG(x, 1);
    QEC_CHECK_ARG(labels, 2);
    QEC_CHECK_ARG(loss && g_x && scratch && err, 4);
    QEC_CHECK_ARG(B > 0 && C > 0, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)(((long long)B * 32 + 255) / 256);
    qec::spread_loss_kernel<<<grid, 256, 0, st>>>(x, labels, margin, scratch, g_x, B, C, err);
    qec::mean_kernel<<<1, 256, 0, st>>>(scratch, loss, B);
    return qec::launch_status();
}
