// adam.cu -- the optimizer step of the training loop (reference train_cls.py:44-51: torch.optim.Adam with
// default betas / eps, no weight decay unless asked) on the FLAT parameter and gradient buffers of the
// data-parallel plumbing (qecnetworks_b200/ddp.py).  QEC-Net has 0.92 M parameters in ~20 tensors, one of
// them 91 % of the total: torch's multi-tensor fused Adam needs 47 us for them inside the 2.7 ms step;
// one grid-stride launch over the flat buffers moves the same 26 MB (L2-resident) in a few microseconds,
// and it absorbs the two elementwise passes around it: the 1/world scaling of the all-reduced gradient
// and the zeroing of the gradient buffer for the next step.
//
// CUDA-graph friendly: the step counter lives on the device; the last block to finish advances it.
#include "adam_math.cuh"
#include "qec_common.cuh"

namespace qec {

// state: [0] = steps taken so far (as int), [1] = blocks finished in this launch
__global__ void __launch_bounds__(256)
    adam_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     long long n, AdamHyper h, int* __restrict__ state) {
    // bias corrections 1 - beta^t in double like torch.optim.Adam (1 - 0.999f loses five digits at t = 1):
    // one thread per block evaluates them, 2 x pow is ~1 us of latency
    __shared__ float bc[2];
    if (threadIdx.x == 0) adam_bias_corrections(h, state[0] + 1, bc[0], bc[1]);
    __syncthreads();
    const float inv_bc1 = bc[0], inv_sqrt_bc2 = bc[1];
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, h, inv_bc1, inv_sqrt_bc2);
        adam_one(pp.y, gg.y, mm.y, vv.y, h, inv_bc1, inv_sqrt_bc2);
        adam_one(pp.z, gg.z, mm.z, vv.z, h, inv_bc1, inv_sqrt_bc2);
        adam_one(pp.w, gg.w, mm.w, vv.w, h, inv_bc1, inv_sqrt_bc2);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (h.zero_grad) g4[i] = gg;
    }
    if (blockIdx.x == 0) {  // n not a multiple of four
        const long long i = (n4 << 2) + threadIdx.x;
        if (i < n) adam_one(p[i], g[i], m[i], v[i], h, inv_bc1, inv_sqrt_bc2);
    }
    // every block has read state[0] before it gets here: the last one to arrive advances the step
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(state + 1, 1);
        if (ticket == (int)gridDim.x - 1) {
            state[1] = 0;
            state[0] += 1;
        }
    }
}

}  // namespace qec

// One Adam step on flat FP32 buffers of n elements (16-byte aligned): the gradient is read as
// grad * grad_scale (1/world after a summing all-reduce), and overwritten with zeros when zero_grad != 0.
// state: two device ints, zero-initialised by the caller before the first step: {steps taken, scratch}.
extern "C" int qec_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                             double beta1, double beta2, float eps, float weight_decay, float grad_scale, int zero_grad,
                             int* state, void* stream) {
    QEC_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, 1);
    QEC_CHECK_ARG(n > 0, 5);
    QEC_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, 7);
    QEC_CHECK_ARG(state, 13);
    QEC_CHECK_ARG(((reinterpret_cast<size_t>(params) | reinterpret_cast<size_t>(grads) | reinterpret_cast<size_t>(exp_avg) |
                    reinterpret_cast<size_t>(exp_avg_sq)) & 15) == 0, 1);
    const qec::AdamHyper h = qec::adam_hyper(lr, beta1, beta2, eps, weight_decay, grad_scale, zero_grad);
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 4 * qec::kNumSMs) blocks = 4 * qec::kNumSMs;
    qec::adam_flat_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq,
                                                                                            n, h, state);
    return qec::launch_status();
}
