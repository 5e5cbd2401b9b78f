// adam_math.cuh -- the element update of qec_adam_flat (adam.cu) and its bias corrections, host/device like
// qec_math.cuh so that tests/host/ can check the arithmetic against torch.optim.Adam on the CPU-only build
// container.  Reference: train_cls.py:44-51 (torch.optim.Adam, default betas / eps).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define QEC_ADAM_HD __host__ __device__ __forceinline__
#else
#define QEC_ADAM_HD inline
#endif

namespace qec {

struct AdamHyper {
    double beta1_d, beta2_d;  // for the bias corrections, which torch evaluates in double
    float lr, beta1, beta2, omb1, omb2, eps, weight_decay, grad_scale;  // omb = 1 - beta, rounded from double
    int zero_grad;
};

QEC_ADAM_HD AdamHyper adam_hyper(float lr, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                                 int zero_grad) {
    AdamHyper h;
    h.beta1_d = beta1; h.beta2_d = beta2;
    h.lr = lr; h.beta1 = (float)beta1; h.beta2 = (float)beta2;
    h.omb1 = (float)(1.0 - beta1); h.omb2 = (float)(1.0 - beta2);
    h.eps = eps; h.weight_decay = weight_decay; h.grad_scale = grad_scale; h.zero_grad = zero_grad;
    return h;
}

// 1 / (1 - beta1^t) and 1 / sqrt(1 - beta2^t) for step t = 1, 2, ...: in double like torch.optim.Adam
// (1 - 0.999f loses five digits at t = 1)
QEC_ADAM_HD void adam_bias_corrections(const AdamHyper& h, int t, float& inv_bc1, float& inv_sqrt_bc2) {
    inv_bc1 = (float)(1.0 / (1.0 - pow(h.beta1_d, (double)t)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(h.beta2_d, (double)t)));
}

QEC_ADAM_HD void adam_one(float& p, float& g, float& m, float& v, const AdamHyper& h, float inv_bc1, float inv_sqrt_bc2) {
    float gr = g * h.grad_scale;
    if (h.weight_decay != 0.f) gr = fmaf(h.weight_decay, p, gr);  // L2 penalty folded into the gradient (torch.optim.Adam)
    m = fmaf(h.beta1, m, h.omb1 * gr);
    v = fmaf(h.beta2, v, h.omb2 * gr * gr);
    const float denom = fmaf(sqrtf(v), inv_sqrt_bc2, h.eps);
    p -= (h.lr * inv_bc1) * (m / denom);
    if (h.zero_grad) g = 0.f;
}

}  // namespace qec
