// sia_loss.cu -- the losses of the two QEC networks and the evaluation-side quaternion ops as single launches
// (SURVEY.md section 8 f.3 / f.4).
//
//  qec_class_pose_loss: spread loss of one branch (reference models/qec_net.py:53-59) or of the two Siamese
//      branches (models/qec_sia_net.py:69-77), plus -- optionally -- the pose-difference loss
//      (qec_sia_net.py:79-82 / qec_net.py:61-64) on the poses of the ground-truth class, which the reference
//      selects with a per-sample Python loop (train_cls_sia.py:72-78: pose_out_act[:, i] = pose_out[:, i, y_i]),
//      all with their gradients, in ONE launch of one CTA:
//          total = sum_s mean_b sum_{c != y_b} relu(m - (x[s,b,y_b] - x[s,b,c]))^2
//                  + w_pose * mean_b (2/pi) acos(min(|<pose[0,b,y_b], pose[1,b,y_b]>|, 0.9999))
//      The reference's ATen sequence for this is ~60 launches + B host-side indexing steps per training step.
//      Deterministic: every sum has a fixed order.
//  qec_quat_eval: q_distance and the relative pose of test_rot_sia.py:32-41, 131-132, batched.
#include "qec_common.cuh"

namespace qec {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum_fixed(float v, float* sh) {  // fixed-order tree: deterministic
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int k = kLossThreads / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    const float r = sh[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kLossThreads)
    class_pose_loss_kernel(const float* __restrict__ x, const long long* __restrict__ labels, float margin,
                           const float* __restrict__ pose, float w_pose, float* __restrict__ loss,
                           float* __restrict__ g_x, float* __restrict__ g_pose, int* __restrict__ err, int S, int B,
                           int C) {
    __shared__ float sh[kLossThreads];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kLossThreads / 32;
    const float scale = 2.f / (float)B;
    // ---- spread loss: one warp per (branch, sample), samples dealt round-robin (fixed order per warp)
    float spread = 0.f;
    for (int sb = warp; sb < S * B; sb += kWarps) {
        const int b = sb % B;
        const long long y = labels[b];
        float* grow = g_x + (size_t)sb * C;
        if (y < 0 || y >= C) {  // the reference's scatter_ / indexing raises here
            if (lane == 0) atomicExch(err, 1);
            for (int c = lane; c < C; c += 32) grow[c] = 0.f;
            continue;
        }
        const float* row = x + (size_t)sb * C;
        const float at = __ldg(row + y);
        float l = 0.f, gsum = 0.f;
        for (int c = lane; c < C; c += 32) {
            if (c == (int)y) continue;
            const float d = fmaxf(margin - (at - __ldg(row + c)), 0.f);
            l = fmaf(d, d, l);
            const float g = scale * d;
            grow[c] = g;
            gsum += g;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, s);
            gsum += __shfl_xor_sync(0xffffffffu, gsum, s);
        }
        if (lane == 0) {
            grow[y] = -gsum;
            spread += l;
        }
    }
    const float spread_tot = block_sum_fixed(lane == 0 ? spread : 0.f, sh) / (float)B;
    // ---- pose-difference loss on the ground-truth class of the two branches
    float pd = 0.f;
    if (pose != nullptr) {
        const size_t branch = (size_t)B * C * 4;
        // zero the gradient of every pose, then fill the two selected ones per sample
        if (g_pose != nullptr) {
            for (size_t i = threadIdx.x; i < 2 * branch; i += kLossThreads) g_pose[i] = 0.f;
            __syncthreads();
        }
        const float gscale = w_pose / (float)B;
        for (int b = threadIdx.x; b < B; b += kLossThreads) {
            const long long y = labels[b];
            if (y < 0 || y >= C) continue;  // already flagged above
            const size_t o = ((size_t)b * C + (size_t)y) * 4;
            const Quat p0 = ldq(pose + o), p1 = ldq(pose + branch + o);
            const float dot = qdot(p0, p1);
            pd += 0.63661977236758f * acosf(fminf(fabsf(dot), QEC_DCLAMP));
            if (g_pose != nullptr) {
                const float g = gscale * dist_grad(dot);  // d dist / d<p0,p1>, 0 beyond the clamp
                stq(g_pose + o, qscale(p1, g));
                stq(g_pose + branch + o, qscale(p0, g));
            }
        }
    }
    const float pd_tot = block_sum_fixed(pd, sh) / (float)B;
    if (threadIdx.x == 0) {
        loss[0] = spread_tot + w_pose * pd_tot;
        loss[1] = spread_tot;
        loss[2] = pd_tot;
    }
}

__global__ void __launch_bounds__(256)
    quat_eval_kernel(const float* __restrict__ p1, const float* __restrict__ p2, float* __restrict__ dist,
                     float* __restrict__ rel, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Quat a = ldq(p1 + i * 4), b = ldq(p2 + i * 4);
    if (dist != nullptr) {  // test_rot_sia.py:32-41: both poses sign-fixed, clamp at 1.0
        const float d = fminf(fabsf(qdot(qfix(a), qfix(b))), 1.f);
        dist[i] = 0.63661977236758f * acosf(d);
    }
    if (rel != nullptr) stq(rel + i * 4, qham(qconj(a), b));  // test_rot_sia.py:131-132
}

}  // namespace qec

// x [S, B, C] agreements of S = 1 (QecNet) or 2 (QecSiaNet) branches (S = 0, x = g_x = NULL: pose term only),
// labels [B] int64 (shared by the branches), pose [2, B, C, 4] or NULL (then no pose term; w_pose ignored).  -> loss [3] = {total, spread, pose_diff},
// g_x [S, B, C] = d total / d x, g_pose [2, B, C, 4] = d total / d pose (NULL to skip; zero off the label class).
// err: device int set to 1 if a label is outside [0, C).
extern "C" int qec_class_pose_loss(const float* x, const long long* labels, float margin, const float* pose,
                                   float w_pose, float* loss, float* g_x, float* g_pose, int* err, int S, int B, int C,
                                   void* stream) {
    QEC_CHECK_ARG(x || S == 0, 1);
    QEC_CHECK_ARG(labels, 2);
    QEC_CHECK_ARG(pose || S > 0, 4);
    QEC_CHECK_ARG(loss && (g_x || S == 0) && err, 6);
    QEC_CHECK_ARG(S >= 0 && S <= 2, 10);
    QEC_CHECK_ARG(B > 0 && C > 0, 11);
    qec::class_pose_loss_kernel<<<1, qec::kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, labels, margin, pose, w_pose, loss, g_x, g_pose, err, S, B, C);
    return qec::launch_status();
}

// pose1, pose2 [n, 4] -> dist [n] (NULL to skip) = (2/pi) acos(min(|<fix(p1), fix(p2)>|, 1)), rel [n, 4] (NULL to
// skip) = conj(pose1) (x) pose2   (reference test_rot_sia.py:32-41, 131-132)
extern "C" int qec_quat_eval(const float* pose1, const float* pose2, float* dist, float* rel, long long n,
                             void* stream) {
    QEC_CHECK_ARG(pose1, 1);
    QEC_CHECK_ARG(pose2, 2);
    QEC_CHECK_ARG(dist || rel, 3);
    QEC_CHECK_ARG(n > 0, 5);
    const unsigned grid = (unsigned)((n + 255) / 256);
    qec::quat_eval_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pose1, pose2, dist, rel, n);
    return qec::launch_status();
}
