// knn.cu -- deterministic K-nearest-neighbour indices of the patch centres, on the device
// (the on-device counterpart of the neighbourhoods the reference reads from its
// offline pooling-index files, my_dataloader/gen_downsample_m40.py:86-150 -> the idx tensors consumed
// by my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160; north_star (d): "kNN indices
// bit-exact").
//
// Contract, chosen so that the result is bit-exact against a plain float32 CPU statement
// (oracle/qec_oracle.py: knn_indices):
//   d(p, n) = ((cx - x_n)^2 + (cy - y_n)^2) + (cz - z_n)^2      each operation rounded to float32
//             separately (no FMA contraction), in this order;
//   neighbours sorted by (d, n) ascending -- ties go to the smaller point index (a stable sort);
//   idx_out [B, P, K] int64.  The centre itself (d = 0) comes first when it is one of the points.
// One warp per (b, p): every lane scans N/32 points keeping its K best in a sorted list in shared
// memory, then the 32 lists are merged with K warp-wide argmin rounds.  Integer/compare work on
// ~25 KB per sample: latency-bound, a few microseconds per batch.
#include "qec_common.cuh"

namespace qec {

constexpr int kKnnWarps = 4;

__device__ __forceinline__ bool knn_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

__global__ void __launch_bounds__(kKnnWarps * 32)
    knn_kernel(const float* __restrict__ points, const float* __restrict__ centres, const long long* __restrict__ centre_idx,
               long long* __restrict__ idx_out, int* __restrict__ err, int B, int N, int P, int K) {
    extern __shared__ unsigned char knn_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* ld = reinterpret_cast<float*>(knn_smem) + (size_t)(warp * 32 + lane) * K;          // this lane's distances
    int* li = reinterpret_cast<int*>(knn_smem + (size_t)kKnnWarps * 32 * K * sizeof(float)) + (size_t)(warp * 32 + lane) * K;
    const long long total = (long long)B * P;
    for (long long bp = (long long)blockIdx.x * kKnnWarps + warp; bp < total; bp += (long long)gridDim.x * kKnnWarps) {
        const long long b = bp / P;
        const float* pts = points + (size_t)b * N * 3;
        float cx, cy, cz;
        if (centre_idx != nullptr) {
            long long ci = centre_idx[bp];
            if (ci < 0 || ci >= N) {
                if (lane == 0) atomicExch(err, 1);
                ci = 0;
            }
            cx = pts[ci * 3]; cy = pts[ci * 3 + 1]; cz = pts[ci * 3 + 2];
        } else {
            cx = centres[bp * 3]; cy = centres[bp * 3 + 1]; cz = centres[bp * 3 + 2];
        }
        int cnt = 0;  // entries of this lane's list
        for (int n = lane; n < N; n += 32) {
            const float dx = __fsub_rn(cx, pts[n * 3]), dy = __fsub_rn(cy, pts[n * 3 + 1]), dz = __fsub_rn(cz, pts[n * 3 + 2]);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (cnt < K || knn_less(d, n, ld[cnt - 1], li[cnt - 1])) {  // insertion into the sorted list
                int pos = (cnt < K) ? cnt : K - 1;
                while (pos > 0 && knn_less(d, n, ld[pos - 1], li[pos - 1])) {
                    ld[pos] = ld[pos - 1];
                    li[pos] = li[pos - 1];
                    --pos;
                }
                ld[pos] = d;
                li[pos] = n;
                if (cnt < K) ++cnt;
            }
        }
        __syncwarp();
        int head = 0;
        for (int r = 0; r < K; ++r) {  // merge: K rounds of warp-wide argmin over the list heads
            float d = (head < cnt) ? ld[head] : __int_as_float(0x7f800000);
            int i = (head < cnt) ? li[head] : 0x7fffffff;
            int who = lane;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                const float d2 = __shfl_xor_sync(0xffffffffu, d, s);
                const int i2 = __shfl_xor_sync(0xffffffffu, i, s);
                const int w2 = __shfl_xor_sync(0xffffffffu, who, s);
                if (knn_less(d2, i2, d, i)) { d = d2; i = i2; who = w2; }
            }
            if (lane == 0) idx_out[bp * K + r] = (i == 0x7fffffff) ? -1 : (long long)i;  // fewer than K points: -1
            if (who == lane) ++head;
        }
        __syncwarp();
    }
}

}  // namespace qec

// points [B, N, 3]; the P centres per sample are either point indices (centre_idx [B, P] int64) or
// coordinates (centres [B, P, 3]) -- exactly one of the two is non-null.  idx_out [B, P, K] int64, K <= 64.
// err: device int set to 1 if a centre index is outside [0, N).
extern "C" int qec_knn(const float* points, const float* centres, const long long* centre_idx, long long* idx_out,
                       int* err, int B, int N, int P, int K, void* stream) {
    QEC_CHECK_ARG(points, 1);
    QEC_CHECK_ARG((centres != nullptr) != (centre_idx != nullptr), 2);
    QEC_CHECK_ARG(idx_out && err, 4);
    QEC_CHECK_ARG(B > 0 && N > 0 && P > 0, 6);
    QEC_CHECK_ARG(K > 0 && K <= 64, 9);
    const size_t smem = (size_t)qec::kKnnWarps * 32 * K * (sizeof(float) + sizeof(int));
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(qec::knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const long long total = (long long)B * P;
    long long blocks = (total + qec::kKnnWarps - 1) / qec::kKnnWarps;
    if (blocks > qec::kNumSMs * 8) blocks = qec::kNumSMs * 8;
    qec::knn_kernel<<<(unsigned)blocks, qec::kKnnWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        points, centres, centre_idx, idx_out, err, B, N, P, K);
    return qec::launch_status();
}
