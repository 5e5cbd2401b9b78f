// gather.cu -- the loader-side neighbour gather (reference
// my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160; Siamese
// variant ..._sia.py:125-167), moved onto the device and batched.
//
//   idx [B,P,K] int64: original point ids, -1 = empty slot
//   act = [idx > 0], zeroed where idx is in wrong_ids for the first pool_size rows,
//         pool_size = #rows whose first slot is active          (:139-150)
//   the LAST row of point_set / lrf_set reads as zero and negative ids index
//   from the end, so -1 picks that zero row                      (:154-160)
// Pure copies and integer compares: bit-exact by construction.  One CTA per
// sample; every (p,k) slot is an independent 12/16-byte row read.
#include "qec_common.cuh"

namespace qec {

constexpr int kGT = 256;

__global__ void __launch_bounds__(kGT)
    gather_kernel(const float* __restrict__ point_set, const float* __restrict__ lrf_set,
                  const long long* __restrict__ idx, const long long* __restrict__ wrong_ids,
                  const int* __restrict__ wrong_count, float* __restrict__ points, float* __restrict__ lrfs,
                  float* __restrict__ act, int* __restrict__ err, int N, int P, int K, int Wmax) {
    const int b = blockIdx.x;
    const long long* ib = idx + (size_t)b * P * K;
    // pool_size: rows whose first slot is active
    int cnt = 0;
    for (int p0 = 0; p0 < P; p0 += kGT) {
        const int p = p0 + threadIdx.x;
        cnt += __syncthreads_count(p < P && ib[(size_t)p * K] > 0);
    }
    const int W = (wrong_ids != nullptr) ? min(wrong_count ? wrong_count[b] : Wmax, Wmax) : 0;
    const long long* wb = (wrong_ids != nullptr) ? wrong_ids + (size_t)b * Wmax : nullptr;
    const float* ps = point_set + (size_t)b * N * 3;
    const float* ls = lrf_set + (size_t)b * N * 4;
    for (int e = threadIdx.x; e < P * K; e += kGT) {
        const long long id = ib[e];
        float a = (id > 0) ? 1.f : 0.f;
        if (a != 0.f && e / K < cnt) {
            for (int k = 0; k < W; ++k)
                if (wb[k] == id) a = 0.f;
        }
        long long row = id < 0 ? id + N : id;
        if (row < 0 || row >= N) {  // the reference would raise IndexError
            atomicExch(err, 1);
            row = N - 1;
        }
        const bool zero = row == N - 1;
        const size_t o = (size_t)b * P * K + e;
        const float* sp = ps + (size_t)row * 3;
        points[o * 3 + 0] = zero ? 0.f : sp[0];
        points[o * 3 + 1] = zero ? 0.f : sp[1];
        points[o * 3 + 2] = zero ? 0.f : sp[2];
        const float4 q = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(ls + (size_t)row * 4);
        *reinterpret_cast<float4*>(lrfs + o * 4) = q;
        act[o] = a;
    }
}

}  // namespace qec

using namespace qec;

extern "C" int qec_gather(const float* point_set, const float* lrf_set, const long long* idx,
                          const long long* wrong_ids, const int* wrong_count, float* points, float* lrfs, float* act,
                          int* err, int B, int N, int P, int K, int Wmax, void* stream) {
    QEC_CHECK_ARG(point_set, 1);
    QEC_CHECK_ARG(lrf_set, 2);
    QEC_CHECK_ARG(idx, 3);
    QEC_CHECK_ARG(points && lrfs && act, 6);
    QEC_CHECK_ARG(err, 9);
    QEC_CHECK_ARG(B > 0, 10);
    QEC_CHECK_ARG(N > 0, 11);
    QEC_CHECK_ARG(P > 0, 12);
    QEC_CHECK_ARG(K > 0, 13);
    QEC_CHECK_ARG(Wmax >= 0, 14);
    gather_kernel<<<B, kGT, 0, static_cast<cudaStream_t>(stream)>>>(point_set, lrf_set, idx, wrong_ids, wrong_count,
                                                                    points, lrfs, act, err, N, P, K, Wmax);
    return launch_status();
}
