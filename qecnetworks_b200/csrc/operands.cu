// operands.cu -- the small glue around the quater_gen2 GEMMs (reference models/qec_module.py:124,
// nn.Linear(out_channels, 4*in*out) and its backward) when they run on the tensor cores at FP32
// accuracy ("3xTF32": x = hi + lo with hi, lo on the TF32 grid, products hi*hi + hi*lo + lo*hi).
//
//   qec_split_operand : [x | extra column] -> [blk0 | blk1 | blk2 | 0] with blk in {hi, lo} at a given
//                       block stride, rows optionally gathered through a permutation; and/or the hi
//                       part in BF16.  One launch replaces ~8 ATen launches (cat, round,
//                       subtract, round, pad, cat ...) per operand.
//   qec_fold_chunks   : sum of the split-K partial products of the two backward GEMMs, folded onto
//                       the K real columns and scattered straight into the parameter gradients.
// HBM-bound elementwise work on small tensors (<= 200 MB per step); no reference counterpart -- the
// reference calls cuBLAS SGEMM through nn.Linear.
#include "qec_common.cuh"

namespace qec {

__device__ __forceinline__ float tf32_rn(float v) {
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ float operand_at(const float* __restrict__ src, const float* __restrict__ extra,
                                            float extra_const, long long sr, int c, int C, int add_col) {
    if (c < C) return __ldg(src + sr * C + c);
    if (c == C && add_col) return extra ? __ldg(extra + sr) : extra_const;
    return 0.f;
}

// one warp per output row (lanes run along the columns: coalesced stores, the permutation entry and the
// source row are read once per warp): out_f32 [R, WT] = nblk blocks of stride BS (block j = hi or lo
// part of the Kc = C + add_col real columns), zero elsewhere; out_bf16 [R, WB] = hi part, zero-padded
__global__ void __launch_bounds__(256)
    split_operand_kernel(const float* __restrict__ src, const float* __restrict__ extra, float extra_const,
                         const long long* __restrict__ perm, float* __restrict__ out_f32,
                         unsigned short* __restrict__ out_bf16, long long R, int C, int add_col, int BS, int WT, int WB,
                         int b0, int b1, int b2, int nblk) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const int lane = threadIdx.x & 31;
    const int Kc = C + (add_col ? 1 : 0);
    const long long sr = perm ? perm[r] : r;
    if (out_f32) {
        float* o = out_f32 + r * WT;
        for (int c = lane; c < BS; c += 32) {  // the same source column feeds block j at column j * BS + c
            float hi = 0.f, lo = 0.f;
            if (c < Kc) {
                const float v = operand_at(src, extra, extra_const, sr, c, C, add_col);
                hi = tf32_rn(v);
                lo = tf32_rn(v - hi);
            }
            if (nblk > 0) o[c] = (b0 == 1) ? hi : lo;
            if (nblk > 1) o[BS + c] = (b1 == 1) ? hi : lo;
            if (nblk > 2) o[2 * BS + c] = (b2 == 1) ? hi : lo;
        }
        for (int c = nblk * BS + lane; c < WT; c += 32) o[c] = 0.f;  // tail padding
    }
    if (out_bf16) {
        unsigned short* o = out_bf16 + r * WB;
        for (int c = lane; c < WB; c += 32) {
            float hi = 0.f;
            if (c < Kc) hi = tf32_rn(operand_at(src, extra, extra_const, sr, c, C, add_col));
            unsigned short h;
            asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(hi));
            o[c] = h;
        }
    }
}

// out[row(r), c] = sum_s ( a[s, r, c] + a[s, r, KP + c] + b[s, r, c] ),  c < K, in ascending s
__global__ void __launch_bounds__(256)
    fold_chunks_kernel(const float* __restrict__ a, const float* __restrict__ b, const long long* __restrict__ perm,
                       float* __restrict__ out_main, float* __restrict__ out_last, int S, long long R, int KP, int K,
                       int ld_main) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * K) return;
    const long long r = idx / K;
    const int c = (int)(idx - r * K);
    // the chunks are summed in ascending order; four are in flight at a time (the loads, not the adds,
    // are what this loop waits for)
    float acc = 0.f;
    int s = 0;
    for (; s + 4 <= S; s += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* pa = a + ((long long)(s + u) * R + r) * (2 * KP);
            const float* pb = b + ((long long)(s + u) * R + r) * KP;
            v[u] = (__ldg(pa + c) + __ldg(pa + KP + c)) + __ldg(pb + c);
        }
        acc += v[0]; acc += v[1]; acc += v[2]; acc += v[3];
    }
    for (; s < S; ++s) {
        const float* pa = a + ((long long)s * R + r) * (2 * KP);
        const float* pb = b + ((long long)s * R + r) * KP;
        acc += (__ldg(pa + c) + __ldg(pa + KP + c)) + __ldg(pb + c);
    }
    const long long orow = perm ? perm[r] : r;
    if (c < ld_main) out_main[orow * ld_main + c] = acc;
    else if (out_last) out_last[orow] = acc;  // c == K - 1 == ld_main: the bias / ones column
}

}  // namespace qec

// src [R, C] (+ one extra column when add_col: extra[R] if given, else extra_const) = Kc real columns.
// out_f32 [R, WT] (nullable): nblk <= 3 blocks at stride BS >= Kc, block j = the TF32 hi part (blkj == 1)
// or the lo part (blkj == 2) of the operand, zeros elsewhere; out_bf16 [R, WB] (nullable): the hi part in
// BF16, zero-padded.  Output row r is read from source row perm[r] when perm is given.
extern "C" int qec_split_operand(const float* src, const float* extra, float extra_const, const long long* perm,
                                 float* out_f32, unsigned short* out_bf16, long long R, int C, int add_col, int BS,
                                 int WT, int WB, int blk0, int blk1, int blk2, int nblk, void* stream) {
    QEC_CHECK_ARG(src, 1);
    QEC_CHECK_ARG(out_f32 || out_bf16, 5);
    QEC_CHECK_ARG(R > 0, 7);
    const int Kc = C + (add_col ? 1 : 0);
    QEC_CHECK_ARG(C > 0, 8);
    QEC_CHECK_ARG(nblk >= 0 && nblk <= 3, 16);
    if (out_f32) QEC_CHECK_ARG(BS >= Kc && WT >= nblk * BS, 10);
    if (out_bf16) QEC_CHECK_ARG(WB >= Kc, 12);
    const unsigned grid = (unsigned)((R + 7) / 8);  // 8 warps = 8 rows per CTA
    qec::split_operand_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, extra, extra_const, perm, out_f32, out_bf16, R, C, add_col, BS, WT, WB, blk0, blk1, blk2, nblk);
    return qec::launch_status();
}

// a [S, R, 2*KP], b [S, R, KP] (split-K partial products) -> out_main [*, ld_main] (columns 0..ld_main-1)
// and, when ld_main == K - 1, column K - 1 -> out_last [*] (nullable: dropped).  Row r lands in row perm[r].
extern "C" int qec_fold_chunks(const float* a, const float* b, const long long* perm, float* out_main,
                               float* out_last, int S, long long R, int KP, int K, int ld_main, void* stream) {
    QEC_CHECK_ARG(a && b, 1);
    QEC_CHECK_ARG(out_main, 4);
    QEC_CHECK_ARG(S > 0 && R > 0, 6);
    QEC_CHECK_ARG(K > 0 && KP >= K && (ld_main == K || ld_main == K - 1), 9);
    const long long n = R * K;
    const unsigned grid = (unsigned)((n + 255) / 256);
    qec::fold_chunks_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, perm, out_main, out_last, S, R,
                                                                                 KP, K, ld_main);
    return qec::launch_status();
}
