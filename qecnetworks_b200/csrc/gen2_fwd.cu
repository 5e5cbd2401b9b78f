// gen2_fwd.cu -- the pool2 transform generator quater_gen2 (reference models/qec_module.py:124: nn.Linear(O, 4*I*O),
// bias included) on the 5th-generation tensor cores at FP32 accuracy, writing the transforms capsule-major:
//
//   t [M, N] = [h, 1] @ [W2[perm], b2[perm]]^T      M = B*P*K rows, N = 4*I*O columns (o, q, i), contraction Kc + 1 <= 48
//
// 671 MB of output at BASELINE config 2 against 1.4 M input floats: the kernel is bound by writing t.  Persistent CTAs
// (one per SM) walk over [128 x 128] output tiles, row block by row block:
//   warp 5  TMA producer: the row block's A operand (h, hi and lo parts, 64 KB) once per row block; the tile's B operand
//           (W2 rows, hi and lo, 64 KB) into a two-stage ring
//   warp 4  allocates all 512 columns of tensor memory (four accumulator buffers of 128 columns); one lane issues
//           tcgen05.mma kind::tf32 (M = N = 128, K = 8): A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T into one accumulator
//   warps 0-3  epilogue: tcgen05.ld (the next chunk in flight while this one is staged), 32-column chunks staged in
//              shared memory (128-byte swizzle, conflict-free) and written by TMA stores, one CTA barrier per chunk;
//              overlaps the following tiles' MMAs through the accumulator ring
// Measured at BASELINE config 2 (M = 8 192, N = 20 480, Kc = 40; 10 240 tiles over 148 CTAs): 158 us including the
// prep launch = 4.6 TB/s of output, error 6e-7 of the largest entry against float64.  The cuBLAS 3xTF32 GEMM it would
// replace takes 120 us + 14 us of operand launches, so the training path keeps cuBLAS by default and this kernel is
// opt-in (QEC_GEN2_FWD_TC=1, qec_module.GEN2_FORWARD_TCGEN05).  What bounds it is shared-memory bandwidth: per 64 KB
// output tile 64 KB of B operand are filled, 144 KB of operands are read by the 18 MMAs and 2 x 64 KB pass through the
// store staging = 336 KB at 128 B/clk.  History: rows written straight from registers (one 16-byte request per thread
// and store): 365 us; TMA stores with two barriers per chunk and a separate hi*lo accumulator: 170 us; this form: 158 us.
// Operands are K-major with the 128-byte swizzle, the contraction padded to 64 = two panels of 32 (zeros beyond
// Kc + 1 are loaded but only ceil((Kc + 1) / 8) k steps are issued).  A prep kernel builds the split operands (hi
// rounded to the TF32 grid, lo = x - hi; the bias column; the row permutation of W2).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdint>
#include <cstdlib>

#include "qec_common.cuh"
#include "tc_common.cuh"

namespace qec {
namespace g2f {

using namespace qec::g2;

constexpr int kThreads = 192;
constexpr int kKP = 64;                       // padded contraction length: two 128-byte-swizzle panels of 32
constexpr int kPanelBytes = 128 * 32 * 4;     // [128 rows x 32 k] floats
constexpr int kABytes = 4 * kPanelBytes;      // panel 0: hi, lo; panel 1: hi, lo
constexpr int kBStageBytes = 4 * kPanelBytes;
constexpr int kBStages = 2;
constexpr int kOutChunkBytes = 128 * 32 * 4;  // [128 rows x 32 columns] of the output tile, staged for a TMA store
constexpr int kOutStages = 2;
constexpr int kSmemBytes = kABytes + kBStages * kBStageBytes + kOutStages * kOutChunkBytes;
constexpr int kAccStages = 4;                 // accumulator buffers of 128 columns: all 512 columns of tensor memory
constexpr int kTmemCols = 128 * kAccStages;

struct Maps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo, out;
};

__global__ void __launch_bounds__(kThreads, 1)
    gen2_fwd_kernel(const __grid_constant__ Maps maps, float* __restrict__ t, int M, int N, int ksteps, int items_per_cta) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 + 2 * kBStages + 2 * kAccStages];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = N / 128;
    const int total = (M / 128) * NT;
    const int first = blockIdx.x * items_per_cta;
    const int last = min(total, first + items_per_cta);
    const uint32_t a_full = smem_addr(&bars[0]), a_empty = smem_addr(&bars[1]), b_full0 = smem_addr(&bars[2]),
                   b_empty0 = smem_addr(&bars[2 + kBStages]), acc_full0 = smem_addr(&bars[2 + 2 * kBStages]),
                   acc_empty0 = smem_addr(&bars[2 + 2 * kBStages + kAccStages]);
    const uint32_t a_base = smem_addr(smem), b_base = a_base + kABytes;
    unsigned char* out_stage = smem + kABytes + kBStages * kBStageBytes;

    if (tid == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(b_full0 + 8 * s, 1);
            mbar_init(b_empty0 + 8 * s, 1);
        }
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(acc_full0 + 8 * s, 1);
            mbar_init(acc_empty0 + 8 * s, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (warp == 5) {
        // ------------------------------------------------ TMA producer (one lane)
        if (lane == 0) {
            int cur_rb = -1, reloads = 0;
            for (int it = first, j = 0; it < last; ++it, ++j) {
                const int rb = it / NT, nt = it - rb * NT;
                if (rb != cur_rb) {  // new row block: its A operand replaces the old one once every MMA reading that is done
                    if (reloads > 0) mbar_wait(a_empty, (uint32_t)(reloads - 1) & 1u);
                    mbar_expect_tx(a_full, kABytes);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        tma_load_2d(a_base + (2 * p) * kPanelBytes, &maps.a_hi, 32 * p, rb * 128, a_full);
                        tma_load_2d(a_base + (2 * p + 1) * kPanelBytes, &maps.a_lo, 32 * p, rb * 128, a_full);
                    }
                    cur_rb = rb;
                    ++reloads;
                }
                const int s = j % kBStages;
                if (j >= kBStages) mbar_wait(b_empty0 + 8 * s, (uint32_t)((j / kBStages) - 1) & 1u);
                const uint32_t st = b_base + (uint32_t)s * kBStageBytes, mb = b_full0 + 8 * s;
                mbar_expect_tx(mb, kBStageBytes);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    tma_load_2d(st + (2 * p) * kPanelBytes, &maps.b_hi, 32 * p, nt * 128, mb);
                    tma_load_2d(st + (2 * p + 1) * kPanelBytes, &maps.b_lo, 32 * p, nt * 128, mb);
                }
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ------------------------------------------------ MMA issuer (one lane)
        const uint32_t idesc = make_idesc(0, 0, 128);
        if (lane == 0) {
            int cur_rb = -1, reloads = 0;
            for (int it = first, j = 0; it < last; ++it, ++j) {
                const int rb = it / NT;
                if (rb != cur_rb) {
                    mbar_wait(a_full, (uint32_t)reloads & 1u);
                    cur_rb = rb;
                    ++reloads;
                }
                const int s = j % kBStages, ab = j % kAccStages;
                mbar_wait(b_full0 + 8 * s, (uint32_t)(j / kBStages) & 1u);
                if (j >= kAccStages) mbar_wait(acc_empty0 + 8 * ab, (uint32_t)((j / kAccStages) - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + (uint32_t)ab * 128u;
                const uint32_t bst = b_base + (uint32_t)s * kBStageBytes;
                uint32_t accumulate = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {  // hi*hi, hi*lo, lo*hi into ONE accumulator: the epilogue is the
                    // critical path of this write-bound kernel, a single tcgen05.ld per element keeps it short
                    const int ap = (term == 2) ? 1 : 0, bp = (term == 1) ? 1 : 0;  // 0 = hi tile, 1 = lo tile of the panel
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const int p = ks >> 2, o = (ks & 3) * 32;
                        const uint64_t da = make_desc(a_base + (2 * p + ap) * kPanelBytes + o, 16, 1024, 2);
                        const uint64_t db = make_desc(bst + (2 * p + bp) * kPanelBytes + o, 16, 1024, 2);
                        mma_tf32(acc, da, db, idesc, accumulate);
                        accumulate = 1;
                    }
                }
                mma_commit(b_empty0 + 8 * s);
                mma_commit(acc_full0 + 8 * ab);
                const bool last_of_rb = (it + 1 >= last) || ((it + 1) / NT != rb);
                if (last_of_rb) mma_commit(a_empty);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ epilogue: (hi*hi + lo*hi) + hi*lo -> t
        // One row of the tile per thread (tcgen05.ld 32x32b).  Writing the rows straight from registers costs one
        // 16-byte request per thread and store (32 lines per warp instruction: measured 365 us for the 671 MB of
        // BASELINE config 2, 1.8 TB/s), so 32-column chunks are staged in shared memory in the 128-byte-swizzle
        // pattern (conflict-free 16-byte stores) and written by TMA (cp.async.bulk.tensor, global <- shared).
        const int row_in_tile = warp * 32 + lane;
        const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
        auto load32 = [&](uint32_t (&v)[32], uint32_t addr) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[16 * q + 0]), "=r"(v[16 * q + 1]), "=r"(v[16 * q + 2]), "=r"(v[16 * q + 3]), "=r"(v[16 * q + 4]),
                      "=r"(v[16 * q + 5]), "=r"(v[16 * q + 6]), "=r"(v[16 * q + 7]), "=r"(v[16 * q + 8]), "=r"(v[16 * q + 9]),
                      "=r"(v[16 * q + 10]), "=r"(v[16 * q + 11]), "=r"(v[16 * q + 12]), "=r"(v[16 * q + 13]),
                      "=r"(v[16 * q + 14]), "=r"(v[16 * q + 15])
                    : "r"(addr + 16u * q));
        };
        int chunk = 0;  // running count of staged chunks: staging buffer = chunk & 1
        for (int it = first, j = 0; it < last; ++it, ++j) {
            const int rb = it / NT, nt = it - rb * NT;
            const int ab = j % kAccStages;
            mbar_wait(acc_full0 + 8 * ab, (uint32_t)(j / kAccStages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = lane_base + (uint32_t)ab * 128u;
            uint32_t v[2][32];
            load32(v[0], ta);
#pragma unroll
            for (int c = 0; c < 4; ++c, ++chunk) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c < 3) load32(v[(c + 1) & 1], ta + 32u * (c + 1));  // the next chunk travels while this one is staged
                unsigned char* buf = out_stage + (chunk & 1) * kOutChunkBytes;
#pragma unroll
                for (int q = 0; q < 8; ++q)  // 16-byte piece q of this row goes to position q ^ (row % 8): 128-byte swizzle
                    *reinterpret_cast<float4*>(buf + row_in_tile * 128 + ((q ^ (row_in_tile & 7)) << 4)) =
                        make_float4(__uint_as_float(v[c & 1][4 * q]), __uint_as_float(v[c & 1][4 * q + 1]),
                                    __uint_as_float(v[c & 1][4 * q + 2]), __uint_as_float(v[c & 1][4 * q + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to TMA
                // One barrier per chunk: before it, thread 0 makes sure every earlier store has finished READING its
                // staging buffer, so after it the other buffer (read by the previous chunk's store) is free again.
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&maps.out),
                                 "r"(nt * 128 + 32 * c), "r"(rb * 128), "r"(smem_addr(buf))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(acc_empty0 + 8 * ab);  // this accumulator buffer may be overwritten
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store has landed
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
    }
}

// split operands, K-major, padded to kKP columns:
//   A_hi/lo [M][64]: [h[m][0..Kc), 1, 0...]           B_hi/lo [N][64]: [W2[perm[n]][0..Kc), b2[perm[n]], 0...]
__global__ void __launch_bounds__(256)
    prep_fwd_kernel(const float* __restrict__ h, const float* __restrict__ W2, const float* __restrict__ b2,
                    const long long* __restrict__ perm, float* __restrict__ a_hi, float* __restrict__ a_lo,
                    float* __restrict__ b_hi, float* __restrict__ b_lo, int M, int N, int Kc) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long na = (long long)M * kKP;
    float x;
    float *ph, *pl;
    if (e < na) {
        const int m = (int)(e / kKP), j = (int)(e % kKP);
        x = (j < Kc) ? __ldg(h + (size_t)m * Kc + j) : ((j == Kc) ? 1.f : 0.f);
        ph = a_hi + e;
        pl = a_lo + e;
    } else if (e < na + (long long)N * kKP) {
        const long long f = e - na;
        const int n = (int)(f / kKP), j = (int)(f % kKP);
        const long long row = (perm != nullptr) ? __ldg(perm + n) : n;
        x = (j < Kc) ? __ldg(W2 + (size_t)row * Kc + j) : ((j == Kc) ? __ldg(b2 + row) : 0.f);
        ph = b_hi + f;
        pl = b_lo + f;
    } else {
        return;
    }
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // round to nearest on the TF32 grid
    *ph = hi;
    *pl = x - hi;
}

}  // namespace g2f
}  // namespace qec

using namespace qec;

// 1 if qec_gen2_fwd serves the shape: M, N multiples of 128, Kc + 1 <= 64
extern "C" int qec_gen2_fwd_supported(int M, int N, int Kc) {
    return (M > 0 && N > 0 && M % 128 == 0 && N % 128 == 0 && Kc >= 1 && Kc + 1 <= g2f::kKP) ? 1 : 0;
}

// floats of workspace qec_gen2_fwd needs (the split operands)
extern "C" long long qec_gen2_fwd_workspace(int M, int N, int Kc) {
    if (!qec_gen2_fwd_supported(M, N, Kc)) return 0;
    return 2ll * g2f::kKP * ((long long)M + N);
}

// h [M, Kc], W2 [N, Kc] and b2 [N] in the reference's row order, perm [N] int64 | NULL (row of W2 / b2 behind output
// column n) -> t [M, N] = h W2[perm]^T + b2[perm].  workspace: qec_gen2_fwd_workspace floats.
extern "C" int qec_gen2_fwd(const float* h, const float* W2, const float* b2, const long long* perm, float* t,
                            float* workspace, int M, int N, int Kc, void* stream) {
    QEC_CHECK_ARG(h, 1);
    QEC_CHECK_ARG(W2, 2);
    QEC_CHECK_ARG(b2, 3);
    QEC_CHECK_ARG(t, 5);
    QEC_CHECK_ARG(workspace, 6);
    QEC_CHECK_ARG(qec_gen2_fwd_supported(M, N, Kc), 7);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* a_hi = workspace;
    float* a_lo = a_hi + (size_t)M * g2f::kKP;
    float* b_hi = a_lo + (size_t)M * g2f::kKP;
    float* b_lo = b_hi + (size_t)N * g2f::kKP;
    {
        const long long tot = (long long)g2f::kKP * ((long long)M + N);
        g2f::prep_fwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h, W2, b2, perm, a_hi, a_lo, b_hi, b_lo, M, N, Kc);
    }
    static PerDeviceInt sms;  // SM count of the device (persistent grid) + opt-in shared memory, once per device
    int nsm = 0;
    if (std::atomic<int>* slot = sms.slot(); slot && slot->load(std::memory_order_relaxed) > 0) {
        nsm = slot->load(std::memory_order_relaxed);
    } else {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(g2f::gen2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g2f::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        if (slot) slot->store(nsm, std::memory_order_relaxed);
    }
    static PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }();
    if (encode == nullptr) return (int)cudaErrorNotSupported;
    auto make_map = [&](CUtensorMap* map, const float* base, unsigned long long rows) {
        const cuuint64_t dims[2] = {(cuuint64_t)g2f::kKP, rows};
        const cuuint64_t strides[1] = {g2f::kKP * sizeof(float)};
        const cuuint32_t box[2] = {32, 128};
        const cuuint32_t estr[2] = {1, 1};
        return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    g2f::Maps mp;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(float)};
        const cuuint32_t box[2] = {32, 128};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&mp.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return (int)cudaErrorInvalidValue;
    }
    if (!make_map(&mp.a_hi, a_hi, M) || !make_map(&mp.a_lo, a_lo, M) || !make_map(&mp.b_hi, b_hi, N) ||
        !make_map(&mp.b_lo, b_lo, N))
        return (int)cudaErrorInvalidValue;
    const int total = (M / 128) * (N / 128);
    int ctas = nsm < total ? nsm : total;
    const int per = (total + ctas - 1) / ctas;
    ctas = (total + per - 1) / per;
    const int ksteps = (Kc + 1 + 7) / 8;
    g2f::gen2_fwd_kernel<<<ctas, g2f::kThreads, g2f::kSmemBytes, st>>>(mp, t, M, N, ksteps, per);
    return launch_status();
}
