// gen2_bwd.cu -- backward of the pool2 transform generator quater_gen2 (reference models/qec_module.py:124,
// nn.Linear(O, 4*I*O) autograd) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulator in
// tensor memory, at FP32 accuracy (3-term split), reading the FP32 transform gradient the routing backward wrote.
//
//   g  [M, N]  = dL/dt_oqi, M = B*P*K rows (b,p,k), N = 4*I*O columns in capsule-major order (o, q, i)
//   dh [M, Kc] = g  @ Wp          Wp[n, :] = W2[perm[n], :]                     (contraction over N = 20 480)
//   dW [N, Kc+1] = g^T @ [h, 1]   -> dL/dW2[perm[n], :], dL/db2[perm[n]]        (contraction over M =  8 192)
// Both outputs are tiny (Kc = 40 classes); the work is streaming g, 671 MB at BASELINE config 2.  The cuBLAS route it
// replaces ran four batched GEMMs over a pre-split copy of g (TF32 plane + BF16 remainder: 1.0 GB, read twice) plus
// split / fold launches: 0.34 + 0.07 ms per step; here g is plain FP32, read once per product.
//
// One kernel, two modes (one launch each), both with the big operand A = a tile of g and the small one as B (N = 48):
//   mode 0 (dh): CTA = 128 rows of g x a slice of the columns; chunk = [128 x 32]; A is K-major (k = column).
//   mode 1 (dW): CTA = 128 columns of g x a slice of the rows;  chunk = [32 x 128]; A is MN-major (k = row): the
//                tensor core reads g^T straight from the same kind of tile -- no transposed copy of g anywhere.
// Shared-memory tiles are written by TMA in the tensor core's canonical swizzled layouts: 128-byte swizzle for the
// K-major operands (rows of 128 B = 32 k), and -- for g^T -- the 128-byte swizzle with a 32-byte base, which is the
// one MN-major layout tcgen05 accepts for 32-bit operands (the swizzle-free "interleaved" MN-major form computes
// zeros for tf32: measured).
//
// FP32 accuracy: tf32 keeps 10 mantissa bits, so every operand is split x = hi + lo with hi on the TF32 grid and
//   A B = A_hi [B_hi | B_lo] + A_lo B_hi         (the dropped lo*lo term is 2^-22 relative)
// accumulated in FP32 in tensor memory (96 columns: the hi*lo products keep an accumulator of their own and are added
// in the epilogue).  The small operands are split once by a prep kernel (hi rounded to the TF32
// grid); for the tile of g, hi is what the tensor core itself reads from an FP32 word -- the upper 19 bits, i.e.
// truncation, verified bit-for-bit against an explicitly masked copy -- and lo = x - trunc(x) is written to a twin
// tile by four splitter warps.
//
// Warp roles (192 threads): warp 5 = TMA producer (cp.async.bulk.tensor into a ring of kStages stages, completion
// on an mbarrier), warps 0-3 = splitters and finally the epilogue (tcgen05.ld of the accumulator to a partial-sum
// buffer), warp 4 allocates tensor memory and one of its lanes issues the MMAs; tcgen05.commit returns a stage to the
// producer when the MMAs that read it have finished.  (A first version fed the ring with 16-byte cp.async from 128
// threads: 731 us for both products; TMA: 405 us -- the SM could not keep enough 16-byte requests in flight.)  A small fold kernel adds the partial sums of the column / row slices in a fixed order and
// scatters them to dL/dh, dL/dW2, dL/db2: deterministic, no atomics.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdint>
#include <cstdlib>

#include "qec_common.cuh"
#include "tc_common.cuh"

namespace qec {
namespace g2 {

constexpr int kCvtThreads = 128;  // warps 0-3: split the tile of g, finally drain the accumulator
constexpr int kThreads = 192;     // + warp 4: MMA issuer / tensor-memory allocator, warp 5: TMA producer
constexpr int kNB = 48;           // padded width of the small operand (Kc + 1 <= 48): the MMA's N
#ifndef QEC_G2_STAGES
#define QEC_G2_STAGES 5
#endif
constexpr int kStages = QEC_G2_STAGES;  // ring depth: 5 x 44 KB fill the SM; fewer leave room for a co-resident kernel
constexpr int kTileBytes = 16 * 1024;           // a chunk of g: [128 x 32] (mode 0) or [32 x 128] (mode 1) floats
constexpr int kSmallBytes = kNB * 32 * 4;       // the matching chunk of the small operand: 48 rows x 32 k, K-major
constexpr int kStageBytes = 2 * kTileBytes + 2 * kSmallBytes;  // g hi, g lo, small hi, small lo
constexpr int kTxBytes = kTileBytes + 2 * kSmallBytes;         // what TMA delivers per stage
constexpr int kTmemCols = 128;                  // power of two >= 2 kNB: columns [0,48) hi*hi + lo*hi, [48,96) hi*lo

struct Maps {  // TMA descriptors of one launch
    CUtensorMap g, small_hi, small_lo;
};

// MODE 0 (dh): this CTA owns rows [mn0, mn0 + 128) of g and the 32-wide column chunks slice, slice + nslices, ...
//   A = g tile [128 rows x 32 k], K-major, 128-byte swizzle (TMA box {32, 128});  B = WT tile [48 x 32 k], same format.
// MODE 1 (dW): this CTA owns columns [mn0, mn0 + 128) of g and the 32-tall row chunks slice, slice + nslices, ...
//   A = (g tile [32 k x 128 columns])^T, MN-major: four TMA boxes {32 columns, 32 rows} in the 128-byte swizzle with
//   32-byte base, the one MN-major layout the tensor core takes for 32-bit operands;  B = hT tile [48 x 32 k], K-major.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
    gen2_bwd_kernel(const __grid_constant__ Maps maps, float* __restrict__ partial, int M, int N, int klen, int nslices,
                    uint32_t flags) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[3 * kStages + 1];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1-D grid, slice index fastest: the CTAs that share a row / column block are neighbours in launch order and take
    // the k chunks round-robin, so that at any time they read ADJACENT pieces of the same rows of g
    const int slice = blockIdx.x % nslices;
    const int mn0 = (blockIdx.x / nslices) * 128;
    const int nchunks = klen / 32;
    const uint32_t tma_full0 = smem_addr(&bars[0]), cvt_full0 = smem_addr(&bars[kStages]),
                   empty0 = smem_addr(&bars[2 * kStages]), done = smem_addr(&bars[3 * kStages]);
    const uint32_t stage0 = smem_addr(smem);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(tma_full0 + 8 * s, 1);
            mbar_init(cvt_full0 + 8 * s, kCvtThreads);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(done, 1);
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
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages;
                if (c >= kStages) mbar_wait(empty0 + 8 * s, (uint32_t)((c / kStages) - 1) & 1u);
                const uint32_t st = stage0 + (uint32_t)s * kStageBytes, mb = tma_full0 + 8 * s;
                const int kc = (c * nslices + slice) * 32;
                mbar_expect_tx(mb, kTxBytes);
                if (MODE == 0) {
                    tma_load_2d(st, &maps.g, kc, mn0, mb);              // box {32 columns, 128 rows}
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) tma_load_2d(st + a * 4096, &maps.g, mn0 + 32 * a, kc, mb);  // {32, 32}
                }
                tma_load_2d(st + 2 * kTileBytes, &maps.small_hi, kc, 0, mb);              // box {32 k, 48 rows}
                tma_load_2d(st + 2 * kTileBytes + kSmallBytes, &maps.small_lo, kc, 0, mb);
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ------------------------------------------------ MMA issuer (one lane)
        // The hi and lo tiles of the small operand lie back to back (48 rows = six 1024-byte swizzle groups each), so
        // one MMA with N = 96 multiplies the tile of g by both: A_hi [B_hi | B_lo] -> columns [0, 96), then
        // A_lo B_hi -> columns [0, 48).  Two reads of a g tile per chunk instead of three: the kernel is bound by
        // shared-memory bandwidth (TMA fill + split + operand reads), not by the tensor pipe.
        const uint32_t idesc2 = make_idesc(MODE, 0, 2 * kNB), idesc1 = make_idesc(MODE, 0, kNB);
        if (lane == 0) {
            uint32_t acc = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages;
                mbar_wait(cvt_full0 + 8 * s, (uint32_t)(c / kStages) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = stage0 + (uint32_t)s * kStageBytes, a_lo = a_hi + kTileBytes;
                const uint32_t b_hi = a_hi + 2 * kTileBytes;
                if (!(flags & 1u))
#pragma unroll
                for (int term = 0; term < 2; ++term) {  // hi * [hi | lo], lo * hi
                    const uint32_t ab = (term == 1) ? a_lo : a_hi, bb = b_hi;
                    const uint32_t idesc = (term == 1) ? idesc1 : idesc2;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {    // 32 k per chunk, 8 per MMA
                        // K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart; 8 k = 32 B along the row
                        // MN-major (mode 1's A): k-quads 512 B apart (SBO), 32-column atoms 4096 B apart (LBO); 8 k = 1024 B
                        const uint64_t da = (MODE == 0) ? make_desc(ab + ks * 32, 16, 1024, 2)
                                                        : make_desc(ab + ks * 1024, 4096, 512, 1);
                        const uint64_t db = make_desc(bb + ks * 32, 16, 1024, 2);
                        mma_tf32(tmem_d, da, db, idesc, acc);
                        acc = 1;
                    }
                }
                mma_commit(empty0 + 8 * s);  // arrives when the MMAs above have finished reading stage s
            }
            mma_commit(done);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ splitters: x -> hi (in place, on the TF32 grid) + lo (twin tile)
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % kStages;
            mbar_wait(tma_full0 + 8 * s, (uint32_t)(c / kStages) & 1u);
            float4* t = reinterpret_cast<float4*>(smem + (size_t)s * kStageBytes);
            if (!(flags & 2u))
#pragma unroll
            for (int i = 0; i < kTileBytes / 16 / kCvtThreads; ++i) {  // element-wise: the tile's layout does not matter
                float4* p = t + tid + i * kCvtThreads;
                const float4 x = *p;
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); l.x = x.x - h.x;
                h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u); l.y = x.y - h.y;
                h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); l.z = x.z - h.z;
                h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u); l.w = x.w - h.w;
                // The raw FP32 tile itself serves as the hi operand: kind::tf32 reads the upper 19 bits of each word, i.e.
                // exactly the mask above (measured: results bit-identical with and without storing h; flag 4 stores it).
                if (flags & 4u) *p = h;
                p[kTileBytes / 16] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
            mbar_arrive(cvt_full0 + 8 * s);
        }
        // ------------------------------------------------ epilogue: accumulator [128 lanes x 48 columns] -> partial
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = warp * 32 + lane;
        float* out = partial + ((size_t)slice * ((MODE == 0) ? M : N) + mn0 + row) * kNB;
#pragma unroll
        for (int cb = 0; cb < kNB; cb += 16) {
            uint32_t v[16], w[16];
            const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(ta));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                  "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "r"(ta + (uint32_t)kNB));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; i += 4)  // (hi*hi + lo*hi) + hi*lo
                *reinterpret_cast<float4*>(out + cb + i) = make_float4(
                    __uint_as_float(v[i]) + __uint_as_float(w[i]), __uint_as_float(v[i + 1]) + __uint_as_float(w[i + 1]),
                    __uint_as_float(v[i + 2]) + __uint_as_float(w[i + 2]), __uint_as_float(v[i + 3]) + __uint_as_float(w[i + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols) : "memory");
    }
}

// the small operands, split on the TF32 grid:
//   WT_hi/lo [48][N]: WT[j][n] = W2[perm[n]][j] (j < Kc), b2[perm[n]] (j == Kc: unused by dh, kept 0), 0 beyond
//   hT_hi/lo [48][M]: hT[j][m] = h[m][j] (j < Kc), 1 (j == Kc: the bias column), 0 beyond
__global__ void __launch_bounds__(256)
    prep_small_kernel(const float* __restrict__ W2, const long long* __restrict__ perm, const float* __restrict__ h,
                      float* __restrict__ wt_hi, float* __restrict__ wt_lo, float* __restrict__ h_hi,
                      float* __restrict__ h_lo, int M, int N, int Kc) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long nw = (long long)kNB * N;
    float x;
    float *dh_, *dl_;
    if (e < nw) {
        const int j = (int)(e / N);
        const int n = (int)(e - (long long)j * N);
        x = (j < Kc) ? __ldg(W2 + (size_t)__ldg(perm + n) * Kc + j) : 0.f;
        dh_ = wt_hi + e;
        dl_ = wt_lo + e;
    } else if (e < nw + (long long)M * kNB) {
        const long long f = e - nw;
        const int j = (int)(f / M), m = (int)(f - (long long)j * M);
        x = (j < Kc) ? __ldg(h + (size_t)m * Kc + j) : ((j == Kc) ? 1.f : 0.f);
        dh_ = h_hi + f;
        dl_ = h_lo + f;
    } else {
        return;
    }
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // round to nearest on the TF32 grid
    *dh_ = hi;
    *dl_ = x - hi;
}

// sum the partial results of the slices in a fixed order and scatter:
//   dh: g_h[m, j] = sum_s pdh[s][m][j]                          (j < Kc)
//   dW: g_W2[perm[n], j] = sum_s pdw[s][n][j], g_b2[perm[n]] = sum_s pdw[s][n][Kc]
__global__ void __launch_bounds__(256)
    fold_kernel(const float* __restrict__ pdh, const float* __restrict__ pdw, const long long* __restrict__ perm,
                float* __restrict__ g_h, float* __restrict__ g_W2, float* __restrict__ g_b2, int M, int N, int Kc, int SN,
                int SM) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long nh = (long long)M * kNB;
    if (e < nh) {
        const int m = (int)(e / kNB), j = (int)(e - (long long)m * kNB);
        if (j >= Kc || g_h == nullptr) return;
        float s = 0.f;
        for (int k = 0; k < SN; ++k) s += pdh[(size_t)k * nh + e];
        g_h[(size_t)m * Kc + j] = s;
    } else if (e < nh + (long long)N * kNB) {
        const long long f = e - nh;
        const int n = (int)(f / kNB), j = (int)(f - (long long)n * kNB);
        if (j > Kc || g_W2 == nullptr) return;
        float s = 0.f;
        for (int k = 0; k < SM; ++k) s += pdw[(size_t)k * N * kNB + f];
        const long long row = __ldg(perm + n);
        if (j < Kc) g_W2[(size_t)row * Kc + j] = s;
        else g_b2[row] = s;
    }
}

}  // namespace g2
}  // namespace qec

using namespace qec;

static void g2_slices(int M, int N, int& SN, int& SM) {
    // mode 0 launches (M / 128) x SN CTAs, mode 1 (N / 128) x SM.  Slices also bound the length of one tensor-memory
    // accumulation (the FP32 accumulator loses ~4e-5 relative over 10^4 terms; the fold kernel adds the slices in FP32)
    // and even out the waves over the 148 SMs: aim at >= 4 waves.  Overridable for experiments (QEC_G2_SN / QEC_G2_SM).
    // (measured at M = 8192, N = 20480 with three MMAs per k step: 16 x 8 slices: 405 us, errors 1e-5 / 8e-6; 4 x 2: 433 us,
    // 4e-5 / 3e-5.  With the hi*hi and hi*lo products merged into one N = 96 MMA: 341 us, 7e-6 / 5e-6; the same launches
    // with the MMAs and the split compiled out, i.e. TMA streaming alone: 305 us)
    SN = 1;
    while (((M / 128) * SN < 4 * 148 || N / SN > 1280) && (N / (2 * SN)) % 32 == 0 && N / (2 * SN) >= 256 && SN < 16) SN *= 2;
    SM = 1;
    while (((N / 128) * SM < 4 * 148 || M / SM > 1024) && (M / (2 * SM)) % 32 == 0 && M / (2 * SM) >= 256 && SM < 16) SM *= 2;
    if (const char* e = getenv("QEC_G2_SN")) { const int v = atoi(e); if (v >= 1 && N % (32 * v) == 0) SN = v; }
    if (const char* e = getenv("QEC_G2_SM")) { const int v = atoi(e); if (v >= 1 && M % (32 * v) == 0) SM = v; }
}

// 1 if qec_gen2_bwd serves the shape: M, N multiples of 128 (and of the slice granularity), Kc + 1 <= 48
extern "C" int qec_gen2_bwd_supported(int M, int N, int Kc) {
    return (M > 0 && N > 0 && M % 128 == 0 && N % 128 == 0 && Kc >= 1 && Kc + 1 <= g2::kNB) ? 1 : 0;
}

// floats of workspace qec_gen2_bwd needs
extern "C" long long qec_gen2_bwd_workspace(int M, int N, int Kc) {
    if (!qec_gen2_bwd_supported(M, N, Kc)) return 0;
    int SN, SM;
    g2_slices(M, N, SN, SM);
    return 2ll * g2::kNB * N + 2ll * M * g2::kNB + (long long)SN * M * g2::kNB + (long long)SM * N * g2::kNB;
}

// g [M, N] FP32, h [M, Kc], W2 [N, Kc] (rows in the reference's order), perm [N] int64 (row of W2 behind column n of g)
// -> g_h [M, Kc] | NULL, g_W2 [N, Kc] + g_b2 [N] | NULL (both or neither).  workspace: qec_gen2_bwd_workspace floats.
// `variant` selects the descriptor stride convention (0 = documented; 1..3 = swapped LBO/SBO of A / B: bring-up aid).
extern "C" int qec_gen2_bwd(const float* g, const float* h, const float* W2, const long long* perm, float* g_h,
                            float* g_W2, float* g_b2, float* workspace, int M, int N, int Kc, int variant, void* stream) {
    QEC_CHECK_ARG(g, 1);
    QEC_CHECK_ARG(h, 2);
    QEC_CHECK_ARG(W2, 3);
    QEC_CHECK_ARG(perm, 4);
    QEC_CHECK_ARG((g_W2 == nullptr) == (g_b2 == nullptr), 6);
    QEC_CHECK_ARG(workspace, 8);
    QEC_CHECK_ARG(qec_gen2_bwd_supported(M, N, Kc), 9);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int SN, SM;
    g2_slices(M, N, SN, SM);
    float* wt_hi = workspace;
    float* wt_lo = wt_hi + (size_t)g2::kNB * N;
    float* h_hi = wt_lo + (size_t)g2::kNB * N;
    float* h_lo = h_hi + (size_t)M * g2::kNB;
    float* pdh = h_lo + (size_t)M * g2::kNB;
    float* pdw = pdh + (size_t)SN * M * g2::kNB;
    {
        const long long tot = (long long)g2::kNB * N + (long long)M * g2::kNB;
        g2::prep_small_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(W2, perm, h, wt_hi, wt_lo, h_hi, h_lo, M, N, Kc);
    }
    const size_t smem = (size_t)g2::kStages * g2::kStageBytes;
    static PerDeviceInt configured;
    if (std::atomic<int>* slot = configured.slot(); !slot || slot->load(std::memory_order_relaxed) < 0) {
        cudaError_t e = cudaFuncSetAttribute(g2::gen2_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(g2::gen2_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (slot) slot->store(1, std::memory_order_relaxed);
    }
    // TMA descriptors (host-side encode, passed by value as kernel parameters): FP32, 2-D, inner dimension first
    static PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }();
    if (encode == nullptr) return (int)cudaErrorNotSupported;
    auto make_map = [&](CUtensorMap* map, const float* base, unsigned long long inner, unsigned long long outer,
                        unsigned box_inner, unsigned box_outer, CUtensorMapSwizzle sw) {
        const cuuint64_t dims[2] = {inner, outer};
        const cuuint64_t strides[1] = {inner * sizeof(float)};
        const cuuint32_t box[2] = {box_inner, box_outer};
        const cuuint32_t estr[2] = {1, 1};
        return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    const uint32_t flags = (uint32_t)(variant >> 4);
    if (g_h != nullptr) {
        g2::Maps mp;
        if (!make_map(&mp.g, g, N, M, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&mp.small_hi, wt_hi, N, g2::kNB, 32, g2::kNB, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&mp.small_lo, wt_lo, N, g2::kNB, 32, g2::kNB, CU_TENSOR_MAP_SWIZZLE_128B))
            return (int)cudaErrorInvalidValue;
        g2::gen2_bwd_kernel<0><<<(M / 128) * SN, g2::kThreads, smem, st>>>(mp, pdh, M, N, N / SN, SN, flags);
    }
    if (g_W2 != nullptr) {
        g2::Maps mp;
        if (!make_map(&mp.g, g, N, M, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
            !make_map(&mp.small_hi, h_hi, M, g2::kNB, 32, g2::kNB, CU_TENSOR_MAP_SWIZZLE_128B) ||
            !make_map(&mp.small_lo, h_lo, M, g2::kNB, 32, g2::kNB, CU_TENSOR_MAP_SWIZZLE_128B))
            return (int)cudaErrorInvalidValue;
        g2::gen2_bwd_kernel<1><<<(N / 128) * SM, g2::kThreads, smem, st>>>(mp, pdw, M, N, M / SM, SM, flags);
    }
    {
        const long long tot = (long long)M * g2::kNB + (long long)N * g2::kNB;
        g2::fold_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pdh, pdw, perm, g_h, g_W2, g_b2, M, N, Kc, SN, SM);
    }
    return launch_status();
}
