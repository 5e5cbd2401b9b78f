// tc_common.cuh -- sm_100a building blocks shared by the tensor-core kernels (gen2_bwd.cu, gen2_fwd.cu): mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 shared-memory / instruction descriptors, tcgen05.mma kind::tf32, tcgen05.commit.
#pragma once
#include <cuda.h>

#include <cstdint>

namespace qec {
namespace g2 {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mb, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mb) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mb, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "G2WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra G2WAIT_%=;\n\t}" ::"r"(mb), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mb) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(mb)
                 : "memory");
}

// shared-memory matrix descriptor (sm_100 version); offsets in bytes.  layout_type: 2 = 128-byte swizzle (K-major
// tiles, rows of 128 B), 1 = 128-byte swizzle with a 32-byte base (the MN-major layout of 32-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, M = 128, N = n, majors as given (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 2u << 7;                    // a_format = TF32
    d |= 2u << 10;                   // b_format = TF32
    d |= (uint32_t)(a_mn & 1) << 15;
    d |= (uint32_t)(b_mn & 1) << 16;
    d |= (uint32_t)(n >> 3) << 17;    // N
    d |= (uint32_t)(128 >> 4) << 24;  // M
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mb) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
}

}  // namespace g2
}  // namespace qec
