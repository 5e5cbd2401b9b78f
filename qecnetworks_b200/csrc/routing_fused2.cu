// routing_fused2.cu -- the cluster-resident routing kernels of the pool2 regime, second
// generation: PACKED vote pairs (FFMA2) and asynchronous prefetch of the next capsule.
//
// Same contract as routing_fused.cu (reference models/qec_module.py:126-147, 172-194): vote
// generation + all T routing iterations + the activation epilogue of one output capsule with
// its votes resident in the shared memory of a thread-block cluster; the [B,P,O,K*I,4] vote
// tensor never exists in HBM.  What changed, and why (profiles/r01_fused_v5*: the first
// generation is bound by instruction issue -- issue slots 46 % busy, fma pipe 27 %, DRAM 12 %,
// top stall = barrier):
//
//  * Votes are processed as PAIRS (j, j+1) = inputs (k, i), (k, i+1) held in the low / high half
//    of 64-bit register pairs and every FP32 operation of a sweep is one FFMA2 / FMUL2 /
//    FADD2 for both (f32x2.cuh).  A routing sweep costs 23.5 issue slots per vote instead of
//    ~40 (FFMA2 = one issue slot, two fma-pipe cycles: the bound moves from issue to the fma
//    pipe).  Shared memory keeps the pair component-interleaved, {w0 w1 x0 x1} {y0 y1 z0 z1},
//    so a pair is two conflict-free LDS.128 straight into aligned register pairs; the
//    transforms arrive the same way with LDG.64 / 8-byte cp.async because i is the
//    contiguous axis of the capsule-major transform layout t_oqi.
//  * The transforms of the NEXT capsule are copied global -> shared with cp.async directly
//    into the vote slots while the epilogue sweep of the current capsule frees them, one
//    commit group per slot, so HBM latency hides behind the epilogue sweep and the first
//    reduction instead of stalling sweep 0.
//  * The epilogue's two sums ride in the next capsule's first reduction (15 values wide), so a
//    capsule costs T cluster reductions instead of T + 1.
//  * The cluster reduction needs no cluster barrier (PushReducer below: st.async pushes into the
//    peers' receive buffers, completion counted by their mbarriers).
//  * Backward: packed twin of routing_fused_bwd_kernel; the top level is two passes so that neither
//    spills; dL/dt can leave pre-split for the tensor-core GEMMs that consume it (TF32 plane +
//    BF16 remainder, qec_routing_fused_bwd_split).
//
// Requires I even (pairs never straddle a neighbour row and stay 8-byte aligned); other
// geometries use the scalar kernels of routing_fused.cu.
#include <cooperative_groups.h>

#include <cstdlib>

#include "routing_pair.cuh"

namespace cg = cooperative_groups;

// The file is compiled twice: as itself (256 threads per CTA, two CTAs per SM, clusters of up to 8) and through
// routing_fused2_wide.cu (QEC_PT = 512: ONE CTA of 16 warps per SM holding 8 192 votes, clusters of up to 4).
// Each build lives in its own nested namespace; the C entry points belong to the 256-thread build and hand
// the geometries that would need an 8-CTA cluster to the wide one (see qec_fused2_wide_* at the end).
#ifndef QEC_PT
#define QEC_PT 256
#endif
#ifndef QEC_VARIANT_NS
#define QEC_VARIANT_NS pt256
#endif
// Software-pipelined shared-memory operands (1 = on, the default; 0 = the previous schedule, kept for A/B timing).
// Every sweep both reads its slot (votes, activation, weight) and writes shared memory for it (the updated weight,
// the regenerated vote, or -- through cp.async, a compiler barrier -- the next capsule's transforms).  The compiler
// cannot prove that those stores do not alias the NEXT slot's loads, so it issued slot m+1's LDS only after slot m's
// whole dependent chain (dot -> distance polynomial -> weight -> STS) had retired: one slot in flight per warp, ~300
// cycles per pair and sweep against ~65 cycles of fma-pipe work (ncu source page, profiles/r01_fused2_ncu.txt: the
// first instruction after each LDS.128 carries the short-scoreboard stalls, every Horner step a fixed-latency wait).
// QEC_SWP is a bit mask of the sweeps that use it: 1 forward routing sweeps, 2 forward sweep 0, 4 forward epilogue,
// 8 backward accumulation sweeps, 16 backward sweep 0.
// With QEC_SWP the operands of slot m+1 are loaded into registers BEFORE slot m is computed and stored, so two
// slots' chains overlap inside each warp.  Same arithmetic in the same order: results are bit-identical.
#ifndef QEC_SWP
#define QEC_SWP 0
#endif
// Order in which the persistent clusters walk over the capsules.  0: capsule c = bp * O + o in memory order, so
// the ~33 clusters in flight work on the O capsules of ONE patch at a time (its LRFs / activations are shared:
// L1/L2-friendly) -- but in the backward they then all add into the SAME dL/dlrfs rows at about the same time, and
// the L2 atomic unit serialises per address.  1: walk o-major (u -> bp = u % BP, o = u / BP), so concurrent
// clusters sit on different patches: no same-address contention, LRFs / activations come from L2 (20 MB working
// set for 32 patches).
#ifndef QEC_SPREAD
#define QEC_SPREAD 0
#endif
#ifndef QEC_PP_DEFAULT
#define QEC_PP_DEFAULT 0
#endif
#ifndef QEC_RED_SHFL
#define QEC_RED_SHFL 0
#endif
__device__ __forceinline__ long long caps_of(long long u, long long BPn, int O) {
#if QEC_SPREAD
    const long long o = u / BPn;
    return (u - o * BPn) * O + o;
#else
    (void)BPn; (void)O;
    return u;
#endif
}

namespace qec {
namespace QEC_VARIANT_NS {

// Optional timeline instrumentation (build with QEC_NVCC_EXTRA=-DQEC_PROF; scripts/timeline_fused.py):
// thread 0 of a few CTAs stamps clock64() at the phase boundaries of its first capsules.
#ifdef QEC_PROF
__device__ long long g_prof[8 * 64 * 32];
#define QEC_EVT(id)                                                                                   \
    do {                                                                                              \
        if (threadIdx.x == 0 && blockIdx.x < 8 && prof_n < 64) g_prof[(blockIdx.x * 64 + prof_n) * 32 + (id)] = clock64(); \
    } while (0)
#define QEC_EVT_NEXT() (++prof_n)
#else
#define QEC_EVT(id) \
    do {            \
    } while (0)
#define QEC_EVT_NEXT() \
    do {               \
    } while (0)
#endif

constexpr int kPT = QEC_PT;      // threads per CTA
constexpr int kCtasPerSM = (kPT <= 256) ? 2 : 1;
constexpr int kPairs = 8;        // resident vote pairs per thread
constexpr int kNpMax = kPT * kPairs;  // 2048 pairs = 4096 votes per CTA



// ------------------------------------------------------------------------------------
// Cluster-wide sum of <= 16 floats per thread without a cluster barrier.
//   1. every thread parks its N partials in shared memory, one CTA barrier;
//   2. 16 lanes per value add 16 entries each and finish with a 4-step butterfly;
//   3. lane r of each group PUSHES the CTA's sum of that value into the receive buffer of every CTA of the
//      cluster with st.async (remote shared-memory store that signals the destination's
//      mbarrier with its byte count);
//   4. warp 0 of every CTA waits on its own mbarrier and adds the CL partials in a fixed order,
//      so all CTAs continue with bit-identical totals and nothing has to be broadcast between
//      CTAs.  Receive buffers are double-buffered by phase; the pose broadcast inside the CTA
//      (bcast) is the second and last CTA barrier of a sweep.
// Compared with barrier.cluster + DSMEM loads (ClusterReducer) this removes the cluster-scope
// MEMBAR, one DSMEM round trip and two CTA barriers from the serial path of every sweep.
// (Pushing the 64 WARP partials of a cluster directly -- no CTA barrier, no scratch -- was measured
// too: 8x the remote stores make the wait for the totals longer than the barrier it saves,
// forward 0.51 -> 0.64 ms.)
// ------------------------------------------------------------------------------------
template <int CL>
struct PushReducer {
    static constexpr int NT = kPT;
    static constexpr int kRows = 15;  // widest accumulator; 2 CTAs/SM leave room for no more
    static constexpr int kSmemFloats = kRows * NT + 2 * CL * 16 + 16 + 4;
    float* scr;   // [kRows][NT] per-thread partials
    float* buf;   // [2][CL][16] received CTA sums
    float* out;   // [16] solve result
    unsigned mbar, rank, phase;
    __device__ __forceinline__ PushReducer(float* s, unsigned rank_)
        : scr(s), buf(s + kRows * NT), out(s + kRows * NT + 2 * CL * 16), rank(rank_), phase(0) {
        mbar = smem_u32(out + 16);  // two mbarriers (8 bytes each), one per receive buffer
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 8) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_sync_all();  // every CTA of the cluster is resident and its mbarrier initialised
    }
    template <int N>
    __device__ __forceinline__ void reduce(float (&a)[N], int evt = 0, int prof_n = 64) {
        static_assert(N <= kRows, "accumulator too wide");
        (void)evt; (void)prof_n;
#ifdef QEC_PROBE_NOCHAIN  // timing probe only (wrong results): what the sweeps alone cost, no reduce / solve / broadcast
        return;
#endif
        const int tid = threadIdx.x;
        // phase n uses receive buffer / mbarrier n & 1.  A peer can be at most one phase ahead, so a
        // push of phase n + 1 never lands in the mbarrier that still collects phase n.
        const unsigned mb = mbar + 8u * (phase & 1u);
#if QEC_RED_SHFL
        // Variant: every warp folds its 32 lanes with a shuffle reduce-scatter BEFORE the CTA barrier (in the tail of
        // its own sweep), so that after the barrier only 16 x kW floats are left to add: CL lanes per value read
        // the kW warp sums in a fixed order and push.  Replaces 15 STS + 4 LDS.128 + 4 shuffles per thread.
        {
            constexpr int kW = NT / 32;
            const int lane = tid & 31, warp = tid >> 5;
            float w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] = (i < N) ? a[i] : 0.f;
            const float mine = warp_reduce_scatter16(w, lane);
            if ((lane & 1) == 0) scr[idx16(lane) * kW + warp] = mine;
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(N * CL * 4) : "memory");
            __syncthreads();
            QEC_EVT(evt);
            const int v = tid / CL, r = tid - v * CL;
            if (v < N) {
                float s = 0.f;
#pragma unroll
                for (int x = 0; x < kW; ++x) s += scr[v * kW + x];  // fixed order
                const unsigned slot = smem_u32(buf + ((phase & 1u) * CL + rank) * 16 + v);
                st_async_f32(mapa_u32(slot, r), s, mapa_u32(mb, r));
            }
        }
#else
#pragma unroll
        for (int i = 0; i < N; ++i) scr[i * NT + tid] = a[i];
        if (tid == 0)  // arm this phase: N floats from each of the CL CTAs
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(N * CL * 4) : "memory");
        __syncthreads();
        QEC_EVT(evt);
        const int v = tid >> 4, part = tid & 15;
        float s = 0.f;
        if (v < N) {
            constexpr int kPer = NT / 16;  // entries per lane: 16 (256 threads) or 32 (512 threads)
            const float4* row = reinterpret_cast<const float4*>(scr + v * NT + part * kPer);
            const int rot = (kPer == 16) ? (part >> 1) : part;
#pragma unroll
            for (int k = 0; k < kPer / 4; ++k) {  // rotated start: conflict-free LDS.128
                const float4 q = row[(rot + k) & (kPer / 4 - 1)];
                s += (q.x + q.y) + (q.z + q.w);
            }
        }
        if (evt == 2) QEC_EVT(22);
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (evt == 2) QEC_EVT(23);
        if (v < N && part < CL) {  // every lane of the group holds the sum: lane r pushes to CTA r
            const unsigned slot = smem_u32(buf + ((phase & 1u) * CL + rank) * 16 + v);
            st_async_f32(mapa_u32(slot, part), s, mapa_u32(mb, part));
        }
#endif
        QEC_EVT(evt + 1);
        if (tid < 32) {
            const unsigned parity = (phase >> 1) & 1u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@!p bra WAIT_%=;\n\t}" ::"r"(mb), "r"(parity)
                : "memory");
            if (evt == 2) QEC_EVT(24);
            float t = 0.f;
            if (tid < 16) {
                const float* b = buf + (phase & 1u) * CL * 16 + tid;
#pragma unroll
                for (int r = 0; r < CL; ++r) t += b[r * 16];  // fixed order: identical bits in every CTA
            }
#pragma unroll
            for (int i = 0; i < N; ++i) a[i] = __shfl_sync(0xffffffffu, t, i);
        }
        QEC_EVT(evt + 2);
        ++phase;
    }
#ifdef QEC_PROBE_NOCHAIN
    __device__ __forceinline__ bool solver() const { return false; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = 0.5f;
    }
    __device__ __forceinline__ void sync() {}
#else
    __device__ __forceinline__ bool solver() const { return threadIdx.x < 32; }
    template <int N>
    __device__ __forceinline__ void bcast(float (&r)[N]) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) out[i] = r[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = out[i];
    }
    // a reduction whose result only warp 0 consumes still needs the CTA to stay behind the solver
    // before scr is reused: callers follow it with bcast() or sync()
    __device__ __forceinline__ void sync() { __syncthreads(); }
#endif
};


// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
struct PairGeom {
    int np;              // pairs of this CTA
    int off[kPairs];     // k * trow + i of the pair's first vote (element offset inside the bp block of t_oqi)
    int I;
};

// issue the asynchronous copy of capsule (bp, o)'s transforms into the vote slots, one commit group per slot
// Raw transforms travel global -> shared as 8-byte cp.async into the two halves of 16-byte slots: with every lane writing
// the same half, a warp's 32 writes fall on half of the banks (2-way conflict: 7.5 M extra wavefronts per forward, 11 M
// per backward at config 2).  QEC_RAW_SWZ=1 exchanges the halves for every other group of eight lanes (bit 3 of the
// thread index: constant per thread, since slots advance by the CTA size), so that sixteen lanes cover all 32 banks;
// sweep 0 reads the raw slot back as 4 x LDS.64 from the exchanged addresses (same wavefronts as 2 x LDS.128) and
// stores the votes in the regular layout.
#ifndef QEC_RAW_SWZ
#define QEC_RAW_SWZ 0
#endif
__device__ __forceinline__ int raw_half() { return QEC_RAW_SWZ ? (int)((threadIdx.x >> 3) & 1u) * 2 : 0; }
__device__ __forceinline__ Quat2 lds_pair_raw(const float4* sA, const float4* sB, int ps) {
#if QEC_RAW_SWZ
    const int h = raw_half();
    const float* a = reinterpret_cast<const float*>(sA + ps);
    const float* b = reinterpret_cast<const float*>(sB + ps);
    Quat2 v;
    v.w = lds_f2(reinterpret_cast<const float2*>(a + h));
    v.x = lds_f2(reinterpret_cast<const float2*>(a + 2 - h));
    v.y = lds_f2(reinterpret_cast<const float2*>(b + h));
    v.z = lds_f2(reinterpret_cast<const float2*>(b + 2 - h));
    return v;
#else
    return lds_pair(sA, sB, ps);
#endif
}

template <bool FULL>
__device__ __forceinline__ void prefetch_slot(const PairGeom& G, const float* tcap, float4* sA, float4* sB, int m) {
    const int ps = threadIdx.x + m * kPT;
    if (FULL || ps < G.np) {
        const float* tp = tcap + G.off[m];
        float* a = reinterpret_cast<float*>(sA + ps);
        float* b = reinterpret_cast<float*>(sB + ps);
        const int h = raw_half();
        cp_async8(a + h, tp);
        cp_async8(a + 2 - h, tp + G.I);
        cp_async8(b + h, tp + 2 * G.I);
        cp_async8(b + 2 - h, tp + 3 * G.I);
    }
    cp_async_commit();
}

template <int M>
__device__ __forceinline__ void wait_slot() {  // slot M's group is complete (groups are committed in slot order)
    cp_async_wait<kPairs - 1 - M>();
}

// FULL: every thread owns exactly kPairs resident pairs (J == CL * 4096, e.g. pool2 of QEC-Net):
// no per-slot guards, so the unrolled sweeps are straight-line code the scheduler can interleave.
template <int CL, bool FULL>
__global__ void __launch_bounds__(kPT, kCtasPerSM)
    routing_fused2_fwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                              const float* __restrict__ act, const float* __restrict__ alpha_p,
                              const float* __restrict__ beta_p, float* __restrict__ pose,
                              float* __restrict__ agreement, float* __restrict__ poses_all,
                              float* __restrict__ stats, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* __restrict__ sA = reinterpret_cast<float4*>(smem_raw);   // {w0 w1 x0 x1} per pair
    float4* __restrict__ sB = sA + kNpMax;                           // {y0 y1 z0 z1}
    float2* __restrict__ sa = reinterpret_cast<float2*>(sB + kNpMax);  // activations of the pair
    float2* __restrict__ sb = sa + kNpMax;                           // routing weights b^t of the pair
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    PushReducer<CL> red(reinterpret_cast<float*>(sb + kNpMax), rank);
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = 2 * ((J / 2 + CL - 1) / CL);  // even slice per CTA
    const int j_begin = min(J, (int)rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const size_t trow = (size_t)4 * I * O;
    const bool writer = (rank == 0) && (threadIdx.x == 0);

    PairGeom G;
    G.np = nloc / 2;
    G.I = I;
#pragma unroll
    for (int m = 0; m < kPairs; ++m) {
        const int j = j_begin + 2 * (threadIdx.x + m * kPT);
        const int k = j / I;
        G.off[m] = (int)(k * trow) + (j - k * I);
    }
    auto tcap_of = [&](long long c) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        return t_oqi + (size_t)bp * K * trow + (size_t)o * 4 * I;
    };
    auto live = [&](int m) { return FULL || (threadIdx.x + m * kPT < G.np); };

    const long long BPn = Nc / O;
    if (cid < Nc) {
        const float* tc = tcap_of(caps_of(cid, BPn, O));
#pragma unroll
        for (int m = 0; m < kPairs; ++m) prefetch_slot<FULL>(G, tc, sA, sB, m);
    }
#ifdef QEC_PROF
    int prof_n = 0;
#endif
    bool pend = false;       // the previous capsule's epilogue sums wait for this capsule's first reduction
    long long pend_c = 0;
    float e0 = 0.f, e1 = 0.f;
    Quat p = qmake(0.f, 0.f, 0.f, 0.f);

    auto finish = [&](long long c, Quat pf, float es, float en) {  // reference qec_module.py:188-193
        const float s = (en > 0.f) ? es / en : 0.f;
        stq(pose + (size_t)c * 4, pf);
        agreement[c] = (en > 0.f) ? sigmoidf(fmaf(alpha, s, beta - 1.f)) : 0.f;
        stats[2 * c] = s;
        stats[2 * c + 1] = en;
    };

    for (long long u = cid; u < Nc; u += ncl) {
        const long long c = caps_of(u, BPn, O);
        const long long bp = c / O;
        const float* lrf_c = lrfs + ((size_t)bp * J + j_begin) * 4;
        const float* act_c = act + (size_t)bp * J + j_begin;
        // ---- sweep 0: votes from the prefetched transforms, M^0 = sum a v v^T, b^0 = a ----------
        QEC_EVT(0);
        float acc[15];
        {
            f2 m2[10], ssum = bc(0.f), comp = bc(0.f);
            float sabs = 0.f;
#pragma unroll
            for (int i = 0; i < 10; ++i) m2[i] = bc(0.f);
            auto gen_t = [&](int m, const Quat2& lrf, f2 a, const Quat2& traw) {
                const int ps = threadIdx.x + m * kPT;
                const Quat2 v = make_vote2<false>(lrf, traw).vote;
                sts_pair(sA, sB, ps, v);
                sts_f2(sa + ps, a);
                sts_f2(sb + ps, a);
                sabs = (sabs + fabsf(lo(a))) + fabsf(hi(a));
                ssum = add2(ssum, a);
                acc_rank1_2(m2, a, v);
                comp = add2(comp, add2(add2(v.w, v.x), add2(v.y, v.z)));
            };
            auto gen = [&](int m, const Quat2& lrf, f2 a) { gen_t(m, lrf, a, lds_pair_raw(sA, sB, threadIdx.x + m * kPT)); };
            auto ldl = [&](int m, Quat2& lrf, f2& a) {  // the pair's LRFs and activations (L2-resident: shared by all O)
                const int ps = threadIdx.x + m * kPT;
                if (live(m)) {
                    const Quat l0 = ldq_nc(lrf_c + (size_t)ps * 8), l1 = ldq_nc(lrf_c + (size_t)ps * 8 + 4);
                    lrf = pk(l0, l1);
                    a = ldg_f2(act_c + 2 * ps);
                }
            };
            Quat2 lr[3];  // loads run two slots ahead of their use
            f2 av[3];
            ldl(0, lr[0], av[0]);
            ldl(1, lr[1], av[1]);
#if (QEC_SWP & 2)
            // the transforms of slot M+1 leave shared memory before slot M is generated and stored
            Quat2 tr[2];
            wait_slot<0>();
            if (live(0)) tr[0] = lds_pair_raw(sA, sB, threadIdx.x);
#define QEC_GEN_SLOT(M)                                                                        \
    if ((M) + 2 < kPairs) ldl((M) + 2, lr[((M) + 2) % 3], av[((M) + 2) % 3]);                  \
    if ((M) + 1 < kPairs) {                                                                    \
        wait_slot<((M) + 1 < kPairs ? (M) + 1 : 0)>();                                         \
        if (live((M) + 1)) tr[((M) + 1) & 1] = lds_pair_raw(sA, sB, threadIdx.x + ((M) + 1) * kPT); \
    }                                                                                          \
    if (live(M)) gen_t((M), lr[(M) % 3], av[(M) % 3], tr[(M) & 1]);
#else
#define QEC_GEN_SLOT(M)                                                       \
    if ((M) + 2 < kPairs) ldl((M) + 2, lr[((M) + 2) % 3], av[((M) + 2) % 3]); \
    wait_slot<(M)>();                                                         \
    if (live(M)) gen((M), lr[(M) % 3], av[(M) % 3]);
#endif
            QEC_GEN_SLOT(0) QEC_GEN_SLOT(1) QEC_GEN_SLOT(2) QEC_GEN_SLOT(3)
            QEC_GEN_SLOT(4) QEC_GEN_SLOT(5) QEC_GEN_SLOT(6) QEC_GEN_SLOT(7)
#undef QEC_GEN_SLOT
#pragma unroll
            for (int i = 0; i < 10; ++i) acc[i] = hsum(m2[i]);
            acc[10] = sabs;
            acc[11] = hsum(ssum);
            acc[12] = hsum(comp);
            acc[13] = e0;
            acc[14] = e1;
        }
        QEC_EVT(1);
#ifdef QEC_PROF
        red.reduce(acc, 2, prof_n);
#else
        red.reduce(acc);
#endif
        bool valid = false;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (red.solver()) {
            if (pend && writer) finish(pend_c, p, acc[13], acc[14]);
            valid = acc[12] != 0.f;  // reference qec_module.py:72-73
            const Quat q = solve_pose2(acc, valid);
            r[0] = q.w; r[1] = q.x; r[2] = q.y; r[3] = q.z;
        }
        QEC_EVT(5);
        red.bcast(r);
        QEC_EVT(6);
        p = qmake(r[0], r[1], r[2], r[3]);
        if (writer) stq(poses_all + (size_t)c * 4, p);
        // ---- sweeps 1..T-1: b <- b a (1 - d(pose, v)), M^t = sum b v v^T, all from shared memory
        for (int t = 1; t < T; ++t) {
            f2 m2[10], ssum = bc(0.f);
            float sabs = 0.f;
#pragma unroll
            for (int i = 0; i < 10; ++i) m2[i] = bc(0.f);
#if (QEC_SWP & 1)
            Quat2 vn;
            f2 an = bc(0.f), bn = bc(0.f);
            auto lds_slot = [&](int m) {
                const int ps = threadIdx.x + m * kPT;
                if (live(m)) {
                    vn = lds_pair(sA, sB, ps);
                    an = lds_f2(sa + ps);
                    bn = lds_f2(sb + ps);
                }
            };
            lds_slot(0);
#pragma unroll
            for (int m = 0; m < kPairs; ++m) {
                const int ps = threadIdx.x + m * kPT;
                const Quat2 v = vn;
                const f2 a = an, b0 = bn;
                if (m + 1 < kPairs) lds_slot(m + 1);  // before this slot's store: the next chain starts under this one
                if (live(m)) {
                    const f2 b = mul2(b0, mul2(a, one_minus_dist2(qdot2(p, v))));
                    sts_f2(sb + ps, b);
                    sabs = (sabs + fabsf(lo(b))) + fabsf(hi(b));
                    ssum = add2(ssum, b);
                    acc_rank1_2(m2, b, v);
                }
            }
#else
#pragma unroll
            for (int m = 0; m < kPairs; ++m) {
                const int ps = threadIdx.x + m * kPT;
                if (live(m)) {
                    const Quat2 v = lds_pair(sA, sB, ps);
                    const f2 a = lds_f2(sa + ps), b0 = lds_f2(sb + ps);
                    const f2 b = mul2(b0, mul2(a, one_minus_dist2(qdot2(p, v))));
                    sts_f2(sb + ps, b);
                    sabs = (sabs + fabsf(lo(b))) + fabsf(hi(b));
                    ssum = add2(ssum, b);
                    acc_rank1_2(m2, b, v);
                }
            }
#endif
            float a2[12];
#pragma unroll
            for (int i = 0; i < 10; ++i) a2[i] = hsum(m2[i]);
            a2[10] = sabs;
            a2[11] = hsum(ssum);
            QEC_EVT(7 * t);
#ifdef QEC_PROF
            red.reduce(a2, 7 * t + 1, prof_n);
#else
            red.reduce(a2);
#endif
            if (red.solver()) {
                const Quat q = solve_pose2(a2, valid);
                r[0] = q.w; r[1] = q.x; r[2] = q.y; r[3] = q.z;
            }
            QEC_EVT(7 * t + 4);
            red.bcast(r);
            QEC_EVT(7 * t + 5);
            p = qmake(r[0], r[1], r[2], r[3]);
            if (writer) stq(poses_all + ((size_t)t * Nc + c) * 4, p);
        }
        // ---- epilogue sweep (reference qec_module.py:188-193); each slot is refilled with the next
        // capsule's transforms as soon as its vote has been read
        {
            const bool more = u + ncl < Nc;
            const float* tn = more ? tcap_of(caps_of(u + ncl, BPn, O)) : t_oqi;
            f2 es = bc(0.f);
            float en = 0.f;
#if (QEC_SWP & 4)
            Quat2 vn;
            f2 bn = bc(0.f);
            auto lds_slot = [&](int m) {
                const int ps = threadIdx.x + m * kPT;
                if (live(m)) {
                    vn = lds_pair(sA, sB, ps);
                    bn = lds_f2(sb + ps);
                }
            };
            lds_slot(0);
#pragma unroll
            for (int m = 0; m < kPairs; ++m) {
                const Quat2 v = vn;
                const f2 b = bn;
                if (m + 1 < kPairs) lds_slot(m + 1);  // ahead of the cp.async below (a compiler barrier)
                if (more) prefetch_slot<FULL>(G, tn, sA, sB, m);  // slot m is in registers: refill it right away
                if (live(m)) {
                    es = fma2(b, one_minus_dist2(qdot2(p, v)), es);
                    en += ((lo(b) != 0.f) ? 1.f : 0.f) + ((hi(b) != 0.f) ? 1.f : 0.f);
                }
            }
#else
#pragma unroll
            for (int m = 0; m < kPairs; ++m) {
                const int ps = threadIdx.x + m * kPT;
                if (live(m)) {
                    const Quat2 v = lds_pair(sA, sB, ps);
                    const f2 b = lds_f2(sb + ps);
                    es = fma2(b, one_minus_dist2(qdot2(p, v)), es);
                    en += ((lo(b) != 0.f) ? 1.f : 0.f) + ((hi(b) != 0.f) ? 1.f : 0.f);
                }
                if (more) prefetch_slot<FULL>(G, tn, sA, sB, m);
            }
#endif
            e0 = hsum(es);
            e1 = en;
            pend = true;
            pend_c = c;
            QEC_EVT(21);
            QEC_EVT_NEXT();
        }
    }
    if (pend) {
        float e[2] = {e0, e1};
        red.reduce(e);
        if (red.solver() && writer) finish(pend_c, p, e[0], e[1]);
    }
    cluster_sync_all();  // no CTA may exit while a peer can still push into its receive buffers
}

template <int CL>
constexpr size_t fwd2_smem() {
    return (2 * sizeof(float4) + 2 * sizeof(float2)) * kNpMax + sizeof(float) * (PushReducer<CL>::kSmemFloats + 32);
}

// like fused_grid (routing_fused_shared.cuh) but the cluster attribute is set for CL = 1 too: the
// push reducer addresses its own CTA through the cluster window as well
template <class Kern>
static int fused2_grid(Kern kern, int CL, size_t smem, long long Nc, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr,
                       PerDeviceInt& cache) {
    cfg.blockDim = dim3(kPT);
    cfg.dynamicSmemBytes = smem;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
    cfg.attrs = attr;
    std::atomic<int>* slot = cache.slot();  // per device: the shared-memory opt-in below is a per-device attribute
    int clusters = slot ? slot->load(std::memory_order_relaxed) : -1;
    if (clusters < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cfg.gridDim = dim3(CL * 4 * kNumSMs);  // upper bound for the occupancy query
        e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
        if (e != cudaSuccess) return (int)e;
        if (clusters < 1) return (int)cudaErrorLaunchOutOfResources;
        if (slot) slot->store(clusters, std::memory_order_relaxed);
    }
    if (clusters > Nc) clusters = (int)Nc;
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    return 0;
}

template <int CL>
static int launch_fused2_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                             const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                             long long Nc, int K, int I, int O, int T, cudaStream_t st) {
    const bool full = ((long long)K * I == (long long)CL * 2 * kNpMax);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static PerDeviceInt cached[2];
    cudaError_t e;
    if (full) {
        auto kern = routing_fused2_fwd_kernel<CL, true>;
        const int rc = fused2_grid(kern, CL, fwd2_smem<CL>(), Nc, cfg, attr, cached[1]);
        if (rc) return rc;
        cfg.stream = st;
        e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I,
                               O, T);
    } else {
        auto kern = routing_fused2_fwd_kernel<CL, false>;
        const int rc = fused2_grid(kern, CL, fwd2_smem<CL>(), Nc, cfg, attr, cached[0]);
        if (rc) return rc;
        cfg.stream = st;
        e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I,
                               O, T);
    }
    return e == cudaSuccess ? launch_status() : (int)e;
}


// ------------------------------------------------------------------------------------
// backward (T <= kTB = 3), packed twin of routing_fused_bwd_kernel (routing_fused.cu; algebra in
// routing_core.cuh / SURVEY.md App. B).  Levels t = T-1..0; sweep k completes level TT = T-k (its
// adjoint vector u^TT is known from the previous reduction) and accumulates level TT-1.
// Shared memory per pair: the two votes (32 B) and the cached 1 - d^t of the levels below the top
// (16 B).  dL/db^TT is rebuilt per sweep from per-capsule registers by walking down from the top
// level.  The transforms of the next capsule are prefetched into the vote slots during the last
// sweep, as in the forward.
// ------------------------------------------------------------------------------------
// element offsets instead of six 64-bit pointers per capsule: the bases stay in the kernel-parameter bank
// (offA and the in-block walker are 32-bit: qec_routing_fused_packed_supported bounds one bp block)
struct Bwd2Geom {
    const float* act;
    const float* lrfs;
    const float* t_oqi;
    float* g_toqi;
    unsigned short* g_lo;  // nullable (bf16): if set, g_toqi receives the TF32-rounded gradient, g_lo the remainder
    float* g_lrfs;   // nullable
    float* g_act;    // nullable: no activation gradient -> only the top level is processed
    unsigned offA;   // bp * J + j_begin: first vote of this CTA's slice in act / g_act (x4 in lrfs / g_lrfs)
    size_t offT;     // bp * K * trow + o * 4I: the capsule's block in t_oqi / g_toqi (64-bit: > 2^32 elements at large B)
    int I, dko, di, wrap;  // pair walker: per slot j += 2 * kPT
};

// (k * trow + i) of the pair in slot m, advanced slot by slot (j += 2 * kPT) without a division
struct OffWalk {
    int off, i;
    __device__ __forceinline__ void next(const Bwd2Geom& B) {
        i += B.di;
        off += B.dko;
        if (i >= B.I) {
            i -= B.I;
            off += B.wrap;
        }
    }
};

template <bool FULL>
__device__ __forceinline__ void prefetch_walk(const Bwd2Geom& B, size_t offT_next, OffWalk& w, float4* sA, float4* sB,
                                              int np, int m) {
    const int ps = threadIdx.x + m * kPT;
    if (FULL || ps < np) {
        const float* tp = B.t_oqi + (offT_next + (size_t)w.off);
        float* a = reinterpret_cast<float*>(sA + ps);
        float* b = reinterpret_cast<float*>(sB + ps);
        const int h = raw_half();
        cp_async8(a + h, tp);
        cp_async8(a + 2 - h, tp + B.I);
        cp_async8(b + h, tp + 2 * B.I);
        cp_async8(b + 2 - h, tp + 3 * B.I);
    }
    cp_async_commit();
    w.next(B);
}

// ---- top level, pass 1: dL/dvote of the level whose votes are not detached -> dL/dt_raw, dL/dlrf.
// No accumulators are live here; the accumulation of level T-2 is a separate pass (bwd2_acc_sweep),
// which keeps both under the 128-register budget of two CTAs per SM.
// Without PREFETCH (the chained backward) the pair's transforms and activations are STAGED: while slot m is
// computed, slot m + 1's 40 bytes per thread travel global -> shared by cp.async into the reducer's scratch,
// which is idle between two reductions (each thread reads back only what it copied itself: no barrier inside the pass,
// one at its end -- see there).
// One pair in flight leaves ~100 instructions between a direct load and its first use, and no registers
// for a second pair.
template <int T, bool FULL, bool PREFETCH>
__device__ __forceinline__ void bwd2_grad_pass(int np, const Bwd2Geom& B, const Bwd2Caps& C, OffWalk w0, float4* sA,
                                               float4* sB, size_t offT_next, float* stage_f) {
    constexpr int TT = T - 1;
    constexpr bool STAGE = !PREFETCH;  // PREFETCH owns the cp.async group count of the next capsule's sweep 0
    float2* som0 = reinterpret_cast<float2*>(sB + kNpMax);
    float2* som1 = som0 + kNpMax;
    float2* stage = reinterpret_cast<float2*>(stage_f) + threadIdx.x;  // [5][kPT] float2: t rows w,x,y,z and a
    OffWalk w = w0, wp = w0, ws = w0;
    auto stage_slot = [&](int m) {  // ws walks one slot ahead of w
        const int ps = threadIdx.x + m * kPT;
        if (FULL || ps < np) {
            const float* tp = B.t_oqi + (B.offT + (size_t)ws.off);
            cp_async8(stage, tp);
            cp_async8(stage + kPT, tp + B.I);
            cp_async8(stage + 2 * kPT, tp + 2 * B.I);
            cp_async8(stage + 3 * kPT, tp + 3 * B.I);
            cp_async8(stage + 4 * kPT, B.act + (size_t)B.offA + 2 * ps);
        }
        cp_async_commit();
        ws.next(B);
    };
    if (STAGE) stage_slot(0);
#pragma unroll 1  // one pair in flight: ~70 live registers per pair leave no room to overlap two
    for (int m = 0; m < kPairs; ++m) {
        const int ps = threadIdx.x + m * kPT;
        if (STAGE) cp_async_wait<0>();
        if (FULL || ps < np) {
            Quat2 traw;
            f2 a;
            const float* lp = B.lrfs + ((size_t)B.offA + 2 * ps) * 4;
            const Quat2 lrf = pk(ldq_nc(lp), ldq_nc(lp + 4));
            if (STAGE) {
                traw.w = lds_f2(stage); traw.x = lds_f2(stage + kPT); traw.y = lds_f2(stage + 2 * kPT); traw.z = lds_f2(stage + 3 * kPT);
                a = lds_f2(stage + 4 * kPT);
            } else {
                // issued first: the level walk below hides part of their latency
                const float* tp = B.t_oqi + (B.offT + (size_t)w.off);
                traw.w = ldg_f2(tp); traw.x = ldg_f2(tp + B.I); traw.y = ldg_f2(tp + 2 * B.I); traw.z = ldg_f2(tp + 3 * B.I);
                a = ldg_f2(B.act + (size_t)B.offA + 2 * ps);
            }
            if (STAGE && m + 1 < kPairs) stage_slot(m + 1);  // after this slot's operands have been read back
            const Quat2 v = lds_pair(sA, sB, ps);
            f2 b = a;
            if (T >= 2) b = mul2(b, mul2(a, lds_f2(som0 + ps)));
            if (T >= 3) b = mul2(b, mul2(a, lds_f2(som1 + ps)));
            const f2 x = qdot2(C.ph[TT], v);
            const f2 uv = qdot2(C.uh[TT], v);
            const f2 gx = mul2(mul2(bc(-C.gsn), b), dist_grad2(x));
            const f2 wh = mul2(b, bc(C.rS[TT]));
            Quat2 gv = qscale2(C.ph[TT], fma2(wh, uv, gx));  // p (gx + w <u,v>)
            gv = qfma2(mul2(wh, x), C.uh[TT], gv);           // + u w <p,v>
            const VotePair vp = make_vote2<true>(lrf, traw);
            Quat2 gt, gq;
            vote_bwd2(lrf, vp, gv, gt, gq);
            float* gp = B.g_toqi + (B.offT + (size_t)w.off);
            if (B.g_lo != nullptr) {
                // pre-split for the tensor-core GEMMs that consume it (quater_gen2's backward at FP32
                // accuracy): hi on the TF32 grid (an exact tensor-core input); the remainder g - hi (exact in
                // FP32, at most 2^-11 |g|) only has to carry ~11 more bits and travels as BF16
                unsigned* gq2 = reinterpret_cast<unsigned*>(B.g_lo + (B.offT + (size_t)w.off));
                const f2 hw = tf32_round2(gt.w), hx = tf32_round2(gt.x), hy = tf32_round2(gt.y), hz = tf32_round2(gt.z);
                const int Ih = B.I >> 1;  // pair stride in 32-bit words
                gq2[0] = bf16x2(fma2(hw, bc(-1.f), gt.w));
                gq2[Ih] = bf16x2(fma2(hx, bc(-1.f), gt.x));
                gq2[2 * Ih] = bf16x2(fma2(hy, bc(-1.f), gt.y));
                gq2[3 * Ih] = bf16x2(fma2(hz, bc(-1.f), gt.z));
                gt.w = hw; gt.x = hx; gt.y = hy; gt.z = hz;
            }
            *reinterpret_cast<float2*>(gp) = make_float2(lo(gt.w), hi(gt.w));
            *reinterpret_cast<float2*>(gp + B.I) = make_float2(lo(gt.x), hi(gt.x));
            *reinterpret_cast<float2*>(gp + 2 * B.I) = make_float2(lo(gt.y), hi(gt.y));
            *reinterpret_cast<float2*>(gp + 3 * B.I) = make_float2(lo(gt.z), hi(gt.z));
            if (B.g_lrfs != nullptr) {  // sum over the O capsules of the patch
                float* glp = B.g_lrfs + ((size_t)B.offA + 2 * ps) * 4;
                red_add_q(glp, lo(gq));
                red_add_q(glp + 4, hi(gq));
            }
        }
        w.next(B);
        if (PREFETCH) prefetch_walk<FULL>(B, offT_next, wp, sA, sB, np, m);
    }
    // The staging area is the reducer's scratch, and the next PushReducer::reduce() parks every thread's partial sums
    // there BEFORE its own barrier.  The float2 staging columns alias OTHER threads' partial-sum words, so a warp that
    // ran a whole accumulation sweep ahead of a straggler (one cp.async of the straggler stuck behind a TLB miss is
    // enough) overwrote the straggler's staged transforms: one k x 32 channels x 4 components of dL/dt garbage, about
    // once per 1 000 launches at config 2.  Found and verified with scripts/race_hunt_bwd.py (0 deviations in 15 000
    // launches with the barrier).  Measured alternatives: an mbarrier guard (arrive here, wait in reduce()): the same
    // +13 us per backward, more live registers; staging word by word into the thread's OWN partial-sum words (no
    // synchronisation at all): correct, but ten 4-byte cp.async + LDS per slot instead of five: +50 us.
    if (STAGE) __syncthreads();
}

// ---- sweep that completes level TT (u^TT known) and accumulates level TT-1 (chain only)
template <int T, int TT, bool FULL, bool PREFETCH>
__device__ __forceinline__ void bwd2_acc_sweep(int np, const Bwd2Geom& B, const Bwd2Caps& C, OffWalk w0, float4* sA,
                                               float4* sB, size_t offT_next, float (&acc)[15]) {
    float2* som0 = reinterpret_cast<float2*>(sB + kNpMax);
    float2* som1 = som0 + kNpMax;
    f2 a2[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) a2[i] = bc(0.f);
    float sabs = 0.f;
    OffWalk wp = w0;
    auto lda = [&](int m, f2& a) {  // activations run ahead of their use (L2-resident: shared by all O)
        const int ps = threadIdx.x + m * kPT;
        if (FULL || ps < np) a = ldg_f2(B.act + (size_t)B.offA + 2 * ps);
    };
    f2 an[3];  // two slots ahead: one was not enough under the prefetch traffic of the last sweep
    lda(0, an[0]);
    lda(1, an[1]);
#if (QEC_SWP & 8)
    Quat2 vn;
    f2 o0n = bc(1.f), o1n = bc(1.f);
    auto lds_slot = [&](int m) {
        const int ps = threadIdx.x + m * kPT;
        if (FULL || ps < np) {
            vn = lds_pair(sA, sB, ps);
            if (T >= 2) o0n = lds_f2(som0 + ps);
            if (T >= 3) o1n = lds_f2(som1 + ps);
        }
    };
    lds_slot(0);
#pragma unroll
    for (int m = 0; m < kPairs; ++m) {
        if (m + 2 < kPairs) lda(m + 2, an[(m + 2) % 3]);
        const int ps = threadIdx.x + m * kPT;
        const Quat2 v = vn;
        const f2 om0 = o0n, om1 = o1n;
        if (m + 1 < kPairs) lds_slot(m + 1);  // ahead of this slot's red / cp.async (compiler barriers)
        if (FULL || ps < np) {
            const f2 a = an[m % 3];
            const f2 gate = pk((lo(a) != 0.f) ? 1.f : 0.f, (hi(a) != 0.f) ? 1.f : 0.f);
            f2 gb, ga, x, uv, b[kTB];
            complete_levels2<T, TT>(C, v, a, gate, om0, om1, gb, ga, x, uv, b);
            if (TT >= 1) {  // level TT-1: dL/dd_j = -g_b^TT b^{TT-1} a
                const f2 xl = qdot2(C.ph[TT >= 1 ? TT - 1 : 0], v);
                const f2 bl = b[TT >= 1 ? TT - 1 : 0];
                const f2 gx = mul2(mul2(mul2(gb, bc(-1.f)), mul2(bl, a)), dist_grad2(xl));
                a2[0] = fma2(gx, v.w, a2[0]); a2[1] = fma2(gx, v.x, a2[1]);
                a2[2] = fma2(gx, v.y, a2[2]); a2[3] = fma2(gx, v.z, a2[3]);
                acc_rank1_2(a2 + 4, bl, v);
                sabs = (sabs + fabsf(lo(bl))) + fabsf(hi(bl));
            }
            if (TT == 0) {  // + dL/db^0 (b^0 = a)
                const f2 g2 = add2(ga, gb);
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(B.g_act + (size_t)B.offA + 2 * ps),
                             "f"(lo(g2)), "f"(hi(g2))
                             : "memory");
            }
        }
        if (PREFETCH) prefetch_walk<FULL>(B, offT_next, wp, sA, sB, np, m);
    }
#else
#pragma unroll
    for (int m = 0; m < kPairs; ++m) {
        if (m + 2 < kPairs) lda(m + 2, an[(m + 2) % 3]);
        const int ps = threadIdx.x + m * kPT;
        if (FULL || ps < np) {
            const f2 a = an[m % 3];
            const Quat2 v = lds_pair(sA, sB, ps);
            const f2 om0 = (T >= 2) ? lds_f2(som0 + ps) : bc(1.f);
            const f2 om1 = (T >= 3) ? lds_f2(som1 + ps) : bc(1.f);
            const f2 gate = pk((lo(a) != 0.f) ? 1.f : 0.f, (hi(a) != 0.f) ? 1.f : 0.f);
            f2 gb, ga, x, uv, b[kTB];
            complete_levels2<T, TT>(C, v, a, gate, om0, om1, gb, ga, x, uv, b);
            if (TT >= 1) {  // level TT-1: dL/dd_j = -g_b^TT b^{TT-1} a
                const f2 xl = qdot2(C.ph[TT >= 1 ? TT - 1 : 0], v);
                const f2 bl = b[TT >= 1 ? TT - 1 : 0];
                const f2 gx = mul2(mul2(mul2(gb, bc(-1.f)), mul2(bl, a)), dist_grad2(xl));
                a2[0] = fma2(gx, v.w, a2[0]); a2[1] = fma2(gx, v.x, a2[1]);
                a2[2] = fma2(gx, v.y, a2[2]); a2[3] = fma2(gx, v.z, a2[3]);
                acc_rank1_2(a2 + 4, bl, v);
                sabs = (sabs + fabsf(lo(bl))) + fabsf(hi(bl));
            }
            if (TT == 0) {  // + dL/db^0 (b^0 = a)
                const f2 g2 = add2(ga, gb);
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(B.g_act + (size_t)B.offA + 2 * ps),
                             "f"(lo(g2)), "f"(hi(g2))
                             : "memory");
            }
        }
        if (PREFETCH) prefetch_walk<FULL>(B, offT_next, wp, sA, sB, np, m);
    }
#endif
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = hsum(a2[i]);
    acc[14] = sabs;
}

// sweep 0: generate the votes (from the prefetched transforms), cache 1 - d^t of the levels below the
// top, accumulate level T-1
template <int T, bool FULL>
__device__ __forceinline__ void bwd2_sweep0(int np, const Bwd2Geom& B, const Bwd2Caps& C, float4* sA, float4* sB,
                                            float (&acc)[15]) {
    float2* som0 = reinterpret_cast<float2*>(sB + kNpMax);
    float2* som1 = som0 + kNpMax;
    f2 a2[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) a2[i] = bc(0.f);
    float sabs = 0.f;
    auto ldl = [&](int m, Quat2& lrf, f2& a) {
        const int ps = threadIdx.x + m * kPT;
        if (FULL || ps < np) {
            const float* lp = B.lrfs + ((size_t)B.offA + 2 * ps) * 4;
            lrf = pk(ldq_nc(lp), ldq_nc(lp + 4));
            a = ldg_f2(B.act + (size_t)B.offA + 2 * ps);
        }
    };
    auto gen_t = [&](int m, const Quat2& lrf, f2 a, const Quat2& traw) {
        const int ps = threadIdx.x + m * kPT;
        const Quat2 v = make_vote2<false>(lrf, traw).vote;
        sts_pair(sA, sB, ps, v);
        f2 b = a;
        if (T >= 2) {
            const f2 om = one_minus_dist2(qdot2(C.ph[0], v));
            sts_f2(som0 + ps, om);
            b = mul2(b, mul2(a, om));
        }
        if (T >= 3) {
            const f2 om = one_minus_dist2(qdot2(C.ph[1], v));
            sts_f2(som1 + ps, om);
            b = mul2(b, mul2(a, om));
        }
        // level T-1 (final distance): dL/dd_j = -gsn b^{T-1}_j
        const f2 gx = mul2(mul2(bc(-C.gsn), b), dist_grad2(qdot2(C.ph[T - 1], v)));
        a2[0] = fma2(gx, v.w, a2[0]); a2[1] = fma2(gx, v.x, a2[1]);
        a2[2] = fma2(gx, v.y, a2[2]); a2[3] = fma2(gx, v.z, a2[3]);
        acc_rank1_2(a2 + 4, b, v);
        sabs = (sabs + fabsf(lo(b))) + fabsf(hi(b));
    };
    Quat2 lr[3];
    f2 av[3];
    ldl(0, lr[0], av[0]);
    ldl(1, lr[1], av[1]);
#if (QEC_SWP & 16)
    Quat2 tr[2];  // the transforms of slot M+1 leave shared memory before slot M is generated and stored
    wait_slot<0>();
    if (FULL || threadIdx.x < np) tr[0] = lds_pair_raw(sA, sB, threadIdx.x);
#define QEC_GEN_SLOT(M)                                                                              \
    if ((M) + 2 < kPairs) ldl((M) + 2, lr[((M) + 2) % 3], av[((M) + 2) % 3]);                        \
    if ((M) + 1 < kPairs) {                                                                          \
        wait_slot<((M) + 1 < kPairs ? (M) + 1 : 0)>();                                               \
        if (FULL || (threadIdx.x + ((M) + 1) * kPT < np)) tr[((M) + 1) & 1] = lds_pair_raw(sA, sB, threadIdx.x + ((M) + 1) * kPT); \
    }                                                                                                \
    if (FULL || (threadIdx.x + (M) * kPT < np)) gen_t((M), lr[(M) % 3], av[(M) % 3], tr[(M) & 1]);
#else
#define QEC_GEN_SLOT(M)                                                       \
    if ((M) + 2 < kPairs) ldl((M) + 2, lr[((M) + 2) % 3], av[((M) + 2) % 3]); \
    wait_slot<(M)>();                                                         \
    if (FULL || (threadIdx.x + (M) * kPT < np)) gen_t((M), lr[(M) % 3], av[(M) % 3], lds_pair_raw(sA, sB, threadIdx.x + (M) * kPT));
#endif
    QEC_GEN_SLOT(0) QEC_GEN_SLOT(1) QEC_GEN_SLOT(2) QEC_GEN_SLOT(3)
    QEC_GEN_SLOT(4) QEC_GEN_SLOT(5) QEC_GEN_SLOT(6) QEC_GEN_SLOT(7)
#undef QEC_GEN_SLOT
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = hsum(a2[i]);
    acc[14] = sabs;
}

template <int CL, bool FULL>
__global__ void __launch_bounds__(kPT, kCtasPerSM)
    routing_fused2_bwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs,
                              const float* __restrict__ act, const float* __restrict__ alpha_p,
                              const float* __restrict__ beta_p, const float* __restrict__ poses_all,
                              const float* __restrict__ stats, const float* __restrict__ g_pose,
                              const float* __restrict__ g_agr, float* __restrict__ g_toqi, unsigned short* g_lo, float* g_lrfs,
                              float* g_act, float* g_alpha_beta, long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* __restrict__ sA = reinterpret_cast<float4*>(smem_raw);
    float4* __restrict__ sB = sA + kNpMax;
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    float* red_base = reinterpret_cast<float*>(smem_raw + (2 * sizeof(float4) + 2 * sizeof(float2)) * kNpMax);
    PushReducer<CL> red(red_base, rank);
    float* capsS = red_base + PushReducer<CL>::kSmemFloats;  // [kTB][8]: u^t (4), 1/S^t
    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const int J = K * I;
    const int Jc = 2 * ((J / 2 + CL - 1) / CL);
    const int j_begin = min(J, (int)rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const int np = nloc / 2;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    const unsigned trow = 4u * I * O;
    const bool need_chain = g_act != nullptr;
    float ga_tot = 0.f, gb_tot = 0.f;
#ifdef QEC_PROF
    int prof_n = 0;
#endif

    Bwd2Geom B;
    B.act = act; B.lrfs = lrfs; B.t_oqi = t_oqi; B.g_toqi = g_toqi; B.g_lo = g_lo; B.g_lrfs = g_lrfs; B.g_act = g_act;
    B.I = I;
    {
        const int step = 2 * kPT, dk = step / I;
        B.di = step - dk * I;
        B.dko = dk * (int)trow + B.di;
        B.wrap = (int)trow - I;
    }
    OffWalk w0;
    {
        const int j = j_begin + 2 * threadIdx.x;
        const int k = j / I;
        w0.i = j - k * I;
        w0.off = k * (int)trow + w0.i;
    }
    auto offT_of = [&](long long c) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        return (size_t)bp * K * trow + (size_t)o * 4 * I;
    };
    const long long BPn = Nc / O;
    if (cid < Nc) {
        OffWalk w = w0;
        const size_t o0 = offT_of(caps_of(cid, BPn, O));
#pragma unroll
        for (int m = 0; m < kPairs; ++m) prefetch_walk<FULL>(B, o0, w, sA, sB, np, m);
    }

    for (long long u = cid; u < Nc; u += ncl) {
        const long long c = caps_of(u, BPn, O);
        const long long bp = c / O;
        B.offA = (unsigned)bp * (unsigned)J + (unsigned)j_begin;
        B.offT = offT_of(c);
        const bool more = u + ncl < Nc;
        const size_t offT_next = more ? offT_of(caps_of(u + ncl, BPn, O)) : (size_t)0;
        // Per-capsule state lives in memory between sweeps, not in registers: the poses are re-read from
        // poses_all (L2) and the adjoint vectors u^t / 1/S^t from a 32-float shared array, so that each
        // sweep only keeps what it uses live (the 28 registers of the full state made the sweeps spill).
        const float s = __ldg(stats + 2 * c), n = __ldg(stats + 2 * c + 1);
        const float g_agr_c = __ldg(g_agr + c);  // unconditional: one L2 round trip for all three, not two in series
        const bool has = n > 0.f;
        const float sg = sigmoidf(fmaf(alpha, s, beta - 1.f));
        const float gz = has ? g_agr_c * sg * (1.f - sg) : 0.f;
        ga_tot += gz * s;
        gb_tot += gz;
        const float gsn = has ? gz * alpha / n : 0.f;
        auto caps = [&](int t_lo, int u_lo) {  // poses of levels >= t_lo, adjoints of levels >= u_lo
            Bwd2Caps C;
#pragma unroll
            for (int t = 0; t < kTB; ++t) {
                C.ph[t] = (t < T && t >= t_lo) ? ldq_nc(poses_all + ((size_t)t * Nc + c) * 4) : qmake(0.f, 0.f, 0.f, 0.f);
                const bool hu = (t < T && t >= u_lo);
                C.uh[t] = hu ? qmake(capsS[t * 8], capsS[t * 8 + 1], capsS[t * 8 + 2], capsS[t * 8 + 3])
                             : qmake(0.f, 0.f, 0.f, 0.f);
                C.rS[t] = hu ? capsS[t * 8 + 4] : 0.f;
            }
            C.gsn = gsn;
            return C;
        };

        float acc[15];
        QEC_EVT(0);
        {
            const Bwd2Caps C = caps(0, kTB);
            if (T == 3) bwd2_sweep0<3, FULL>(np, B, C, sA, sB, acc);
            else if (T == 2) bwd2_sweep0<2, FULL>(np, B, C, sA, sB, acc);
            else bwd2_sweep0<1, FULL>(np, B, C, sA, sB, acc);
        }

        const int nsweeps = need_chain ? T : 1;
        for (int k = 1; k <= nsweeps; ++k) {
            const int t = T - k;  // level being completed
            QEC_EVT(7 * (k - 1) + 1);
            // the solver's own operands are requested before the reduction, not on the serial path behind it
            // (an L2 round trip per chain: backward 1.32 -> 1.28 ms at config 2)
            Quat p = qmake(0.f, 0.f, 0.f, 0.f), gp_top = qmake(0.f, 0.f, 0.f, 0.f);
            if (red.solver()) {
                p = ldq_nc(poses_all + ((size_t)t * Nc + c) * 4);
                if (t == T - 1) gp_top = ldq_nc(g_pose + (size_t)c * 4);
            }
#ifdef QEC_PROF
            red.reduce(acc, 7 * (k - 1) + 2, prof_n);
#else
            red.reduce(acc);      // sums of the level accumulated by the previous sweep
#endif
            if (red.solver()) {
                const float rs = 1.f / fmaxf(acc[14], 1e-12f);
                const Sym4 m = acc_to_sym(acc + 4, rs);
                Quat gp = qmake(acc[0], acc[1], acc[2], acc[3]);
                if (t == T - 1) gp = qadd(gp, gp_top);
                const float lam = qdot(p, sym_mul(m, p));
                Quat u = top_eigvec_adjoint(m, p, lam, gp);
                if (!(qdot(p, p) > 0.5f)) u = qmake(0.f, 0.f, 0.f, 0.f);  // invalid / sign-zeroed pose
                if (threadIdx.x == 0) {
                    capsS[t * 8] = u.w; capsS[t * 8 + 1] = u.x; capsS[t * 8 + 2] = u.y; capsS[t * 8 + 3] = u.z;
                    capsS[t * 8 + 4] = rs;
                }
            }
            QEC_EVT(7 * (k - 1) + 5);
            red.sync();  // u^t, 1/S^t visible to the CTA; the reducer's scratch may be reused
            QEC_EVT(7 * (k - 1) + 6);
            // the votes die with the last pass over them: that pass refills the slots for the next capsule
            const bool pf_grad = more && !need_chain;
            const bool pf_acc = more && (k == nsweeps);
#define QEC_BWD_GRAD(T_)                                                                       \
    {                                                                                          \
        const Bwd2Caps C = caps(T_ - 1, T_ - 1);                                               \
        if (pf_grad) bwd2_grad_pass<T_, FULL, true>(np, B, C, w0, sA, sB, offT_next, red.scr);          \
        else bwd2_grad_pass<T_, FULL, false>(np, B, C, w0, sA, sB, offT_next, red.scr);                 \
        QEC_EVT(7);                                                                            \
    }
#define QEC_BWD_ACC(TT_, T_)                                                                   \
    {                                                                                          \
        const Bwd2Caps C = caps(TT_ >= 1 ? TT_ - 1 : 0, TT_);                                  \
        if (pf_acc) bwd2_acc_sweep<T_, TT_, FULL, true>(np, B, C, w0, sA, sB, offT_next, acc); \
        else bwd2_acc_sweep<T_, TT_, FULL, false>(np, B, C, w0, sA, sB, offT_next, acc);       \
    }
            switch (T * 4 + t) {
                case 3 * 4 + 2: QEC_BWD_GRAD(3) if (need_chain) QEC_BWD_ACC(2, 3) break;
                case 3 * 4 + 1: QEC_BWD_ACC(1, 3) break;
                case 3 * 4 + 0: QEC_BWD_ACC(0, 3) break;
                case 2 * 4 + 1: QEC_BWD_GRAD(2) if (need_chain) QEC_BWD_ACC(1, 2) break;
                case 2 * 4 + 0: QEC_BWD_ACC(0, 2) break;
                default: QEC_BWD_GRAD(1) if (need_chain) QEC_BWD_ACC(0, 1) break;
            }
#undef QEC_BWD_GRAD
#undef QEC_BWD_ACC
        }
        QEC_EVT(25);
        QEC_EVT_NEXT();
        // the next capsule's sweep 0 touches this thread's own slots only; the reducer's scratch is
        // protected by the barriers inside reduce() / bcast()
    }
    if (rank == 0 && threadIdx.x == 0) {
        atomicAdd(g_alpha_beta, ga_tot);
        atomicAdd(g_alpha_beta + 1, gb_tot);
    }
    cluster_sync_all();
}

template <int CL>
static int launch_fused2_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                             const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                             const float* g_agr, float* g_toqi, unsigned short* g_lo, float* g_lrfs, float* g_act,
                             float* g_ab,
                             long long Nc, int K, int I, int O, int T, cudaStream_t st) {
    const bool full = ((long long)K * I == (long long)CL * 2 * kNpMax);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    static PerDeviceInt cached[2];
    cudaError_t e;
    if (full) {
        auto kern = routing_fused2_bwd_kernel<CL, true>;
        const int rc = fused2_grid(kern, CL, fwd2_smem<CL>(), Nc, cfg, attr, cached[1]);
        if (rc) return rc;
        cfg.stream = st;
        e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agr, g_toqi, g_lo,
                               g_lrfs, g_act, g_ab, Nc, K, I, O, T);
    } else {
        auto kern = routing_fused2_bwd_kernel<CL, false>;
        const int rc = fused2_grid(kern, CL, fwd2_smem<CL>(), Nc, cfg, attr, cached[0]);
        if (rc) return rc;
        cfg.stream = st;
        e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agr, g_toqi, g_lo,
                               g_lrfs, g_act, g_ab, Nc, K, I, O, T);
    }
    return e == cudaSuccess ? launch_status() : (int)e;
}

}  // namespace QEC_VARIANT_NS
}  // namespace qec

using namespace qec;
using namespace qec::QEC_VARIANT_NS;

#ifdef QEC_WIDE
// ------------------------------------------------------------------------------------
// The 512-thread build (routing_fused2_wide.cu): the same kernels with ONE CTA of 16 warps per SM that
// holds 8 192 votes, so a capsule of 32 768 votes needs a cluster of 4 instead of 8.  Clusters of 4 tile the
// GPCs without a remainder (37 capsules in flight on 148 SMs instead of 33 on 132), all sixteen warps of an
// SM sweep the same capsule (the fma pipe is saturated without a second, independent CTA) and the
// reduce -> solve -> publish chain runs unloaded -- but nothing runs in its shadow.  Net effect measured at
// use_wide() below.  Internal entry points, called by the public ones.
// ------------------------------------------------------------------------------------
static int pick_cluster_wide(long long J) {
    for (int cl = 1; cl <= 4; cl *= 2)
        if ((J + cl - 1) / cl <= 2ll * kNpMax) return cl;
    return 0;
}
#define QEC_DISPATCH_CL_WIDE(FN, ...)          \
    switch (CL) {                              \
        case 1: return FN<1>(__VA_ARGS__);     \
        case 2: return FN<2>(__VA_ARGS__);     \
        default: return FN<4>(__VA_ARGS__);    \
    }

extern "C" __attribute__((visibility("hidden"))) int qec_fused2_wide_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                   const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                                   int BP, int K, int I, int O, int T, void* stream) {
    const int CL = pick_cluster_wide((long long)K * I);
    QEC_CHECK_ARG(CL > 0, 11);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL_WIDE(launch_fused2_fwd, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I,
                         O, T, st)
}

extern "C" __attribute__((visibility("hidden"))) int qec_fused2_wide_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                   const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                                   const float* g_agreement, float* g_toqi, unsigned short* g_lo, float* g_lrfs,
                                   float* g_act, float* g_alpha_beta, int BP, int K, int I, int O, int T, void* stream) {
    const int CL = pick_cluster_wide((long long)K * I);
    QEC_CHECK_ARG(CL > 0, 15);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL_WIDE(launch_fused2_bwd, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi,
                         g_lo, g_lrfs, g_act, g_alpha_beta, Nc, K, I, O, T, st)
}

#else  // ---------------- the 256-thread build owns the public entry points ----------------

extern "C" __attribute__((visibility("hidden"))) int qec_fused2_wide_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                   const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                                   int BP, int K, int I, int O, int T, void* stream);
extern "C" __attribute__((visibility("hidden"))) int qec_fused2_wide_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                   const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                                   const float* g_agreement, float* g_toqi, unsigned short* g_lo, float* g_lrfs,
                                   float* g_act, float* g_alpha_beta, int BP, int K, int I, int O, int T, void* stream);

// Geometries that need an 8-CTA cluster of the 256-thread build CAN go to the 512-thread build (clusters of 4).
// Measured at BASELINE config 2: in isolation (scripts/time_pool2.py) forward 0.514 -> 0.526 ms (its chains have
// nothing to hide behind when the CTA is alone on its SM: the gain of 37 instead of 33 capsules in flight is
// spent on them), backward 1.134 -> 1.124 ms plain, 1.188 -> 1.176 ms pre-split; inside the training step
// (bench.py) the backward measures the same either way (1.196 vs 1.197 ms).  So the 256-thread build stays
// the default and the 512-thread one is opt-in:
// QEC_FUSED_WIDE = 0 (never, default) | 1 (both) | f (forward only) | b (backward only).
static std::atomic<int> g_wide_mode{[] {  // 0 never | 1 both | 2 forward only | 3 backward only
    const char* e = getenv("QEC_FUSED_WIDE");
    const char c = (e != nullptr) ? e[0] : '0';
    return c == '1' ? 1 : c == 'f' ? 2 : c == 'b' ? 3 : 0;
}()};
static bool use_wide(long long J, bool backward) {
    const int m = g_wide_mode.load(std::memory_order_relaxed);
    const bool on = m == 1 || (backward ? m == 3 : m == 2);
    return on && pick_cluster((int)J) == 8;
}
// select the build at run time (tests, A/B timing); returns the previous mode
extern "C" int qec_routing_fused_set_wide(int mode) {
    if (mode >= 0 && mode <= 3) return g_wide_mode.exchange(mode, std::memory_order_relaxed);
    return g_wide_mode.load(std::memory_order_relaxed);
}

// Third generation (routing_fused3.cu: two capsules in flight per cluster, dedicated solver warp).  Bit 0: forward,
// bit 1: reserved for a backward (none was built).  Initial value from the environment (QEC_FUSED_PP), default
// QEC_PP_DEFAULT = 0: off (measured slower than the second generation).
extern "C" int qec_routing_fused_pp_supported(int K, int I, int O, int T);
extern "C" int qec_routing_fused_pp_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                        const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                                        int BP, int K, int I, int O, int T, void* stream);
static std::atomic<int> g_pp_mode{[] {
    const char* e = getenv("QEC_FUSED_PP");
    return (e != nullptr && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : QEC_PP_DEFAULT;
}()};
extern "C" int qec_routing_fused_set_pp(int mode) {
    if (mode >= 0 && mode <= 3) return g_pp_mode.exchange(mode, std::memory_order_relaxed);
    return g_pp_mode.load(std::memory_order_relaxed);
}

// 1 if the packed-pair kernels serve this geometry (callers normally just call qec_routing_fused_*,
// which dispatch here when possible)
extern "C" int qec_routing_fused_packed_supported(int K, int I, int O, int T) {
    if (K <= 0 || I <= 0 || O <= 0 || (I & 1)) return 0;
    const long long J = (long long)K * I;
    if (J > 8ll * kJcMax) return 0;
    if (4ll * J * O >= (1ll << 31)) return 0;  // 32-bit element offsets inside one bp block of t_oqi
    return (T >= 1 && T <= kTB && pick_cluster((int)J) > 0) ? 1 : 0;
}

extern "C" int qec_routing_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                     const float* beta, float* pose, float* agreement, float* poses_all,
                                     float* stats, int BP, int K, int I, int O, int T, void* stream) {
    if (!qec_routing_fused_packed_supported(K, I, O, T))
        return qec_routing_fused_fwd_scalar(t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, BP, K, I,
                                            O, T, stream);
    QEC_CHECK_ARG(t_oqi, 1);
    QEC_CHECK_ARG(lrfs, 2);
    QEC_CHECK_ARG(act && alpha && beta, 3);
    QEC_CHECK_ARG(pose && agreement && poses_all && stats, 6);
    QEC_CHECK_ARG(BP > 0, 10);
    if ((g_pp_mode.load(std::memory_order_relaxed) & 1) && qec_routing_fused_pp_supported(K, I, O, T))
        return qec_routing_fused_pp_fwd(t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, BP, K, I, O, T,
                                        stream);
    if (use_wide((long long)K * I, false))
        return qec_fused2_wide_fwd(t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, BP, K, I, O, T, stream);
    const int CL = pick_cluster(K * I);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL(launch_fused2_fwd, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I, O,
                    T, st)
}

#ifdef QEC_PROF
extern "C" int qec_prof_read(long long* dst_host) {
    return (int)cudaMemcpyFromSymbol(dst_host, g_prof, sizeof(long long) * 8 * 64 * 32);
}
#endif

static int fused2_bwd_entry(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                           const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                           const float* g_agreement, float* g_toqi, unsigned short* g_lo, float* g_lrfs, float* g_act,
                           float* g_alpha_beta, int BP, int K, int I, int O, int T, void* stream) {
    QEC_CHECK_ARG(t_oqi, 1);
    QEC_CHECK_ARG(lrfs, 2);
    QEC_CHECK_ARG(act && alpha && beta && poses_all && stats, 3);
    QEC_CHECK_ARG(g_pose && g_agreement, 8);
    QEC_CHECK_ARG(g_toqi && g_alpha_beta, 10);
    QEC_CHECK_ARG(BP > 0, 14);
    if (use_wide((long long)K * I, true))
        return qec_fused2_wide_bwd(t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi, g_lo,
                                   g_lrfs, g_act, g_alpha_beta, BP, K, I, O, T, stream);
    const int CL = pick_cluster(K * I);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    QEC_DISPATCH_CL(launch_fused2_bwd, t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi,
                    g_lo, g_lrfs, g_act, g_alpha_beta, Nc, K, I, O, T, st)
}

extern "C" int qec_routing_fused_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                     const float* beta, const float* poses_all, const float* stats,
                                     const float* g_pose, const float* g_agreement, float* g_toqi, float* g_lrfs,
                                     float* g_act, float* g_alpha_beta, int BP, int K, int I, int O, int T,
                                     void* stream) {
    if (!qec_routing_fused_packed_supported(K, I, O, T))
        return qec_routing_fused_bwd_scalar(t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement,
                                            g_toqi, g_lrfs, g_act, g_alpha_beta, BP, K, I, O, T, stream);
    return fused2_bwd_entry(t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi, nullptr,
                            g_lrfs, g_act, g_alpha_beta, BP, K, I, O, T, stream);
}

// Same as qec_routing_fused_bwd, but dL/dt_oqi is delivered pre-split for tensor-core consumers:
// g_toqi_hi = the gradient rounded to the TF32 grid (FP32 storage), g_toqi_lo = the remainder g - hi
// rounded to BF16 (|g - hi - lo| <= 2^-20 |g|).
// Only for geometries with qec_routing_fused_packed_supported(...) == 1 (argument 15 otherwise).
extern "C" int qec_routing_fused_bwd_split(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                           const float* beta, const float* poses_all, const float* stats,
                                           const float* g_pose, const float* g_agreement, float* g_toqi_hi,
                                           unsigned short* g_toqi_lo, float* g_lrfs, float* g_act, float* g_alpha_beta, int BP,
                                           int K, int I, int O, int T, void* stream) {
    QEC_CHECK_ARG(qec_routing_fused_packed_supported(K, I, O, T), 15);
    QEC_CHECK_ARG(g_toqi_lo, 11);
    return fused2_bwd_entry(t_oqi, lrfs, act, alpha, beta, poses_all, stats, g_pose, g_agreement, g_toqi_hi, g_toqi_lo,
                            g_lrfs, g_act, g_alpha_beta, BP, K, I, O, T, stream);
}
#endif  // QEC_WIDE
