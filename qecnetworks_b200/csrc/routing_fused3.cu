// routing_fused3.cu -- cluster-resident routing FORWARD of the pool2 regime, third generation (opt-in, experimental):
// TWO capsules in flight per cluster and DEDICATED SOLVER WARPS, so that the reduce -> solve -> publish chain of one
// capsule runs in the shadow of the other capsule's sweep instead of stalling the whole CTA.
//
// Why it exists, and why it is not the default (B200, BASELINE config 2, pool2 = 1 280 capsules x 32 768 votes; all
// numbers in profiles/r02_summary.md).  The second generation (routing_fused2.cu) spends 0.505 ms in the forward; the
// same kernel with the chains compiled out (-DQEC_PROBE_NOCHAIN) takes 0.240 ms: more than half of it is the three
// serial chains per capsule (CTA barrier, cross-CTA sum, 4x4 eigen-solve by one warp, broadcast), during which seven
// of eight warps wait, and the sweeps that remain are bound by the shared-memory / LSU pipe (ncu: 53.5 % busy over
// the whole kernel, i.e. saturated while sweeping), not by latency -- which is why software pipelining and the
// Estrin polynomial measured neutral there.  This kernel removes the waiting: its timeline shows the sweepers
// reaching every "pose published" point a few cycles after the pose.  It is nevertheless SLOWER end to end
// (0.553 ms): clusters of 8 whole SMs tile only 120 of the 148 SMs (15 clusters; the second generation's clusters
// of 4 SMs tile 132), a sweep covers 4 pairs per thread instead of 8, so the cross-lane reduction (16 shuffles + ~70
// ALU instructions per warp, on the same LSU pipe that bounds the sweeps) is paid twice as often (0.08 ms: 0.47 ms
// with it compiled out), and two slots hide a chain of ~3 000 cycles (inter-CTA skew + solve) only partly behind a
// sweep of ~2 000.  Its sweeps alone run at 0.299 ms on 120 SMs (= 0.242 ms on 148: the same LSU-bound rate).  Kept
// selectable (qec_routing_fused_set_pp / QEC_FUSED_PP=1), parity-tested, as the measured answer to "hide the chains
// with two capsules per CTA and a solver warp".
//
// Structure.  One CTA per SM: 16 sweeper warps + one solver warp per slot (576 threads, 96 registers).  A cluster of CL
// CTAs owns two capsules at a time ("slots" 0 / 1; consecutive output channels of one patch, so they share LRFs
// and activations); each CTA keeps 4 096 votes of EACH capsule in shared memory (2 x 96 KB: votes 16 B + activation
// 4 B + routing weight 4 B per vote), a sweeper thread owns 4 vote pairs per slot.
//   sweepers:  sweep(slot) -> warp-shuffle reduction -> 16 partials per warp into shared memory -> bar.arrive
//              (they do NOT wait) -> sweep of the OTHER slot -> wait on the mbarrier "pose of this slot published" -> ...
//   solver:    bar.sync on "partials of slot s parked" -> fixed-order sum over the 16 warps -> st.async push into
//              every CTA of the cluster -> mbarrier wait -> fixed-order sum over the CTAs (bit-identical in every
//              CTA, no pose broadcast between CTAs) -> eigen-solve -> pose to shared memory -> mbarrier arrive
// Same arithmetic per vote as routing_fused2.cu (routing_pair.cuh); the reduction tree differs (shuffle tree per
// warp, then warps, then CTAs -- all in fixed order: results are bit-reproducible run to run and agree with the
// second generation to rounding).
//
// Contract: reference models/qec_module.py:126-147 (get_votes after the Linears) + 172-194 (routing loop and
// activation epilogue).  Requires even I (pairs), J = K*I <= 8 * 4096, T <= 10.  Forward only: the backward stays
// on the second generation.
#include <cstdio>
#include <cstdlib>

// Here the sweepers never wait for a chain, so a sweep's own latency is what the kernel time consists of: the
// distance polynomial in Estrin form (dependency depth 3 instead of 8) and software-pipelined shared-memory operands
// (PP_SWP: the next pair's loads are issued before the current pair's store, which the compiler must otherwise keep
// behind it) pay here, while they measured neutral in the second generation, whose time is in its chains.
#ifndef QEC_ESTRIN
#define QEC_ESTRIN 1
#endif
#ifndef PP_SWP
#define PP_SWP 1
#endif

#include "routing_pair.cuh"

namespace qec {
namespace pp {

// Warp roles.  QEC_PP_QUIET = 0: warps 0..15 sweep, warps 16 / 17 solve for slot 0 / 1 (576 threads).
// QEC_PP_QUIET = 1: 16 warps; a warp's scheduler (SM sub-partition) is warp % 4, and sub-partition 3 is kept free of
// sweepers -- warps 3 / 7 solve for slot 0 / 1, warps 11 / 15 idle, the other twelve sweep -- because the solve is a
// serial chain of ~800 dependent instructions that runs several times slower when it has to share its scheduler and
// fma pipe with four sweeping warps.
#ifndef QEC_PP_QUIET
#define QEC_PP_QUIET 0
#endif
constexpr int kSlots = 2;
constexpr int kNp = 2048;              // 2 048 pairs = 4 096 votes per slot and CTA
#if QEC_PP_QUIET
constexpr int kSW = 12;
constexpr int kThreads = 512;
__device__ __forceinline__ bool is_sweeper(int warp) { return (warp & 3) != 3; }
__device__ __forceinline__ int sweeper_warp(int warp) { return warp - (warp >> 2); }
__device__ __forceinline__ int solver_slot(int warp) { return (warp == 3) ? 0 : ((warp == 7) ? 1 : -1); }
#else
constexpr int kSW = 16;
constexpr int kThreads = (kSW + kSlots) * 32;
__device__ __forceinline__ bool is_sweeper(int warp) { return warp < kSW; }
__device__ __forceinline__ int sweeper_warp(int warp) { return warp; }
__device__ __forceinline__ int solver_slot(int warp) { return warp - kSW; }
#endif
constexpr int kSweep = kSW * 32;                   // sweeper threads
constexpr int kBarThreads = kSweep + 32;           // a named barrier joins the sweepers and ONE solver warp
constexpr int kP = (kNp + kSweep - 1) / kSweep;    // vote pairs per sweeper thread and slot (the last one may be partial)
constexpr bool kLastPartial = (kP * kSweep != kNp);
constexpr int kBarPart = 1;            // + slot: partial sums parked (sweepers arrive, solver syncs)
constexpr int kBarPose = 3;            // + slot: pose published (solver arrives, sweepers sync)
constexpr int kClMax = 8;

__device__ __forceinline__ void bar_arrive(int id) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kBarThreads) : "memory");
}
__device__ __forceinline__ void bar_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kBarThreads) : "memory");
}
// "pose of slot s published" is an mbarrier, not a named barrier: a sweeper warp waits for the SOLVER only, never
// for the other sweeper warps (a bar.sync would re-align all sixteen of them at every sweep, and warps in lock-step
// hit the shared-memory pipe and the fma pipe in turns instead of overlapping them: 0.18 ms of the 0.59 ms forward).
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITP_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAITP_%=;\n\t}" ::"r"(mb), "r"(parity)
        : "memory");
}
#ifdef PP_PROBE_NOWAIT  // timing probe only (wrong results): the sweepers never wait for a pose and never signal
#define PP_SWEEPER_SYNC(s_) do { } while (0)
#else
#define PP_SWEEPER_SYNC(s_) do { mbar_wait(smem_u32(&C.pose_mbar[s_]), pose_par[s_] & 1u); ++pose_par[s_]; } while (0)
#endif

// Optional timeline instrumentation (-DQEC_PROF; scripts/timeline_pp.py): lane 0 of sweeper warp 0 and of the two
// solver warps stamp clock64() at the phase boundaries of the first capsule pairs of the first CTAs.
#ifdef QEC_PROF
__device__ long long g_prof3[8 * 48 * 96];
#define PP_EVT(id)                                                                                      \
    do {                                                                                                \
        if ((threadIdx.x & 31) == 0 && blockIdx.x < 8 && prof_n < 48) g_prof3[(blockIdx.x * 48 + prof_n) * 96 + (id)] = clock64(); \
    } while (0)
#else
#define PP_EVT(id) \
    do {           \
    } while (0)
#endif

struct SlotMem {  // one capsule's resident state in this CTA
    float4 A[kNp];  // {w0 w1 x0 x1} per pair
    float4 B[kNp];  // {y0 y1 z0 z1}
    float2 a[kNp];  // activations of the pair
    float2 b[kNp];  // routing weights b^t (forward) / cached 1 - d^0 (backward)
};
struct Ctrl {
    float scr[kSlots][kSW][16];             // per-warp partial sums
    float buf[kSlots][2][kClMax][16];       // received CTA sums, double-buffered by phase
    float out[kSlots][8];                   // published pose (forward) / adjoint vector + 1/S (backward)
    unsigned long long mbar[kSlots][2];
    unsigned long long pose_mbar[kSlots];   // "pose of slot s published": the solver arrives, every sweeper warp waits
};
constexpr size_t kSmemBytes = kSlots * sizeof(SlotMem) + sizeof(Ctrl);

struct PairGeom3 {
    int sid;      // sweeper thread index 0..kSweep-1
    int np;       // live pairs of this CTA (per slot)
    int off[kP];  // k * trow + i of the pair's first vote inside the bp block of t_oqi
    int I;
};

template <bool FULL>
__device__ __forceinline__ void prefetch_pair(const PairGeom3& G, const float* tcap, SlotMem& S, int m, bool go) {
    const int ps = G.sid + m * kSweep;
    if (go && (((FULL && !(kLastPartial && m == kP - 1))) || ps < G.np)) {
        const float* tp = tcap + G.off[m];
        float* a = reinterpret_cast<float*>(S.A + ps);
        float* b = reinterpret_cast<float*>(S.B + ps);
        cp_async8(a, tp);
        cp_async8(a + 2, tp + G.I);
        cp_async8(b, tp + 2 * G.I);
        cp_async8(b + 2, tp + 3 * G.I);
    }
    cp_async_commit();  // always: the wait counts below assume one group per (slot, pair)
}

// at most n of the most recently committed cp.async groups may still be pending (n < 12, folded at compile time)
__device__ __forceinline__ void wait_pending(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        case 7: cp_async_wait<7>(); break;
        case 8: cp_async_wait<8>(); break;
        case 9: cp_async_wait<9>(); break;
        case 10: cp_async_wait<10>(); break;
        default: cp_async_wait<11>(); break;
    }
}

// sweepers: park this warp's sums of N <= 16 values and signal the solver; no waiting
template <int N>
__device__ __forceinline__ void post_partials(Ctrl& C, int slot, const float (&a)[N]) {
#ifdef PP_PROBE_NOWAIT
    return;
#endif
    const int lane = threadIdx.x & 31, warp = sweeper_warp(threadIdx.x >> 5);
#ifdef PP_PROBE_NOPOST  // timing probe only (wrong results): no shuffle reduction, only the signal
    if ((lane & 1) == 0) C.scr[slot][warp][idx16(lane)] = a[0];
    __syncwarp();
    bar_arrive(kBarPart + slot);
    return;
#endif
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = (i < N) ? a[i] : 0.f;
    const float mine = warp_reduce_scatter16(w, lane);  // lanes 2i, 2i+1 hold the warp total of value i
    if ((lane & 1) == 0) C.scr[slot][warp][idx16(lane)] = mine;
    __syncwarp();
    bar_arrive(kBarPart + slot);
}

// solver warp: cluster-wide totals of the N values the sweepers parked for `slot`; valid in all lanes afterwards
template <int CL>
__device__ __forceinline__ void solver_totals(Ctrl& C, int slot, int N, unsigned rank, unsigned& phase, float (&a)[16],
                                              int evt = 0, int prof_n = 48) {
    (void)evt; (void)prof_n;
    const int lane = threadIdx.x & 31;
    const unsigned par = phase & 1u;
    const unsigned mb = smem_u32(&C.mbar[slot][par]);
    if (lane == 0)  // arm this phase: N floats from each of the CL CTAs
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(N * CL * 4) : "memory");
    bar_sync(kBarPart + slot);  // every sweeper warp has parked its sums
    PP_EVT(evt);
    float s = 0.f;
    if (lane < 16) {
#pragma unroll
        for (int w = 0; w < kSW; ++w) s += C.scr[slot][w][lane];  // fixed order
    }
    if (lane < N) {
        const unsigned dst = smem_u32(&C.buf[slot][par][rank][lane]);
#pragma unroll
        for (int r = 0; r < CL; ++r) st_async_f32(mapa_u32(dst, r), s, mapa_u32(mb, r));
    }
    PP_EVT(evt + 1);
    const unsigned parity = (phase >> 1) & 1u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(mb), "r"(parity)
        : "memory");
    PP_EVT(evt + 2);
    float t = 0.f;
    if (lane < 16) {
#pragma unroll
        for (int r = 0; r < CL; ++r) t += C.buf[slot][par][r][lane];  // fixed order: identical bits in every CTA
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = __shfl_sync(0xffffffffu, t, i);
    ++phase;
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int CL, bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
    routing_pp_fwd_kernel(const float* __restrict__ t_oqi, const float* __restrict__ lrfs, const float* __restrict__ act,
                          const float* __restrict__ alpha_p, const float* __restrict__ beta_p, float* __restrict__ pose,
                          float* __restrict__ agreement, float* __restrict__ poses_all, float* __restrict__ stats,
                          long long Nc, int K, int I, int O, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SlotMem* S = reinterpret_cast<SlotMem*>(smem_raw);
    Ctrl& C = *reinterpret_cast<Ctrl*>(smem_raw + kSlots * sizeof(SlotMem));
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
#pragma unroll
            for (int p = 0; p < 2; ++p)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&C.mbar[s][p])) : "memory");
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&C.pose_mbar[s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // every CTA of the cluster is resident and its mbarriers initialised

    const long long cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    const long long npair = (Nc + 1) / 2;
    const int J = K * I;
    const int Jc = 2 * ((J / 2 + CL - 1) / CL);  // even slice per CTA
    const int j_begin = min(J, (int)rank * Jc);
    const int nloc = min(J, j_begin + Jc) - j_begin;
    const size_t trow = (size_t)4 * I * O;
    const float alpha = __ldg(alpha_p), beta = __ldg(beta_p);
    auto tcap_of = [&](long long c) {
        const long long bp = c / O;
        const int o = (int)(c - bp * O);
        return t_oqi + (size_t)bp * K * trow + (size_t)o * 4 * I;
    };

    const int warp_id = threadIdx.x >> 5;
    if (!is_sweeper(warp_id)) {
        // =============================== solver warps: one per slot ===============================
        const int lane = threadIdx.x & 31;
#ifdef PP_PROBE_NOWAIT
        const int s = -1;
#else
        const int s = solver_slot(warp_id);
#endif
        if (s >= 0 && s < kSlots) {
        const bool writer = (rank == 0) && (lane == 0);
        unsigned phase = 0u;
        bool pend = false, valid = false;
        long long pend_c = 0;
        Quat pend_p = qmake(0.f, 0.f, 0.f, 0.f);
        auto finish = [&](long long c, Quat pf, float es, float en) {  // reference qec_module.py:188-193
            const float sm = (en > 0.f) ? es / en : 0.f;
            stq(pose + (size_t)c * 4, pf);
            agreement[c] = (en > 0.f) ? sigmoidf(fmaf(alpha, sm, beta - 1.f)) : 0.f;
            stats[2 * c] = sm;
            stats[2 * c + 1] = en;
        };
#ifdef QEC_PROF
        int prof_n = 0;
#endif
        for (long long q = cid; q < npair; q += ncl) {
            const long long c = 2 * q + s;
            if (c >= Nc) break;
            for (int t = 0; t < T; ++t) {
                float a[16];
#ifdef QEC_PROF
                solver_totals<CL>(C, s, (t == 0) ? 15 : 12, rank, phase, a, 50 + s * 16 + t * 4, prof_n);
#else
                solver_totals<CL>(C, s, (t == 0) ? 15 : 12, rank, phase, a);
#endif
                if (t == 0) {
                    if (pend && writer) finish(pend_c, pend_p, a[13], a[14]);
                    valid = a[12] != 0.f;  // reference qec_module.py:72-73
                }
#if defined(QEC_PP_PROBE_NOSOLVE)  // timing probe only (wrong results): the chain without the eigen-solve
                const float rn_ = frsqrt(fmaf(a[0], a[0], fmaf(a[1], a[1], fmaf(a[2], a[2], a[3] * a[3]))) + 1e-30f);
                const Quat p = qmake(fabsf(a[0]) * rn_, a[1] * rn_, a[2] * rn_, a[3] * rn_);
#else
                const Quat p = solve_pose2(a, valid);
#endif
                if (lane == 0) {
                    C.out[s][0] = p.w; C.out[s][1] = p.x; C.out[s][2] = p.y; C.out[s][3] = p.z;
                    if (rank == 0) stq(poses_all + ((size_t)t * Nc + c) * 4, p);
                }
                if (t == T - 1) {
                    pend = true;
                    pend_c = c;
                    pend_p = p;
                }
                __syncwarp();
                PP_EVT(50 + s * 16 + t * 4 + 3);
                if (lane == 0)  // release: the pose stores above are visible to whoever observes the phase flip
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&C.pose_mbar[s])) : "memory");
            }
#ifdef QEC_PROF
            ++prof_n;
#endif
        }
        if (pend) {  // the last capsule of this slot: its epilogue sums arrive alone
            float a[16];
            solver_totals<CL>(C, s, 2, rank, phase, a);
            if (writer) finish(pend_c, pend_p, a[0], a[1]);
        }
        }  // (idle warps of the quiet sub-partition fall through to the final cluster barrier)
    } else {
        // =============================== sweeper warps ===============================
        PairGeom3 G;
        G.sid = sweeper_warp(warp_id) * 32 + (threadIdx.x & 31);
        G.np = nloc / 2;
        G.I = I;
#pragma unroll
        for (int m = 0; m < kP; ++m) {
            const int j = j_begin + 2 * (G.sid + m * kSweep);
            const int k = j / I;
            G.off[m] = (int)(k * trow) + (j - k * I);
        }
        auto live = [&](int m) { return (FULL && !(kLastPartial && m == kP - 1)) || (G.sid + m * kSweep < G.np); };
        if (cid < npair) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const long long c = 2 * cid + s;
#pragma unroll
                for (int m = 0; m < kP; ++m) prefetch_pair<FULL>(G, tcap_of(c < Nc ? c : 0), S[s], m, c < Nc);
            }
        }
        bool pend[kSlots] = {false, false};
        float e0[kSlots] = {0.f, 0.f}, e1[kSlots] = {0.f, 0.f};
        unsigned pose_par[kSlots] = {0u, 0u};

#ifdef QEC_PROF
        int prof_n = (threadIdx.x < 32) ? 0 : 48;
#endif
        for (long long q = cid; q < npair; q += ncl) {
            PP_EVT(0);
            const long long bp = (2 * q) / O;  // (a pair never straddles two patches when O is even; else per slot)
            // ---- sweep 0 of both slots: votes from the prefetched transforms, M^0 = sum a v v^T, b^0 = a
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const long long c = 2 * q + s;
                if (c >= Nc) continue;
                const long long bps = (O & 1) ? c / O : bp;
                const float* lrf_c = lrfs + ((size_t)bps * J + j_begin) * 4;
                const float* act_c = act + (size_t)bps * J + j_begin;
                SlotMem& M = S[s];
                f2 m2[10], ssum = bc(0.f), comp = bc(0.f);
                float sabs = 0.f;
#pragma unroll
                for (int i = 0; i < 10; ++i) m2[i] = bc(0.f);
                Quat2 lr[2];
                f2 av[2];
                auto ldl = [&](int m, Quat2& lrf, f2& a) {  // LRFs / activations of the pair (L2: shared by all O)
                    const int ps = G.sid + m * kSweep;
                    if (live(m)) {
                        const Quat l0 = ldq_nc(lrf_c + (size_t)ps * 8), l1 = ldq_nc(lrf_c + (size_t)ps * 8 + 4);
                        lrf = pk(l0, l1);
                        a = ldg_f2(act_c + 2 * ps);
                    }
                };
                ldl(0, lr[0], av[0]);
#if PP_SWP
                Quat2 tr[2];
                wait_pending((s == 0 ? kP : 0) + (kP - 1));
                if (live(0)) tr[0] = lds_pair(M.A, M.B, G.sid);
#pragma unroll
                for (int m = 0; m < kP; ++m) {
                    if (m + 1 < kP) {
                        ldl(m + 1, lr[(m + 1) & 1], av[(m + 1) & 1]);
                        wait_pending((s == 0 ? kP : 0) + (kP - 2 - m));
                        if (live(m + 1)) tr[(m + 1) & 1] = lds_pair(M.A, M.B, G.sid + (m + 1) * kSweep);
                    }
                    if (live(m)) {
                        const int ps = G.sid + m * kSweep;
                        const f2 a = av[m & 1];
                        const Quat2 v = make_vote2<false>(lr[m & 1], tr[m & 1]).vote;
                        sts_pair(M.A, M.B, ps, v);
                        sts_f2(M.a + ps, a);
                        sts_f2(M.b + ps, a);
                        sabs = (sabs + fabsf(lo(a))) + fabsf(hi(a));
                        ssum = add2(ssum, a);
                        acc_rank1_2(m2, a, v);
                        comp = add2(comp, add2(add2(v.w, v.x), add2(v.y, v.z)));
                    }
                }
#else
#pragma unroll
                for (int m = 0; m < kP; ++m) {
                    if (m + 1 < kP) ldl(m + 1, lr[(m + 1) & 1], av[(m + 1) & 1]);
                    // groups are committed in (slot, pair) order, 4 per slot: slot 0 waits with slot 1's still pending
                    wait_pending((s == 0 ? kP : 0) + (kP - 1 - m));
                    if (live(m)) {
                        const int ps = G.sid + m * kSweep;
                        const Quat2 traw = lds_pair(M.A, M.B, ps);
                        const f2 a = av[m & 1];
                        const Quat2 v = make_vote2<false>(lr[m & 1], traw).vote;
                        sts_pair(M.A, M.B, ps, v);
                        sts_f2(M.a + ps, a);
                        sts_f2(M.b + ps, a);
                        sabs = (sabs + fabsf(lo(a))) + fabsf(hi(a));
                        ssum = add2(ssum, a);
                        acc_rank1_2(m2, a, v);
                        comp = add2(comp, add2(add2(v.w, v.x), add2(v.y, v.z)));
                    }
                }
#endif
                float acc[15];
#pragma unroll
                for (int i = 0; i < 10; ++i) acc[i] = hsum(m2[i]);
                acc[10] = sabs;
                acc[11] = hsum(ssum);
                acc[12] = hsum(comp);
                acc[13] = e0[s];  // the previous capsule's epilogue sums ride along
                acc[14] = e1[s];
                PP_EVT(10 + s * 4 + 1);
                post_partials(C, s, acc);
                PP_EVT(10 + s * 4 + 2);
            }
            // ---- sweeps 1..T-1: b <- b a (1 - d(pose, v)), M^t = sum b v v^T, all from shared memory
            for (int t = 1; t < T; ++t) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    if (2 * q + s >= Nc) continue;
                    SlotMem& M = S[s];
                    PP_EVT(10 + t * 8 + s * 4 + 3);
                    PP_SWEEPER_SYNC(s);  // pose^{t-1} of this slot is published
                    PP_EVT(10 + t * 8 + s * 4 + 0);
                    const Quat p = qmake(C.out[s][0], C.out[s][1], C.out[s][2], C.out[s][3]);
                    f2 m2[10], ssum = bc(0.f);
                    float sabs = 0.f;
#pragma unroll
                    for (int i = 0; i < 10; ++i) m2[i] = bc(0.f);
#if PP_SWP
                    Quat2 vn;
                    f2 an = bc(0.f), bn = bc(0.f);
                    auto lds_slot = [&](int m) {
                        const int ps = G.sid + m * kSweep;
                        if (live(m)) {
                            vn = lds_pair(M.A, M.B, ps);
                            an = lds_f2(M.a + ps);
                            bn = lds_f2(M.b + ps);
                        }
                    };
                    lds_slot(0);
#pragma unroll
                    for (int m = 0; m < kP; ++m) {
                        const int ps = G.sid + m * kSweep;
                        const Quat2 v = vn;
                        const f2 a = an, b0 = bn;
                        if (m + 1 < kP) lds_slot(m + 1);  // before this pair's store
                        if (live(m)) {
                            const f2 b = mul2(b0, mul2(a, one_minus_dist2(qdot2(p, v))));
                            sts_f2(M.b + ps, b);
                            sabs = (sabs + fabsf(lo(b))) + fabsf(hi(b));
                            ssum = add2(ssum, b);
                            acc_rank1_2(m2, b, v);
                        }
                    }
#else
#pragma unroll
                    for (int m = 0; m < kP; ++m) {
                        const int ps = G.sid + m * kSweep;
                        if (live(m)) {
                            const Quat2 v = lds_pair(M.A, M.B, ps);
                            const f2 a = lds_f2(M.a + ps), b0 = lds_f2(M.b + ps);
                            const f2 b = mul2(b0, mul2(a, one_minus_dist2(qdot2(p, v))));
                            sts_f2(M.b + ps, b);
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
                    PP_EVT(10 + t * 8 + s * 4 + 1);
                    post_partials(C, s, a2);
                    PP_EVT(10 + t * 8 + s * 4 + 2);
                }
            }
            // ---- epilogue sweeps (reference qec_module.py:188-193); each pair slot is refilled with the next capsule's
            // transforms as soon as its vote has been read
            const long long qn = q + ncl;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const long long c = 2 * q + s, cn = 2 * qn + s;
                const bool more = qn < npair;            // (commit groups for both slots whenever a next pair exists)
                const bool fill = more && cn < Nc;
                const float* tn = fill ? tcap_of(cn) : t_oqi;
                SlotMem& M = S[s];
                if (c < Nc) {
                    PP_EVT(40 + s * 3 + 2);
                    PP_SWEEPER_SYNC(s);  // pose^{T-1}
                    PP_EVT(40 + s * 3);
                    const Quat p = qmake(C.out[s][0], C.out[s][1], C.out[s][2], C.out[s][3]);
                    f2 es = bc(0.f);
                    float en = 0.f;
#if PP_SWP
                    Quat2 vn;
                    f2 bn = bc(0.f);
                    auto lds_slot = [&](int m) {
                        const int ps = G.sid + m * kSweep;
                        if (live(m)) {
                            vn = lds_pair(M.A, M.B, ps);
                            bn = lds_f2(M.b + ps);
                        }
                    };
                    lds_slot(0);
#pragma unroll
                    for (int m = 0; m < kP; ++m) {
                        const Quat2 v = vn;
                        const f2 b = bn;
                        if (m + 1 < kP) lds_slot(m + 1);  // ahead of the cp.async below (a compiler barrier)
                        if (more) prefetch_pair<FULL>(G, tn, M, m, fill);  // pair m is in registers: refill its slot
                        if (live(m)) {
                            es = fma2(b, one_minus_dist2(qdot2(p, v)), es);
                            en += ((lo(b) != 0.f) ? 1.f : 0.f) + ((hi(b) != 0.f) ? 1.f : 0.f);
                        }
                    }
#else
#pragma unroll
                    for (int m = 0; m < kP; ++m) {
                        const int ps = G.sid + m * kSweep;
                        if (live(m)) {
                            const Quat2 v = lds_pair(M.A, M.B, ps);
                            const f2 b = lds_f2(M.b + ps);
                            es = fma2(b, one_minus_dist2(qdot2(p, v)), es);
                            en += ((lo(b) != 0.f) ? 1.f : 0.f) + ((hi(b) != 0.f) ? 1.f : 0.f);
                        }
                        if (more) prefetch_pair<FULL>(G, tn, M, m, fill);
                    }
#endif
                    e0[s] = hsum(es);
                    e1[s] = en;
                    pend[s] = true;
                    PP_EVT(40 + s * 3 + 1);
                } else if (more) {
#pragma unroll
                    for (int m = 0; m < kP; ++m) prefetch_pair<FULL>(G, tn, M, m, fill);
                }
            }
#ifdef QEC_PROF
            ++prof_n;
#endif
        }
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            if (pend[s]) {
                float e[2] = {e0[s], e1[s]};
                post_partials(C, s, e);
            }
        }
    }
    cluster_sync_all();  // no CTA may exit while a peer can still push into its receive buffers
}

}  // namespace pp
}  // namespace qec

using namespace qec;
using namespace qec::pp;

#ifdef QEC_PROF
extern "C" int qec_prof3_read(long long* dst_host) {
    return (int)cudaMemcpyFromSymbol(dst_host, g_prof3, sizeof(long long) * 8 * 48 * 96);
}
#endif

static int pp_cluster(long long J) {
    for (int cl = 1; cl <= kClMax; cl *= 2)
        if ((J + cl - 1) / cl <= 2ll * kNp) return cl;
    return 0;
}

// 1 if the two-capsules-in-flight kernels serve this geometry (forward; the backward additionally needs T <= 3)
extern "C" int qec_routing_fused_pp_supported(int K, int I, int O, int T) {
    if (K <= 0 || I <= 0 || O <= 0 || (I & 1) || T < 1 || T > 10) return 0;
    const long long J = (long long)K * I;
    if (4ll * J * O >= (1ll << 31)) return 0;  // 32-bit element offsets inside one bp block of t_oqi
    return pp_cluster(J) > 0 ? 1 : 0;
}

template <int CL, bool FULL>
static int launch_pp_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha, const float* beta,
                         float* pose, float* agreement, float* poses_all, float* stats, long long Nc, int K, int I, int O,
                         int T, cudaStream_t st) {
    auto kern = routing_pp_fwd_kernel<CL, FULL>;
    static PerDeviceInt cached;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
    cfg.attrs = attr;
    std::atomic<int>* slot = cached.slot();
    int clusters = slot ? slot->load(std::memory_order_relaxed) : -1;
    if (clusters < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        cfg.gridDim = dim3(CL * 2 * kNumSMs);
        e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
        if (e != cudaSuccess) return (int)e;
        if (clusters < 1) return (int)cudaErrorLaunchOutOfResources;
        if (getenv("QEC_PP_VERBOSE")) fprintf(stderr, "routing_pp_fwd<%d>: %d clusters co-resident (%d SMs)\n", CL, clusters, clusters * CL);
        if (slot) slot->store(clusters, std::memory_order_relaxed);
    }
    const long long npair = (Nc + 1) / 2;
    if (clusters > npair) clusters = (int)npair;
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    cfg.stream = st;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats,
                                             Nc, K, I, O, T);
    return e == cudaSuccess ? launch_status() : (int)e;
}

// same contract as qec_routing_fused_fwd (include/qec_b200.h); argument 11 if the geometry is not served
extern "C" int qec_routing_fused_pp_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                        const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                                        int BP, int K, int I, int O, int T, void* stream) {
    QEC_CHECK_ARG(t_oqi, 1);
    QEC_CHECK_ARG(lrfs, 2);
    QEC_CHECK_ARG(act && alpha && beta, 3);
    QEC_CHECK_ARG(pose && agreement && poses_all && stats, 6);
    QEC_CHECK_ARG(BP > 0, 10);
    QEC_CHECK_ARG(qec_routing_fused_pp_supported(K, I, O, T), 11);
    const long long J = (long long)K * I;
    const int CL = pp_cluster(J);
    const bool full = (J == (long long)CL * 2 * kNp);
    const long long Nc = (long long)BP * O;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define QEC_PP_CASE(CL_)                                                                                                  \
    case CL_:                                                                                                             \
        return full ? launch_pp_fwd<CL_, true>(t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I, O, T, st) \
                    : launch_pp_fwd<CL_, false>(t_oqi, lrfs, act, alpha, beta, pose, agreement, poses_all, stats, Nc, K, I, O, T, st);
    switch (CL) {
        QEC_PP_CASE(1)
        QEC_PP_CASE(2)
        QEC_PP_CASE(4)
        default:
            QEC_PP_CASE(8)
    }
#undef QEC_PP_CASE
}
