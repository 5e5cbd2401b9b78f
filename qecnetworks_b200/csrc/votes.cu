// votes.cu -- vote generation and its adjoint (reference models/qec_module.py:126-147,
// models/quat_ops.py:30-48).
//
//   t[b,p,k,o,i] = fix(normalise(t_raw[(b,p,k), q*I*O + i*O + o], q = 0..3))
//   v[b,p,o,k*I+i] = fix(lrf[b,p,k,i] (x) t[b,p,k,o,i])
//
// The Linear emits features with `o` innermost while the routing wants the votes
// of one capsule (b,p,o) contiguous over j = k*I+i, so this is a transpose fused
// with the quaternion math.  One CTA owns a (patch, j-tile, o-tile): it stages
// the raw transform components through shared memory with rows read coalesced
// along `o`, then computes with lanes running along `j` so that every vote is
// one float4 store into a contiguous run of the capsule's row.
#include "vote_math.cuh"

namespace qec {

constexpr int kVT = 256;   // threads per CTA
constexpr int kJT = 32;    // votes (j) per tile
constexpr int kOT = 64;    // output capsules (o) per tile
constexpr int kRS = 4 * kOT + 1;  // odd row stride: conflict-free reads along j

// smem[jl][q][ol] <- t_raw tile (coalesced along o)
__device__ __forceinline__ void stage_traw(float* sm, const float* __restrict__ t_raw, long long bp, int K, int I,
                                           int O, int j0, int jn, int o0, int on) {
    const int total = jn * 4 * on;
    for (int e = threadIdx.x; e < total; e += kVT) {
        const int ol = e % on;
        const int r = e / on;
        const int q = r & 3;
        const int jl = r >> 2;
        const int j = j0 + jl;
        const int k = j / I, i = j - k * I;
        const size_t row = (size_t)(bp * K + k) * 4 * I * O;
        sm[jl * kRS + q * kOT + ol] = __ldg(t_raw + row + ((size_t)q * I + i) * O + o0 + ol);
    }
}

__global__ void __launch_bounds__(kVT)
    votes_fwd_kernel(const float* __restrict__ lrfs, const float* __restrict__ t_raw, float* __restrict__ votes,
                     int K, int I, int O, int ntj, int nto) {
    __shared__ float sm[kJT * kRS];
    const int J = K * I;
    const int to = blockIdx.x % nto;
    const int tj = (blockIdx.x / nto) % ntj;
    const long long bp = blockIdx.x / (nto * ntj);
    const int j0 = tj * kJT, o0 = to * kOT;
    const int jn = min(kJT, J - j0), on = min(kOT, O - o0);
    stage_traw(sm, t_raw, bp, K, I, O, j0, jn, o0, on);
    __syncthreads();
    const int total = jn * on;
    for (int e = threadIdx.x; e < total; e += kVT) {
        const int jl = e % jn, ol = e / jn;
        const float* s = sm + jl * kRS + ol;
        const Quat traw = qmake(s[0], s[kOT], s[2 * kOT], s[3 * kOT]);
        const Quat lrf = ldq_nc(lrfs + ((size_t)bp * J + j0 + jl) * 4);  // [bp][k][i] == [bp][j]
        const VoteParts vp = make_vote(lrf, traw);
        stq_stream(votes + (((size_t)bp * O + o0 + ol) * J + j0 + jl) * 4, vp.vote);
    }
}

__global__ void __launch_bounds__(kVT)
    votes_bwd_kernel(const float* __restrict__ lrfs, const float* __restrict__ t_raw,
                     const float* __restrict__ g_votes, float* __restrict__ g_traw, float* g_lrfs, int K, int I,
                     int O, int ntj, int nto) {
    __shared__ float sm[kJT * kRS];
    __shared__ float sgl[kJT * 4];
    const int J = K * I;
    const int to = blockIdx.x % nto;
    const int tj = (blockIdx.x / nto) % ntj;
    const long long bp = blockIdx.x / (nto * ntj);
    const int j0 = tj * kJT, o0 = to * kOT;
    const int jn = min(kJT, J - j0), on = min(kOT, O - o0);
    stage_traw(sm, t_raw, bp, K, I, O, j0, jn, o0, on);
    if (threadIdx.x < kJT * 4) sgl[threadIdx.x] = 0.f;
    __syncthreads();
    const int total = jn * on;
    for (int e = threadIdx.x; e < total; e += kVT) {
        const int jl = e % jn, ol = e / jn;
        float* s = sm + jl * kRS + ol;
        const Quat traw = qmake(s[0], s[kOT], s[2 * kOT], s[3 * kOT]);
        const Quat lrf = ldq_nc(lrfs + ((size_t)bp * J + j0 + jl) * 4);
        const VoteParts vp = make_vote(lrf, traw);
        const Quat g = ldq_stream(g_votes + (((size_t)bp * O + o0 + ol) * J + j0 + jl) * 4);
        Quat gt, gl;
        vote_bwd(lrf, vp, g, gt, gl);
        s[0] = gt.w; s[kOT] = gt.x; s[2 * kOT] = gt.y; s[3 * kOT] = gt.z;  // same slot this thread read
        if (g_lrfs != nullptr) {
            atomicAdd(sgl + jl * 4 + 0, gl.w);
            atomicAdd(sgl + jl * 4 + 1, gl.x);
            atomicAdd(sgl + jl * 4 + 2, gl.y);
            atomicAdd(sgl + jl * 4 + 3, gl.z);
        }
    }
    __syncthreads();
    const int total4 = jn * 4 * on;
    for (int e = threadIdx.x; e < total4; e += kVT) {
        const int ol = e % on;
        const int r = e / on;
        const int q = r & 3;
        const int jl = r >> 2;
        const int j = j0 + jl;
        const int k = j / I, i = j - k * I;
        const size_t row = (size_t)(bp * K + k) * 4 * I * O;
        g_traw[row + ((size_t)q * I + i) * O + o0 + ol] = sm[jl * kRS + q * kOT + ol];
    }
    if (g_lrfs != nullptr && threadIdx.x < jn * 4) {
        float* dst = g_lrfs + ((size_t)bp * J + j0) * 4 + threadIdx.x;
        if (nto == 1)
            *dst = sgl[threadIdx.x];
        else
            atomicAdd(dst, sgl[threadIdx.x]);
    }
}

}  // namespace qec

using namespace qec;

extern "C" int qec_votes_fwd(const float* lrfs, const float* t_raw, float* votes, int BP, int K, int I, int O,
                             void* stream) {
    QEC_CHECK_ARG(lrfs, 1);
    QEC_CHECK_ARG(t_raw, 2);
    QEC_CHECK_ARG(votes, 3);
    QEC_CHECK_ARG(BP > 0, 4);
    QEC_CHECK_ARG(K > 0, 5);
    QEC_CHECK_ARG(I > 0, 6);
    QEC_CHECK_ARG(O > 0, 7);
    const int J = K * I;
    const int ntj = (J + kJT - 1) / kJT, nto = (O + kOT - 1) / kOT;
    const long long blocks = (long long)BP * ntj * nto;
    QEC_CHECK_ARG(blocks < (1ll << 31), 4);
    votes_fwd_kernel<<<(unsigned)blocks, kVT, 0, static_cast<cudaStream_t>(stream)>>>(lrfs, t_raw, votes, K, I, O,
                                                                                      ntj, nto);
    return launch_status();
}

// g_lrfs (nullable) must be zero-initialised by the caller when O > 64 (it is
// accumulated across o-tiles); otherwise it is overwritten.
extern "C" int qec_votes_bwd(const float* lrfs, const float* t_raw, const float* g_votes, float* g_traw,
                             float* g_lrfs, int BP, int K, int I, int O, void* stream) {
    QEC_CHECK_ARG(lrfs, 1);
    QEC_CHECK_ARG(t_raw, 2);
    QEC_CHECK_ARG(g_votes, 3);
    QEC_CHECK_ARG(g_traw, 4);
    QEC_CHECK_ARG(BP > 0, 6);
    QEC_CHECK_ARG(K > 0, 7);
    QEC_CHECK_ARG(I > 0, 8);
    QEC_CHECK_ARG(O > 0, 9);
    const int J = K * I;
    const int ntj = (J + kJT - 1) / kJT, nto = (O + kOT - 1) / kOT;
    const long long blocks = (long long)BP * ntj * nto;
    QEC_CHECK_ARG(blocks < (1ll << 31), 6);
    votes_bwd_kernel<<<(unsigned)blocks, kVT, 0, static_cast<cudaStream_t>(stream)>>>(lrfs, t_raw, g_votes, g_traw,
                                                                                      g_lrfs, K, I, O, ntj, nto);
    return launch_status();
}
