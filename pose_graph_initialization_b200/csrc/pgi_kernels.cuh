// pgi_kernels.cuh — batched sm_100a kernels of the hypothesis-verification path.
//
//   K0 build_correspondences   a1  createCorrespondenceMatrix          pose_graph_builder.h:864-938
//   K1 score_hypotheses        a2-a6  test + getInliers fused          graph_traversal.h:136-168, :194-233
//   K2 fivept_first_solution   a7  cv::findEssentialMat(RANSAC, thr=inf)  pose_graph_builder.h:1013-1020
//   K4 fallback_solve          a8  minimal five-point solves of the robust loop   pose_graph_builder.h:1037-1044
//   K5 fallback_score          a8  scoring / LO refit / termination of the robust loop
//   K3 decompose_vote          a9-a11 + verdict packing (K8)           pose_utils.h:144-252, pose_graph_builder.h:1069-1075
//
// Data layout in HBM: correspondences are one contiguous FP64 array of [x1 y1 x2 y2] rows (32 B per
// correspondence, 16-B aligned => one LDG.128 pair per row), pair k owning rows corr_offset[k]..corr_offset[k+1].
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/pgi.h"
#include "pgi_math.cuh"

namespace pgi {

constexpr int kCtaThreads = 256;      // == oracle fb::kStride (fixed reduction order)
constexpr int kFbChunk = 125;         // fallback iterations solved per K4 launch
constexpr int kLoRounds = 4;
constexpr int kLoMinInliers = 9;
constexpr int kDkMaxSweeps = 200;
constexpr double kDkTolSq = 1e-26;

enum : uint32_t {
    ST_HAVE_HYP = 1u,       // pair had >= 1 hypothesis in this wave
    ST_TEST_PASSED = 2u,    // last hypothesis passed InTraversalPoseTester::test
    ST_NEED_5PT = 4u,       // path branch continues into the five-point solve
    ST_PATH_OK = 8u,        // path branch succeeded (inlierNumber >= minInliers)
    ST_NEED_FB = 16u,       // fallback must run
    ST_FB_ACTIVE = 32u,     // fallback loop still iterating
    ST_FB_RAN = 64u,
    ST_HAVE_E = 128u,       // state.E is valid and must be decomposed
    ST_5PT_NOMODEL = 256u
};

// Per-wave-slot working state (device only).
struct SlotState {
    double E[9];
    double bestCost;
    double bestE[9];
    uint32_t flags;
    uint32_t testCount;
    uint32_t pathInliers;
    uint32_t inlierCount;
    int32_t bestInl;
    int32_t maxIters;
    int32_t it;
    uint32_t models;
    uint32_t tableIdx;
    uint32_t pad;
};

struct WaveArgs {
    // registered pairs
    const double *corr;        // sum N x 4
    const uint64_t *offset;    // n_pairs + 1
    const double *thr;         // n_pairs
    const uint32_t *pairTable; // n_pairs: sampler/iters table index of the pair's N
    // sampler tables
    const uint32_t *samplerTab;   // tables x maxIters x 5
    const uint16_t *itersTab;     // concatenated (N+1) tables
    const uint64_t *itersTabOff;  // tables
    // wave
    uint32_t n;
    const uint32_t *pairId;     // n
    const uint32_t *hypOffset;  // n + 1
    const double *hyp;          // sum H x 7
    SlotState *state;           // n
    uint32_t *bits;             // n x bitsStride words: inlier bit mask of the path branch
    uint32_t bitsStride;
    uint8_t *masks;             // optional: concatenated byte masks
    const uint64_t *maskOffset; // n + 1 (rows, wave-local)
    pgi_verdict *verdicts;      // n
    uint32_t flags;
    uint32_t minInliers;
    uint32_t testMinInliers;
    uint32_t fbMaxIters;
    double thrMultiplier;
    double thrOverride;  // > 0: tester threshold given explicitly (pgi_test_pose)
    // fallback scratch
    double *fbSols;     // n x kFbChunk x 90
    uint8_t *fbCounts;  // n x kFbChunk
    unsigned long long *counters;  // [0] corr evals, [1] fallback pairs, [2] fallback models
};

// ---------------------------------------------------------------------------------------------
// K0: createCorrespondenceMatrix on the device from the compact layout.
// One thread per correspondence; keypoint gathers hit L2 (<= 64 KB per view).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k0_build_correspondences(
    const double *__restrict__ focal, const double *__restrict__ sizeWH, const uint64_t *__restrict__ kpOffset,
    const float2 *__restrict__ kp, uint64_t nPairs, const uint2 *__restrict__ pairViews,
    const uint64_t *__restrict__ mOffset, const uint2 *__restrict__ matches, double thrPx, double4 *__restrict__ corr,
    double *__restrict__ thrNorm)
{
    const uint64_t p = blockIdx.y + (uint64_t)blockIdx.z * gridDim.y;
    if (p >= nPairs) return;
    const uint2 v = pairViews[p];
    const double f = focal[v.x];  // K = [f 0 w/2; 0 f h/2; 0 0 1]  pose_graph_builder.h:284-286
    const double cx = sizeWH[2 * v.x] / 2.0, cy = sizeWH[2 * v.x + 1] / 2.0;
    const uint64_t m0 = mOffset[p], m1 = mOffset[p + 1];
    const float2 *ks = kp + kpOffset[v.x];
    const float2 *kd = kp + kpOffset[v.y];
    for (uint64_t i = m0 + blockIdx.x * blockDim.x + threadIdx.x; i < m1; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 m = matches[i];
        const float2 a = ks[m.x], b = kd[m.y];
        // destination points are normalised with the SOURCE intrinsics (pose_graph_builder.h:908-912)
        corr[i] = make_double4(((double)a.x - cx) / f, ((double)a.y - cy) / f, ((double)b.x - cx) / f, ((double)b.y - cy) / f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) thrNorm[p] = thrPx / ((f + f + f + f) / 4.0);  // :934-937
}

// ---------------------------------------------------------------------------------------------
// K1: one CTA per wave pair.  For every hypothesis: E = [t]x R, squared Sampson residual of every
// correspondence (one LDG.128 x2 per row, coalesced 32 B/thread), count under the squared (test) and
// un-squared (getInliers, GT:164) thresholds with warp ballots; the last passing hypothesis' inlier
// bits are kept for the sampler of K2.  HBM-bound: 32 B / correspondence, read once per hypothesis.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) k1_score_hypotheses(WaveArgs a)
{
    const uint32_t w = blockIdx.x;
    if (w >= a.n) return;
    const uint32_t pid = a.pairId[w];
    const uint64_t r0 = a.offset[pid];
    const uint32_t N = (uint32_t)(a.offset[pid + 1] - r0);
    const double4 *rows = reinterpret_cast<const double4 *>(a.corr) + r0;
    const double thrT = a.thrOverride > 0.0 ? a.thrOverride : a.thrMultiplier * a.thr[pid];  // 1.5 * thr_norm  (PGB:800, :964)
    const double thrTsq = thrT * thrT;                 // GT:184
    // two bit buffers per slot: the one being written and the one of the last hypothesis that reached
    // estimatePose (so a later failing hypothesis cannot clobber the sampler's input)
    uint32_t *bitsBase = a.bits + (size_t)w * a.bitsStride * 2;
    uint8_t *mask = (a.flags & PGI_WAVE_MASKS) ? a.masks + a.maskOffset[w] : nullptr;
    const bool noTest = (a.flags & PGI_WAVE_NO_TEST) != 0;

    __shared__ double sE[9];
    __shared__ uint32_t sCnt[2][kCtaThreads / 32];
    __shared__ uint32_t sDecision, sValidSel;
    if (threadIdx.x == 0) sValidSel = 0;
    const uint32_t h0 = a.hypOffset[w], h1 = a.hypOffset[w + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (mask)
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) mask[i] = 0;

    uint32_t flags = 0, testCount = 0, pathInliers = 0;
    for (uint32_t h = h0; h < h1; ++h) {
        __syncthreads();
        if (threadIdx.x == 0) essentialFromPose(a.hyp + 7 * (size_t)h, sE);
        __syncthreads();
        double E[9];
#pragma unroll
        for (int k = 0; k < 9; k++) E[k] = sE[k];
        const uint32_t cur = 1u - sValidSel;
        uint32_t *bits = bitsBase + (size_t)cur * a.bitsStride;
        uint32_t cTest = 0, cInl = 0;
        const uint32_t nRound = (N + 31u) & ~31u;
        for (uint32_t i = threadIdx.x; i < nRound; i += blockDim.x) {
            bool t = false, in = false;
            if (i < N) {
                const double4 c = rows[i];
                const double s = sampsonSq(c.x, c.y, c.z, c.w, E);
                t = s < thrTsq;  // GT:214
                in = s < thrT;   // GT:164 (un-squared threshold, reproduced)
            }
            const uint32_t bt = __ballot_sync(0xffffffffu, t);
            const uint32_t bi = __ballot_sync(0xffffffffu, in);
            if (lane == 0) {
                cTest += __popc(bt);
                cInl += __popc(bi);
                bits[i >> 5] = bi;
            }
        }
        if (lane == 0) { sCnt[0][warp] = cTest; sCnt[1][warp] = cInl; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0, in = 0;
            for (int k = 0; k < kCtaThreads / 32; k++) { t += sCnt[0][k]; in += sCnt[1][k]; }
            const bool passed = noTest || (t >= a.testMinInliers);  // GT:221
            testCount = t < a.testMinInliers ? t : a.testMinInliers;  // early exit leaves inlierNumber_ at the minimum
            // the loop of PGB:974-1029 lets the LAST guess decide; guesses are the hypotheses that passed
            if (passed) {
                flags = ST_HAVE_HYP | ST_TEST_PASSED;
                pathInliers = in;
            } else
                flags |= ST_HAVE_HYP;
            sDecision = passed ? 1u : 0u;
            atomicAdd(a.counters + 0, (unsigned long long)N);
        }
        __syncthreads();
        // PGB:1000-1009: the mask accumulates the inliers of every guess that reached estimatePose
        if (mask && sDecision)
            for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
                if ((bits[i >> 5] >> (i & 31)) & 1u) mask[i] = 1;
        __syncthreads();
        if (threadIdx.x == 0 && sDecision) sValidSel = cur;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        SlotState &s = a.state[w];
        s.testCount = testCount;
        s.pathInliers = (flags & ST_TEST_PASSED) ? pathInliers : 0u;
        s.inlierCount = 0;
        s.tableIdx = a.pairTable[pid];
        s.pad = sValidSel;  // which bit buffer K2 samples from
        s.models = 0;
        s.it = 0;
        if ((flags & ST_TEST_PASSED) && pathInliers >= 5) flags |= ST_NEED_5PT;  // ptsetreg: count < modelPoints -> fail
        s.flags = flags;
    }
}

// r-th set bit (0-based) of a bit mask.
__device__ inline uint32_t selectRank(const uint32_t *bits, uint32_t nWords, uint32_t r)
{
    uint32_t acc = 0;
    for (uint32_t wI = 0; wI < nWords; ++wI) {
        const uint32_t c = __popc(bits[wI]);
        if (acc + c > r) {
            uint32_t word = bits[wI];
            for (uint32_t k = r - acc; k > 0; --k) word &= word - 1;
            return wI * 32 + (__ffs(word) - 1);
        }
        acc += c;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K2: legacy cv::findEssentialMat(RANSAC, threshold = DBL_MAX) on the inliers of the hypothesis:
// MWC RNG(2^64-1) draws 5 distinct ranks in [0,k); the first solution of the first sample that yields
// a model wins and every point is its inlier.  One thread per wave pair; the
// 1000-sweep Durand-Kerner chain is latency-bound, so waves should be large.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k2_fivept_first_solution(WaveArgs a)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.n) return;
    SlotState &s = a.state[w];
    uint32_t flags = s.flags;
    if (flags & ST_NEED_5PT) {
        const uint32_t pid = a.pairId[w];
        const uint64_t r0 = a.offset[pid];
        const uint32_t N = (uint32_t)(a.offset[pid + 1] - r0);
        const double4 *rows = reinterpret_cast<const double4 *>(a.corr) + r0;
        const uint32_t *bits = a.bits + ((size_t)w * 2 + s.pad) * a.bitsStride;
        const uint32_t nWords = (N + 31) / 32;
        const int k = (int)s.pathInliers;
        double x1[10], x2[10], E[9];
        bool found = false;
        if ((uint32_t)k < a.minInliers) {
            // PGB:1028 fails whatever the kernel returns; E is overwritten by the fallback (PGB:1037).  The
            // 1000-sweep solve is skipped and inlierNumber_ reported as k (what a successful solve gives).
            s.inlierCount = (uint32_t)k;
            flags &= ~ST_NEED_5PT;
        } else if (k == 5) {
            for (int i = 0; i < 5; i++) {
                const double4 c = rows[selectRank(bits, nWords, i)];
                x1[2 * i] = c.x; x1[2 * i + 1] = c.y; x2[2 * i] = c.z; x2[2 * i + 1] = c.w;
            }
            found = fivePoint<true>(x1, x2, E, 1, 1000, 0.0) > 0;
        } else {
            CvRng rng((uint64_t)-1);
            for (int iter = 0; iter < 1000 && !found; iter++) {
                int idx[5];
                for (int i = 0; i < 5; ++i) {
                    int idx_i;
                    for (;;) {
                        idx_i = rng.uniform(0, k);
                        bool dup = false;
                        for (int j = 0; j < i; j++) dup |= (idx[j] == idx_i);
                        if (!dup) break;
                    }
                    idx[i] = idx_i;
                    const double4 c = rows[selectRank(bits, nWords, (uint32_t)idx_i)];
                    x1[2 * i] = c.x; x1[2 * i + 1] = c.y; x2[2 * i] = c.z; x2[2 * i + 1] = c.w;
                }
                found = fivePoint<true>(x1, x2, E, 1, 1000, 0.0) > 0;
            }
        }
        if (!(flags & ST_NEED_5PT)) {
        } else if (found) {
            for (int q = 0; q < 9; q++) s.E[q] = E[q];
            s.inlierCount = (uint32_t)k;  // threshold^2 = inf: every point is an inlier  (PGB:1022)
            if ((uint32_t)k >= a.minInliers) flags |= ST_PATH_OK | ST_HAVE_E;  // PGB:1028
        } else {
            flags |= ST_5PT_NOMODEL;
            s.inlierCount = 0;
        }
    }
    if (!(flags & ST_PATH_OK) && (a.flags & PGI_WAVE_FALLBACK)) {
        const uint32_t pid = a.pairId[w];
        const uint32_t N = (uint32_t)(a.offset[pid + 1] - a.offset[pid]);
        flags |= ST_NEED_FB | ST_FB_RAN;
        if (N >= 5) flags |= ST_FB_ACTIVE;
        s.bestCost = DBL_MAX;
        s.bestInl = 0;
        s.maxIters = (int)a.fbMaxIters;
        s.it = 0;
        s.models = 0;
        for (int q = 0; q < 9; q++) s.bestE[q] = 0.0;
        atomicAdd(a.counters + 1, 1ull);
    }
    s.flags = flags;
}

// ---------------------------------------------------------------------------------------------
// K4: minimal solves of the fallback.  One thread per (wave pair, iteration of the current chunk).
// The uniform sampler's sequence depends on N only (persistent-pool partial Fisher-Yates driven by
// cv::RNG(0)), so the samples come from a per-N table built at registration.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k4_fallback_solve(WaveArgs a, int chunk)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = g / kFbChunk, j = g % kFbChunk;
    if (w >= a.n) return;
    const SlotState &s = a.state[w];
    if (!(s.flags & ST_FB_ACTIVE)) return;
    const int it = chunk * kFbChunk + (int)j;
    if (it >= s.maxIters) return;
    const uint32_t pid = a.pairId[w];
    const double4 *rows = reinterpret_cast<const double4 *>(a.corr) + a.offset[pid];
    const uint32_t *smp = a.samplerTab + ((size_t)s.tableIdx * a.fbMaxIters + it) * 5;
    double x1[10], x2[10];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const double4 c = rows[smp[i]];
        x1[2 * i] = c.x; x1[2 * i + 1] = c.y; x2[2 * i] = c.z; x2[2 * i + 1] = c.w;
    }
    double *out = a.fbSols + ((size_t)w * kFbChunk + j) * 90;
    const int n = fivePoint<false>(x1, x2, out, 10, kDkMaxSweeps, kDkTolSq);
    a.fbCounts[(size_t)w * kFbChunk + j] = (uint8_t)n;
}

// Fixed-order block reduction of one double per thread (256 threads): intra-warp shuffle-down tree,
// then a stride-4/2/1 tree over the 8 warp sums — the order the oracle's fb::treeReduce restates.
__device__ __forceinline__ double blockReduceFixed(double v, double *sWarp /*8*/)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sWarp[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
        double q[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = sWarp[k];
#pragma unroll
        for (int s = 4; s >= 1; s >>= 1)
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < s) q[k] += q[k + s];
        r = q[0];
    }
    return r;  // valid in thread 0
}

__device__ __forceinline__ uint32_t blockSumU32(uint32_t v, uint32_t *sWarp /*8*/)
{
    v = __reduce_add_sync(0xffffffffu, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sWarp[warp] = v;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r += sWarp[k];
    return r;  // valid in all threads
}

// MSAC cost (truncation (1.5 thr)^2) + inlier count at thr^2 of one model; thread t owns points
// i = t, t+256, ... in ascending order.  Returns cost in thread 0, inlier count in all threads.
__device__ __forceinline__ void scoreModelBlock(const double4 *rows, uint32_t N, const double E[9], double thrSq,
                                                double truncSq, double *sWarpD, uint32_t *sWarpU, double &cost,
                                                uint32_t &inl)
{
    double part = 0.0;
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < N; i += kCtaThreads) {
        const double4 c = rows[i];
        const double r = sampsonSq(c.x, c.y, c.z, c.w, E);
        part += (r < truncSq) ? r : truncSq;
        cnt += (r < thrSq) ? 1u : 0u;
    }
    cost = blockReduceFixed(part, sWarpD);
    inl = blockSumU32(cnt, sWarpU);
}

// Symmetric 9x9 cyclic Jacobi, eigenvector of the smallest eigenvalue (single thread; mirrors the oracle).
__device__ inline void smallestEigvec9(double *A /*81*/, double *V /*81*/, double v[9])
{
    for (int i = 0; i < 9; i++)
        for (int j = 0; j < 9; j++) V[i * 9 + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 9; i++) {
            diag += A[i * 9 + i] * A[i * 9 + i];
            for (int j = i + 1; j < 9; j++) off += A[i * 9 + j] * A[i * 9 + j];
        }
        if (off <= 1e-30 * diag) break;
        for (int p = 0; p < 8; p++)
            for (int q = p + 1; q < 9; q++) {
                const double apq = A[p * 9 + q];
                if (apq == 0.0) continue;
                const double app = A[p * 9 + p], aqq = A[q * 9 + q];
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < 9; k++) {
                    const double akp = A[k * 9 + p], akq = A[k * 9 + q];
                    A[k * 9 + p] = c * akp - s * akq;
                    A[k * 9 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 9; k++) {
                    const double apk = A[p * 9 + k], aqk = A[q * 9 + k];
                    A[p * 9 + k] = c * apk - s * aqk;
                    A[q * 9 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 9; k++) {
                    const double vkp = V[k * 9 + p], vkq = V[k * 9 + q];
                    V[k * 9 + p] = c * vkp - s * vkq;
                    V[k * 9 + q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    for (int i = 1; i < 9; i++)
        if (A[i * 9 + i] < A[m * 9 + m]) m = i;
    for (int k = 0; k < 9; k++) v[k] = V[k * 9 + m];
}

// Least-squares refit on the inliers of Ecur (8-point normal equations, essential projection).
// Whole CTA cooperates; result valid in shared sEls / *ok (after the trailing __syncthreads).
__device__ inline void lsRefitBlock(const double4 *rows, uint32_t N, const double Ecur[9], double thrSq, double *sM /*81*/,
                                    double *sV /*81*/, double *sWarpD, uint32_t *sWarpU, double *sEls /*9*/, int *sOk)
{
    // 9 passes: pass r accumulates the (9 - r) sums  sum a_r a_s, s >= r  (thread t: points t, t+256, ...)
    uint32_t cntLocal = 0;
    for (int r = 0; r < 9; r++) {
        double acc[9];
#pragma unroll
        for (int s = 0; s < 9; s++) acc[s] = 0.0;
        for (uint32_t i = threadIdx.x; i < N; i += kCtaThreads) {
            const double4 c = rows[i];
            if (!(sampsonSq(c.x, c.y, c.z, c.w, Ecur) < thrSq)) continue;
            if (r == 0) cntLocal++;
            const double av[9] = {c.z * c.x, c.z * c.y, c.z, c.w * c.x, c.w * c.y, c.w, c.x, c.y, 1.0};
            double ar = av[0];
#pragma unroll
            for (int s = 1; s < 9; s++) ar = (s == r) ? av[s] : ar;
#pragma unroll
            for (int s = 0; s < 9; s++)
                if (s >= r) acc[s] += ar * av[s];
        }
        for (int s = r; s < 9; s++) {
            const double v = blockReduceFixed(acc[s], sWarpD);
            if (threadIdx.x == 0) { sM[r * 9 + s] = v; sM[s * 9 + r] = v; }
        }
    }
    const uint32_t cnt = blockSumU32(cntLocal, sWarpU);
    if (threadIdx.x == 0) {
        int ok = 0;
        if (cnt >= (uint32_t)kLoMinInliers) {
            double ev[9];
            smallestEigvec9(sM, sV, ev);
            double U[9], V[9], S[3];
            eigenJacobiSvd<3, true, true>(ev, U, V, S);
            if (S[1] > 0.0) {
                double nrm = 0.0, Eo[9];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        const double x = U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1];
                        Eo[i * 3 + j] = x;
                        nrm += x * x;
                    }
                nrm = sqrt(nrm);
                ok = 1;
                for (int k = 0; k < 9; k++) {
                    Eo[k] = Eo[k] / nrm;
                    if (Eo[k] != Eo[k]) ok = 0;
                    sEls[k] = Eo[k];
                }
            }
        }
        *sOk = ok;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// K5: one CTA per wave pair; walks the chunk's iterations IN ORDER (best-so-far / LO / termination
// semantics of a sequential RANSAC), scoring every minimal model over all N correspondences.
// Correspondences are streamed from L2/HBM (32 B / row / model).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) k5_fallback_score(WaveArgs a, int chunk, int lastChunk)
{
    const uint32_t w = blockIdx.x;
    if (w >= a.n) return;
    SlotState &st = a.state[w];
    const uint32_t flags0 = st.flags;
    if (!(flags0 & ST_NEED_FB)) return;
    const uint32_t pid = a.pairId[w];
    const uint64_t r0 = a.offset[pid];
    const uint32_t N = (uint32_t)(a.offset[pid + 1] - r0);
    const double4 *rows = reinterpret_cast<const double4 *>(a.corr) + r0;
    const double thr = a.thr[pid];
    const double thrSq = thr * thr;
    const double trunc = 1.5 * thr;
    const double truncSq = trunc * trunc;

    __shared__ double sWarpD[8];
    __shared__ uint32_t sWarpU[8];
    __shared__ double sM[81], sV[81], sEls[9], sBestE[9];
    __shared__ double sBestCost;
    __shared__ int sBestInl, sMaxIters, sOk, sUpdated, sHave;

    if (threadIdx.x == 0) {
        sBestCost = st.bestCost;
        sBestInl = st.bestInl;
        sMaxIters = st.maxIters;
        sHave = st.bestCost < DBL_MAX ? 1 : 0;
        for (int k = 0; k < 9; k++) sBestE[k] = st.bestE[k];
    }
    __syncthreads();
    uint32_t models = 0;
    int it = st.it;
    if (flags0 & ST_FB_ACTIVE) {
        const uint16_t *itTab = a.itersTab + a.itersTabOff[st.tableIdx];
        const int itEnd = (chunk + 1) * kFbChunk;
        for (; it < itEnd && it < sMaxIters; ++it) {
            const int j = it - chunk * kFbChunk;
            const int ns = a.fbCounts[(size_t)w * kFbChunk + j];
            const double *sols = a.fbSols + ((size_t)w * kFbChunk + j) * 90;
            if (threadIdx.x == 0) sUpdated = 0;
            for (int s = 0; s < ns; s++) {
                double E[9];
#pragma unroll
                for (int k = 0; k < 9; k++) E[k] = sols[s * 9 + k];
                double cost;
                uint32_t inl;
                scoreModelBlock(rows, N, E, thrSq, truncSq, sWarpD, sWarpU, cost, inl);
                models++;
                if (threadIdx.x == 0 && cost < sBestCost) {
                    sBestCost = cost;
                    sBestInl = (int)inl;
                    for (int k = 0; k < 9; k++) sBestE[k] = E[k];
                    sUpdated = 1;
                    sHave = 1;
                }
            }
            __syncthreads();
            if (sUpdated) {
                for (int r = 0; r < kLoRounds; r++) {
                    double Eb[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) Eb[k] = sBestE[k];
                    lsRefitBlock(rows, N, Eb, thrSq, sM, sV, sWarpD, sWarpU, sEls, &sOk);
                    if (!sOk) break;
                    double El[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) El[k] = sEls[k];
                    double cost;
                    uint32_t inl;
                    scoreModelBlock(rows, N, El, thrSq, truncSq, sWarpD, sWarpU, cost, inl);
                    models++;
                    if (threadIdx.x == 0) {
                        if (cost < sBestCost) {
                            sBestCost = cost;
                            sBestInl = (int)inl;
                            for (int k = 0; k < 9; k++) sBestE[k] = El[k];
                            sOk = 1;
                        } else
                            sOk = 0;
                    }
                    __syncthreads();
                    if (!sOk) break;
                }
                if (threadIdx.x == 0) {
                    const int lim = (int)itTab[sBestInl];
                    sMaxIters = sMaxIters < lim ? sMaxIters : lim;
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    const bool done = !(flags0 & ST_FB_ACTIVE) || it >= sMaxIters || lastChunk;
    // finalise: mask == sampson^2(E_final) < thr^2 and its count
    uint32_t finalInl = 0;
    if (done) {
        uint8_t *mask = (a.flags & PGI_WAVE_MASKS) ? a.masks + a.maskOffset[w] : nullptr;
        double E[9];
#pragma unroll
        for (int k = 0; k < 9; k++) E[k] = sBestE[k];
        uint32_t cnt = 0;
        for (uint32_t i = threadIdx.x; i < N; i += kCtaThreads) {
            uint32_t in = 0;
            if (sHave) {
                const double4 c = rows[i];
                in = sampsonSq(c.x, c.y, c.z, c.w, E) < thrSq ? 1u : 0u;
            }
            if (mask) mask[i] = (uint8_t)in;
            cnt += in;
        }
        finalInl = blockSumU32(cnt, sWarpU);
    }
    if (threadIdx.x == 0) {
        st.bestCost = sBestCost;
        st.bestInl = sBestInl;
        st.maxIters = sMaxIters;
        st.it = it;
        st.models += models;
        for (int k = 0; k < 9; k++) st.bestE[k] = sBestE[k];
        uint32_t f = flags0;
        if (done) {
            f &= ~ST_FB_ACTIVE;
            f &= ~ST_NEED_FB;
            st.inlierCount = finalInl;  // PGB:1047-1048
            if (sHave && finalInl >= a.minInliers) {  // PGB:1053
                f |= ST_HAVE_E;
                for (int k = 0; k < 9; k++) st.E[k] = sBestE[k];
            }
            atomicAdd(a.counters + 2, (unsigned long long)st.models);
        }
        st.flags = f;
    }
}

// Whole-CTA E -> candidates -> triangulation vote (pose_utils.h:172-240).  Results in shared memory.
__device__ inline void decomposeVoteBlock(const double E[9], const double4 *rows, uint32_t N, double (*sR)[9], double *sT,
                                          uint32_t *sVotes)
{
    if (threadIdx.x == 0) {
        double R1[9], R2[9], t[3];
        decomposeEssential(E, R1, R2, t);
        for (int k = 0; k < 9; k++) { sR[0][k] = R1[k]; sR[1][k] = R2[k]; }
        for (int k = 0; k < 3; k++) sT[k] = t[k];
    }
    if (threadIdx.x < 4) sVotes[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double4 c = rows[i];
        double best = DBL_MAX;
        int bestPose = 5;
#pragma unroll 1
        for (int cand = 0; cand < 4; cand++) {
            const double sgn = (cand & 1) ? -1.0 : 1.0;
            const double *R = sR[cand >> 1];
            double P2[12];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                P2[r * 4 + 0] = R[r * 3 + 0];
                P2[r * 4 + 1] = R[r * 3 + 1];
                P2[r * 4 + 2] = R[r * 3 + 2];
                P2[r * 4 + 3] = sgn * sT[r];
            }
            double err;
            if (!triangulateAndScore(P2, c.x, c.y, c.z, c.w, err)) continue;
            if (err < best) {
                best = err;
                bestPose = cand;
            }
        }
        if (bestPose < 5) atomicAdd(&sVotes[bestPose], 1u);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// K3 (+K8): E -> (R1,R2,+-t) by the Eigen-faithful 3x3 Jacobi SVD, then the triangulation vote over
// ALL correspondences: 4 candidates x N 4x4 Jacobi SVDs, register resident, one thread per
// correspondence (looping the 4 candidates so the strict-< winner logic of pose_utils.h:226-230 is
// local).  FP64-pipe bound.  Thread 0 picks the first maximum, converts to a unit quaternion and
// packs the 160-byte verdict.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) k3_decompose_vote(WaveArgs a)
{
    const uint32_t w = blockIdx.x;
    if (w >= a.n) return;
    const SlotState &st = a.state[w];
    const uint32_t pid = a.pairId[w];
    const uint64_t r0 = a.offset[pid];
    const uint32_t N = (uint32_t)(a.offset[pid + 1] - r0);
    const uint32_t flags = st.flags;

    __shared__ double sR[2][9], sT[3];
    __shared__ uint32_t sVotes[4];
    if (flags & ST_HAVE_E) {
        double E[9];
        for (int k = 0; k < 9; k++) E[k] = st.E[k];
        decomposeVoteBlock(E, reinterpret_cast<const double4 *>(a.corr) + r0, N, sR, sT, sVotes);
    }
    if (threadIdx.x == 0) {
        pgi_verdict v;
        v.pair_id = pid;
        v.n_corr = N;
        v.n_hypotheses = (uint8_t)min(255u, a.hypOffset[w + 1] - a.hypOffset[w]);
        v.test_passed = (flags & ST_TEST_PASSED) ? 1 : 0;
        v.test_count = st.testCount;
        v.path_inliers = st.pathInliers;
        v.inlier_count = st.inlierCount;
        v.iters = (flags & ST_FB_RAN) ? (uint32_t)st.it : 0u;
        v.status = ((flags & ST_FB_RAN) ? 1u : 0u) | ((flags & ST_5PT_NOMODEL) ? 2u : 0u);
        v.branch = 0;
        v.accepted = 0;
        for (int k = 0; k < 9; k++) v.E[k] = 0.0;
        v.q[0] = v.q[1] = v.q[2] = 0.0; v.q[3] = 1.0;
        v.t[0] = v.t[1] = v.t[2] = 0.0;
        if (flags & ST_HAVE_E) {
            for (int k = 0; k < 9; k++) v.E[k] = st.E[k];
            int maxIdx = 0;
            for (int i = 1; i < 4; i++)
                if (sVotes[i] > sVotes[maxIdx]) maxIdx = i;  // std::max_element: first maximum (PU:243-246)
            double R[9], t[3];
            const double sgn = (maxIdx & 1) ? -1.0 : 1.0;
            for (int k = 0; k < 9; k++) R[k] = sR[maxIdx >> 1][k];
            for (int k = 0; k < 3; k++) t[k] = sgn * sT[k];
            bool nan = false;
            for (int k = 0; k < 9; k++) nan |= (R[k] != R[k]);
            for (int k = 0; k < 3; k++) nan |= (t[k] != t[k]);
            if (nan) {  // PGB:1069-1070
                v.status |= 4u;
            } else {
                rotationToUnitQuat(R, v.q);
                v.t[0] = t[0]; v.t[1] = t[1]; v.t[2] = t[2];
                v.accepted = 1;
                v.branch = (flags & ST_PATH_OK) ? 1 : 2;
            }
        }
        a.verdicts[w] = v;
    }
}

// Stand-alone E -> (R, t, votes) for the parity tests (pgi_dbg_pose_from_essential).
__global__ void __launch_bounds__(kCtaThreads) kdbg_pose_from_essential(const double *E9, const double4 *rows, uint32_t N,
                                                                       double *Rout, double *tout, unsigned long long *votes)
{
    __shared__ double sR[2][9], sT[3];
    __shared__ uint32_t sVotes[4];
    double E[9];
    for (int k = 0; k < 9; k++) E[k] = E9[k];
    decomposeVoteBlock(E, rows, N, sR, sT, sVotes);
    if (threadIdx.x == 0) {
        int maxIdx = 0;
        for (int i = 1; i < 4; i++)
            if (sVotes[i] > sVotes[maxIdx]) maxIdx = i;
        const double sgn = (maxIdx & 1) ? -1.0 : 1.0;
        for (int k = 0; k < 9; k++) Rout[k] = sR[maxIdx >> 1][k];
        for (int k = 0; k < 3; k++) tout[k] = sgn * sT[k];
        for (int k = 0; k < 4; k++) votes[k] = sVotes[k];
    }
}

// FP64 non-fused peak probe (DMUL+DADD chain) / fused (DFMA) — roofline denominator for K2-K5.
__global__ void k_fp64_peak(double *out, int iters, int fused)
{
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
    double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 1.0000001, c = 1e-9;
    if (fused) {
        for (int i = 0; i < iters; i++) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    } else {
        for (int i = 0; i < iters; i++) {
            a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c);
            a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
            a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
            a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace pgi
