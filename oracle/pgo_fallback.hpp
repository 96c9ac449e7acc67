// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_fallback.hpp — the robust-estimation fallback, SURVEY §8a row a8:
//   cv::findEssentialMat(x1, x2, I, cv::USAC_MAGSAC, 0.99, thr_norm, mask)   pose_graph_builder.h:1037-1044
// OpenCV's USAC is an un-vendored, un-pinned black box whose source is not in this container
// (SURVEY §0.4, App. B.5), so this is the TIER-B restatement the survey prescribes: a robust loop
// with the *observable* structure of the reference call —
//   * uniform sampler = partial Fisher-Yates over a persistent pool driven by cv::RNG(0 -> 0xffffffff)
//     [probe-verified sample sequence, SURVEY App. B.5],
//   * maxIterations = 1000, confidence = 0.99, minimal sample = 5 points, five-point minimal solver,
//   * returned mask == (sampson^2(E_final) < thr^2) [probe-verified],
//   * accept iff sum(mask) >= minimum_inlier_number (pose_graph_builder.h:1047-1054),
// with truncated-quadratic (MSAC) scoring and least-squares local optimisation in place of MAGSAC's
// sigma-consensus (BASELINE.json north_star: "hypothesis scoring, and least-squares LO refits").
// PARITY UNPINNED w.r.t. cv2 for this stage: agreement with cv2.findEssentialMat(USAC_MAGSAC) is
// statistical only and is reported by tests/test_oracle_golden.py (the Tier-B test) and scripts/cv2_host_report.py
// (profiles/r02_cv2_host_report_cfg1_50v.json), never assumed.  The CUDA fallback
// kernels must match THIS restatement bit-for-bit.  To make that independent of any summation order, every
// reduction over correspondences is done in exact 64-bit FIXED POINT: each FP64 term is computed with IEEE
// operations, converted to an integer once, and the integers are summed (associative, so the CUDA kernels are
// free to reduce in any order / with atomics).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pgo_cv.hpp"
#include "pgo_eigen.hpp"
#include "pgo_geom.hpp"

namespace pgo {
namespace fb {

constexpr int kMaxIters = 1000;       // cv::findEssentialMat default maxIters for the USAC overload
constexpr double kConfidence = 0.99;  // pose_graph_builder.h:1042
constexpr int kLoRounds = 4;          // max least-squares refits after each new best model
constexpr int kLoMinInliers = 9;      // need > 8 points for the linear fit
constexpr int kDkMaxSweeps = 200;     // Durand-Kerner sweeps for the fallback's minimal solver
constexpr double kDkTolSq = 1e-22;    // stop when max |correction|^2 <= this (quadratic convergence: the next error is ~1e-22)
constexpr double kCostOne = 4294967296.0;       // MSAC term of an outlier in fixed point (2^32)
constexpr double kLsScale = 1099511627776.0;    // 2^40: fixed-point scale of the normal-equation products

// Sampler table: iteration -> 5 indices.  Depends on N only.
inline void samplerTable(int N, int iters, std::vector<uint32_t> &out)
{
    out.resize((size_t)iters * 5);
    std::vector<int> pool(N);
    for (int i = 0; i < N; i++) pool[i] = i;
    cvx::RNG rng(0);
    for (int it = 0; it < iters; it++) {
        int size = N;
        for (int i = 0; i < 5; i++) {
            const int j = rng.uniform(0, size);
            out[(size_t)it * 5 + i] = (uint32_t)pool[j];
            std::swap(pool[j], pool[--size]);
        }
    }
}

// Standard RANSAC termination: iterations needed to see an all-inlier sample with prob. kConfidence.
// Tabulated on the host (glibc log) for inlier counts 0..N; the CUDA path consumes the same table.
inline void itersTable(int N, std::vector<uint16_t> &out)
{
    out.resize(N + 1);
    for (int k = 0; k <= N; k++) {
        const double w = (double)k / (double)N;
        const double p5 = w * w * w * w * w;
        int it;
        if (p5 <= 0.0)
            it = kMaxIters;
        else if (p5 >= 1.0)
            it = 1;
        else {
            const double v = std::log(1.0 - kConfidence) / std::log(1.0 - p5);
            it = v >= (double)kMaxIters ? kMaxIters : (int)std::ceil(v);
            if (it < 1) it = 1;
        }
        out[k] = (uint16_t)it;
    }
}

// MSAC cost with truncation T = (1.5 thr)^2 and inlier count at thr^2.  Fixed point: an outlier (or NaN residual)
// costs 2^32, an inlier (uint64)(r^2 * (1/T) * 2^32).
inline void scoreModel(const double *corr, int N, const double E[9], double thrSq, double truncSq, uint64_t &cost,
                       int &inliers)
{
    const double invT = 1.0 / truncSq;
    uint64_t acc = 0;
    int cnt = 0;
    for (int i = 0; i < N; i++) {
        const double r = sampsonSq(corr + 4 * i, E);
        acc += (r < truncSq) ? (uint64_t)(r * invT * kCostOne) : (uint64_t)kCostOne;
        cnt += (r < thrSq) ? 1 : 0;
    }
    cost = acc;
    inliers = cnt;
}

// Symmetric 9x9 eigen-decomposition by cyclic Jacobi (fixed sweep count policy), returns the
// eigenvector of the smallest eigenvalue.  A is destroyed.
inline void smallestEigvec9(double A[81], double v[9])
{
    double V[81];
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
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
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

// Least-squares E on the points with sampson^2(Ecur) < thrSq: 8-point normal equations
// (rows [x2x1 x2y1 x2 y2x1 y2y1 y2 x1 y1 1]) + projection onto the essential manifold
// (U diag(1,1,0) V^T with the Eigen-style 3x3 JacobiSVD), unit Frobenius norm.
inline bool lsRefit(const double *corr, int N, const double Ecur[9], double thrSq, double Eout[9])
{
    // 45 upper-triangular sums in 2^-40 fixed point (order-free)
    int64_t acc[45];
    for (int e = 0; e < 45; e++) acc[e] = 0;
    int cnt = 0;
    for (int i = 0; i < N; i++) {
        const double *c = corr + 4 * i;
        if (!(sampsonSq(c, Ecur) < thrSq)) continue;
        cnt++;
        const double a[9] = {c[2] * c[0], c[2] * c[1], c[2], c[3] * c[0], c[3] * c[1], c[3], c[0], c[1], 1.0};
        int e = 0;
        for (int r = 0; r < 9; r++)
            for (int q = r; q < 9; q++, e++) acc[e] += (int64_t)std::llrint(a[r] * a[q] * kLsScale);
    }
    if (cnt < kLoMinInliers) return false;
    double M[81];
    int e = 0;
    for (int r = 0; r < 9; r++)
        for (int q = r; q < 9; q++, e++) {
            const double v = (double)acc[e] / kLsScale;
            M[r * 9 + q] = v;
            M[q * 9 + r] = v;
        }
    double ev[9];
    smallestEigvec9(M, ev);
    double U[9], V[9], S[3];
    eig::jacobiSvd<3>(ev, U, V, S);
    if (!(S[1] > 0.0)) return false;
    // E = U diag(1,1,0) V^T
    double nrm = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double x = U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1];
            Eout[i * 3 + j] = x;
            nrm += x * x;
        }
    nrm = std::sqrt(nrm);
    for (int k = 0; k < 9; k++) Eout[k] = Eout[k] / nrm;
    for (int k = 0; k < 9; k++)
        if (Eout[k] != Eout[k]) return false;
    return true;
}

struct FallbackResult {
    int ok = 0;          // a model was found
    double E[9] = {0};
    int inliers = 0;     // sum(mask)
    int iterations = 0;  // iterations executed
    int models = 0;      // minimal models scored
    int loRuns = 0;      // LO refits scored
    uint64_t cost = UINT64_MAX;
};

// Five-point minimal solver for the fallback: same algebra as cvx::fivePointKernel but with a
// tolerance-stopped Durand-Kerner (kDkMaxSweeps / kDkTolSq) instead of the legacy 1000 fixed sweeps.
inline int fivePointFallback(const double *x1, const double *x2, double *Eout)
{
    return cvx::fivePointKernel(x1, x2, 5, Eout, kDkMaxSweeps, kDkTolSq);
}

inline FallbackResult runFallback(const double *corr, int N, double thr, uint8_t *mask /*N or null*/)
{
    FallbackResult res;
    if (N < 5) {
        if (mask) std::memset(mask, 0, N);
        return res;
    }
    const double thrSq = thr * thr;
    const double trunc = 1.5 * thr;
    const double truncSq = trunc * trunc;
    std::vector<uint32_t> samples;
    samplerTable(N, kMaxIters, samples);
    std::vector<uint16_t> iters;
    itersTable(N, iters);

    int maxIters = kMaxIters;
    uint64_t bestCost = UINT64_MAX;
    int bestInl = 0;
    double bestE[9] = {0};
    bool have = false;
    int it = 0;
    for (; it < maxIters; it++) {
        double x1[10], x2[10];
        for (int i = 0; i < 5; i++) {
            const double *c = corr + 4 * samples[(size_t)it * 5 + i];
            x1[2 * i] = c[0]; x1[2 * i + 1] = c[1]; x2[2 * i] = c[2]; x2[2 * i + 1] = c[3];
        }
        double sols[90];
        const int ns = fivePointFallback(x1, x2, sols);
        bool updated = false;
        for (int s = 0; s < ns; s++) {
            uint64_t cost;
            int inl;
            scoreModel(corr, N, sols + 9 * s, thrSq, truncSq, cost, inl);
            res.models++;
            if (cost < bestCost) {
                bestCost = cost;
                bestInl = inl;
                std::memcpy(bestE, sols + 9 * s, sizeof(bestE));
                have = true;
                updated = true;
            }
        }
        if (updated) {
            for (int r = 0; r < kLoRounds; r++) {
                double Els[9];
                if (!lsRefit(corr, N, bestE, thrSq, Els)) break;
                uint64_t cost;
                int inl;
                scoreModel(corr, N, Els, thrSq, truncSq, cost, inl);
                res.loRuns++;
                if (!(cost < bestCost)) break;
                bestCost = cost;
                bestInl = inl;
                std::memcpy(bestE, Els, sizeof(bestE));
            }
            maxIters = std::min(maxIters, (int)iters[bestInl]);
        }
    }
    res.iterations = it;
    res.ok = have ? 1 : 0;
    res.cost = bestCost;
    if (have) {
        std::memcpy(res.E, bestE, sizeof(bestE));
        int cnt = 0;
        for (int i = 0; i < N; i++) {
            const int in = sampsonSq(corr + 4 * i, bestE) < thrSq ? 1 : 0;
            if (mask) mask[i] = (uint8_t)in;
            cnt += in;
        }
        res.inliers = cnt;
    } else if (mask) {
        std::memset(mask, 0, N);
    }
    return res;
}

}  // namespace fb
}  // namespace pgo
