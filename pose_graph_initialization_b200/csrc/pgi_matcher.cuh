// pgi_matcher.cuh — K7: epipolar-hashing guided matcher.
//
//   HashingBasedMatcherWithPose<false, 45>::match     matcher.h:199-405
//   (called from PoseGraphBuilder::guidedMatching       pose_graph_builder.h:717-783)
//
// One CTA per image pair.
//   phase 0 (thread 0): F = Kd^-T E Ks^-1, epipole from the 3x3 Jacobi SVD of F, angular range from the destination
//           image's corners (matcher.h:218-277) — the same operations in the same order as oracle/pgo_matcher.hpp;
//   phase 1: every destination keypoint gets the bin of its epipolar line's normal angle (matcher.h:285-299);
//   phase 2: the bins are laid out as index lists in keypoint order (the reference appends in that order, and the order
//           decides ties between equal descriptor distances): warp w gathers bins w, w + 8, ... with ballot compaction;
//   phase 3: one thread per source keypoint walks its bin: squared symmetric epipolar distance (FP64, matcher.h:351-360)
//           against 0.75^2, then the 128-D squared descriptor distance (float differences accumulated in double,
//           matcher.h:371-375), best / second best, count-corrected Lowe ratio (matcher.h:386-406);
//   phase 4: accepted matches are compacted in source order (block scan).
// atan2 is evaluated in double-double arithmetic and rounded once (pgi_atan2.h): the correctly rounded value, which is
// what glibc returns for all but ~2.5e-4 of arguments (CUDA's own atan2 is a 2-ulp function and differed from the host
// in the angular range of one of five test scenes).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pgi_atan2.h"
#include "pgi_math.cuh"

namespace pgi {

constexpr int kMatchThreads = 256;
constexpr int kMatchMaxBins = 192;

struct MatchArgs {
    const float2 *kpS;   // nS source keypoints (pixels)
    const float *dS;     // nS x dim
    const float2 *kpD;
    const float *dD;
    uint32_t nS, nD, dim;
    double E[9], Ks[9], Kd[9];
    int wS, hS, wD, hD, binNumber;
    uint8_t *binOfD;     // nD scratch
    uint32_t *binList;   // nD scratch: bins concatenated
    uint32_t *cand;      // nS scratch: best index (0xffffffff: none)
    double *candRatio;   // nS scratch
    uint32_t *matches;   // out: n x 2
    double *ratios;      // out
    uint32_t *nOut;
    double *prepOut;     // 14 doubles (diagnostics / parity)
};

PGI_DEV void inverse3(const double *m, double *inv)  // Eigen compute_inverse_size3 (see oracle/pgo_matcher.hpp)
{
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
    const double invdet = 1.0 / det;
    inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
    inv[3] = cof(0, 1) * invdet; inv[4] = cof(1, 1) * invdet; inv[5] = cof(2, 1) * invdet;
    inv[6] = cof(0, 2) * invdet; inv[7] = cof(1, 2) * invdet; inv[8] = cof(2, 2) * invdet;
}
PGI_DEV void mul3(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = sum3(A[i * 3 + 0] * B[0 * 3 + j], A[i * 3 + 1] * B[1 * 3 + j], A[i * 3 + 2] * B[2 * 3 + j]);
}
PGI_DEV double lineAngle(double ny, double nx)
{
    const double kRadianToDegree = 180.0 / 3.14159265358979323846;
    double angle = kRadianToDegree * pgi_atan::atan2cr(ny, nx) + 180.0;
    if (angle > 180) angle -= 180;
    return angle;
}
PGI_DEV int binOfAngle(double angle, double minAngle, double angularRange, int binNumber)
{
    angle = (binNumber - 1) * (angle - minAngle) / angularRange;
    const int b = (int)round(angle);
    const int lo = b < 0 ? 0 : b;  // MAX(0, .)
    return lo < binNumber - 1 ? lo : binNumber - 1;
}

__global__ void __launch_bounds__(kMatchThreads) k7_guided_match(MatchArgs a)
{
    __shared__ double sF[9], sEpi[2], sMinAngle, sRange;
    __shared__ int sBins;
    __shared__ uint32_t sBinStart[kMatchMaxBins + 1], sBinCount[kMatchMaxBins];
    __shared__ uint32_t sScan[kMatchThreads / 32], sTotal;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        double KdInv[9], KsInv[9], KdInvT[9], T[9], F[9];
        inverse3(a.Kd, KdInv);
        inverse3(a.Ks, KsInv);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) KdInvT[i * 3 + j] = KdInv[j * 3 + i];
        mul3(KdInvT, a.E, T);
        mul3(T, KsInv, F);
        double U[1], V[9], S[3];
        eigenJacobiSvd<3, false, true>(F, U, V, S);
        const double ez = V[8];
        const double ex = V[2] / ez, ey = V[5] / ez;
        const bool inImage = ex >= 0 && ex < a.wS && ey >= 0 && ey < a.hS;
        double minAngle = 180, maxAngle = 0;
        if (!inImage) {
            const double corners[8] = {0, 0, (double)a.wD, 0, (double)a.wD, (double)a.hD, 0, (double)a.hD};
            for (int c = 0; c < 8; c += 2) {
                const double x = corners[c], y = corners[c + 1];
                const double nx = F[0] * x + F[3] * y + F[6];
                const double ny = F[1] * x + F[4] * y + F[7];
                const double angle = lineAngle(ny, nx);
                minAngle = minAngle < angle ? minAngle : angle;
                maxAngle = maxAngle > angle ? maxAngle : angle;
            }
        }
        for (int k = 0; k < 9; k++) sF[k] = F[k];
        sEpi[0] = ex; sEpi[1] = ey;
        sMinAngle = minAngle;
        sRange = maxAngle - minAngle;
        int bins = a.binNumber <= 0 ? (int)sRange : a.binNumber;
        sBins = bins < 0 ? 0 : (bins > kMatchMaxBins ? kMatchMaxBins : bins);
        if (a.prepOut) {
            for (int k = 0; k < 9; k++) a.prepOut[k] = F[k];
            a.prepOut[9] = ex; a.prepOut[10] = ey; a.prepOut[11] = minAngle; a.prepOut[12] = sRange; a.prepOut[13] = bins;
        }
    }
    for (int b = threadIdx.x; b < kMatchMaxBins; b += kMatchThreads) sBinCount[b] = 0;
    __syncthreads();
    const int bins = sBins;
    if (bins <= 0) {
        if (threadIdx.x == 0) *a.nOut = 0;
        return;
    }
    const double F0 = sF[0], F1 = sF[1], F2 = sF[2], F3 = sF[3], F4 = sF[4], F5 = sF[5], F6 = sF[6], F7 = sF[7], F8 = sF[8];
    const double minAngle = sMinAngle, range = sRange;
    // ---- phase 1: bin of every destination keypoint
    for (uint32_t i = threadIdx.x; i < a.nD; i += kMatchThreads) {
        const float2 p = a.kpD[i];
        const double x = p.x, y = p.y;
        const double nx = F0 * x + F3 * y + F6;
        const double ny = F1 * x + F4 * y + F7;
        const int b = binOfAngle(lineAngle(ny, nx), minAngle, range, bins);
        a.binOfD[i] = (uint8_t)b;
        atomicAdd(&sBinCount[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < bins; b++) { sBinStart[b] = acc; acc += sBinCount[b]; }
        sBinStart[bins] = acc;
    }
    __syncthreads();
    // ---- phase 2: bins as index lists in keypoint order
    for (int b = warp; b < bins; b += kMatchThreads / 32) {
        uint32_t out = sBinStart[b];
        for (uint32_t base = 0; base < a.nD; base += 32) {
            const uint32_t i = base + lane;
            const bool in = i < a.nD && a.binOfD[i] == (uint8_t)b;
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (in) a.binList[out + __popc(m & ((1u << lane) - 1u))] = i;
            out += __popc(m);
        }
    }
    __syncthreads();
    // ---- phase 3: best / second best per source keypoint
    for (uint32_t i = threadIdx.x; i < a.nS; i += kMatchThreads) {
        const float2 p = a.kpS[i];
        const double x1 = p.x, y1 = p.y;
        const double vx = x1 - sEpi[0], vy = y1 - sEpi[1];
        const int bin = binOfAngle(lineAngle(vx, -vy), minAngle, range, bins);
        double second = DBL_MAX, best = DBL_MAX;
        int bestIndex = -1, countSnn = 0;
        const float *ds = a.dS + (size_t)i * a.dim;
        for (uint32_t k = sBinStart[bin]; k < sBinStart[bin + 1]; k++) {
            const uint32_t nb = a.binList[k];
            const float2 q = a.kpD[nb];
            const double x2 = q.x, y2 = q.y;
            const double rxc = F0 * x2 + F3 * y2 + F6;
            const double ryc = F1 * x2 + F4 * y2 + F7;
            const double rwc = F2 * x2 + F5 * y2 + F8;
            const double r = (x1 * rxc + y1 * ryc + rwc);
            const double rx = F0 * x1 + F1 * y1 + F2;
            const double ry = F3 * x1 + F4 * y1 + F5;
            const double a1 = rxc * rxc + ryc * ryc;
            const double b1 = rx * rx + ry * ry;
            const double d2 = r * r * (a1 + b1) / (a1 * b1);
            if (d2 >= 0.75 * 0.75) continue;
            countSnn += 1;
            const float *dd = a.dD + (size_t)nb * a.dim;
            double acc = 0;
            for (uint32_t m = 0; m < a.dim; m++) {
                const double dist = (double)(ds[m] - dd[m]);  // float subtraction, widened (matcher.h:373)
                acc += dist * dist;
            }
            if (acc < best) { second = best; best = acc; bestIndex = (int)nb; }
        }
        double corr = 1.0;
        if (countSnn < 20) corr = 0.65 * 0.65;
        if (countSnn < 10) corr = 0.6 * 0.6;
        if (countSnn < 5) corr = 0.5 * 0.5;
        if (countSnn < 3) corr = 0.25 * 0.25;
        const double adapted = (best / second) / corr;
        bool ok = false;
        if (!(adapted < 0.00001)) ok = bestIndex > -1 && ((adapted < 0.8 * 0.8) || (countSnn == 1));
        a.cand[i] = ok ? (uint32_t)bestIndex : 0xffffffffu;
        a.candRatio[i] = adapted;
    }
    __syncthreads();
    // ---- phase 4: compaction in source order
    uint32_t base = 0;
    for (uint32_t i0 = 0; i0 < a.nS; i0 += kMatchThreads) {
        const uint32_t i = i0 + threadIdx.x;
        const bool ok = i < a.nS && a.cand[i] != 0xffffffffu;
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) sScan[warp] = __popc(m);
        __syncthreads();
        uint32_t off = base;
        for (int w = 0; w < warp; w++) off += sScan[w];
        if (ok) {
            const uint32_t o = off + __popc(m & ((1u << lane) - 1u));
            a.matches[2 * o] = i;
            a.matches[2 * o + 1] = a.cand[i];
            a.ratios[o] = a.candRatio[i];
        }
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kMatchThreads / 32; w++) t += sScan[w];
            sTotal = t;
        }
        __syncthreads();
        base += sTotal;
    }
    if (threadIdx.x == 0) *a.nOut = base;
}

}  // namespace pgi
