// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_matcher.hpp — restatement of the epipolar-hashing guided matcher:
//   HashingBasedMatcherWithPose<false, 45>::match          matcher.h:199-405
//   PoseGraphBuilder::guidedMatching (caller, top-K cut)    pose_graph_builder.h:717-783
// Quirks kept: the destination keypoints are binned by the angle of their epipolar line's normal in the SOURCE image
// (F^T p2), the angular range comes from the DESTINATION image's corners (:245-268) or is 0 - 180 = -180 when the epipole
// lies inside the source image (:236-241); the symmetric epipolar distance is tested against 0.75^2 px^2 whatever the
// threshold argument (:362); descriptor differences are taken in float and accumulated in double (:371-375); Lowe ratio
// corrected by the candidate count (:386-399); "too good to be true" ratios are dropped (:402); the caller attaches
// descriptorDistances[0] to EVERY match when there are at most kMaximumPointNumberForEpipolarHashing of them (:777-779).
// Eigen pieces (3x3 inverse by cofactors, 3-term products, JacobiSVD) follow pgo_eigen.hpp's conventions.
// PARITY UNPINNED: the reference has no tests or fixtures for the matcher and cannot be compiled here; what
// pins this file is its closeness to matcher.h and behavioural tests (tests/test_oracle_matcher.py: matches are true
// correspondences within 0.75 px of their epipolar lines, selection rule).  atan2 is the host libm's here (the device
// evaluates the correctly rounded value, which glibc returns for all but ~2.5e-4 of arguments: tests/test_atan2_cr.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <tuple>
#include <utility>
#include <vector>

#include "pgo_eigen.hpp"

namespace pgo {
namespace matcher {

// Eigen::internal::compute_inverse_size3 (cofactors of column 0 first, det = their dot product with column 0).
inline void inverse3(const double *m, double *inv)
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
inline void mul3(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + (A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j]);
}

struct Prepared {  // everything match() derives from the pose before it looks at a keypoint
    double F[9];
    double epipole[2];
    double minAngle, angularRange;
    int binNumber;
};
inline double lineAngle(double ny, double nx)  // :256-258, :288-290, :327-329
{
    constexpr double kRadianToDegree = 180.0 / M_PI;
    double angle = kRadianToDegree * std::atan2(ny, nx) + 180.0;
    if (angle > 180) angle -= 180;
    return angle;
}
inline int binOf(const Prepared &P, double angle)
{
    angle = (P.binNumber - 1) * (angle - P.minAngle) / P.angularRange;
    return std::min(std::max(0, static_cast<int>(std::round(angle))), P.binNumber - 1);
}
inline Prepared prepare(const double E[9], const double Ks[9], const double Kd[9], int wS, int hS, int wD, int hD, int binNumber)
{
    Prepared P;
    double KdInv[9], KsInv[9], KdInvT[9], T[9];
    inverse3(Kd, KdInv);
    inverse3(Ks, KsInv);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) KdInvT[i * 3 + j] = KdInv[j * 3 + i];
    mul3(KdInvT, E, T);
    mul3(T, KsInv, P.F);  // :220-221
    double V[9], S[3];
    eig::jacobiSvd<3>(P.F, nullptr, V, S);  // :224-226
    const double ez = V[2 * 3 + 2];
    P.epipole[0] = V[0 * 3 + 2] / ez;  // :228-230
    P.epipole[1] = V[1 * 3 + 2] / ez;
    const bool inImage = P.epipole[0] >= 0 && P.epipole[0] < wS && P.epipole[1] >= 0 && P.epipole[1] < hS;  // :232-234
    double minAngle = 180, maxAngle = 0;
    if (!inImage) {
        const double corners[8] = {0, 0, (double)wD, 0, (double)wD, (double)hD, 0, (double)hD};
        for (int c = 0; c < 8; c += 2) {
            const double x = corners[c], y = corners[c + 1];
            const double nx = P.F[0] * x + P.F[3] * y + P.F[6];
            const double ny = P.F[1] * x + P.F[4] * y + P.F[7];
            const double angle = lineAngle(ny, nx);
            minAngle = std::min(minAngle, angle);
            maxAngle = std::max(maxAngle, angle);
        }
    }
    P.minAngle = minAngle;
    P.angularRange = maxAngle - minAngle;  // :271
    P.binNumber = binNumber <= 0 ? static_cast<int>(P.angularRange) : binNumber;  // :274-277
    return P;
}

// match(): matches (source index, destination index) in source order + the adapted squared ratio of each.
inline void match(const float *kpS, size_t nS, const float *dS, const float *kpD, size_t nD, const float *dD, int dim,
                  const Prepared &P, std::vector<std::pair<uint32_t, uint32_t>> &matches, std::vector<double> &ratios)
{
    matches.clear();
    ratios.clear();
    std::vector<std::vector<uint32_t>> bins(std::max(P.binNumber, 0));
    for (size_t i = 0; i < nD; i++) {  // :285-299
        const double x = kpD[2 * i], y = kpD[2 * i + 1];
        const double nx = P.F[0] * x + P.F[3] * y + P.F[6];
        const double ny = P.F[1] * x + P.F[4] * y + P.F[7];
        bins[binOf(P, lineAngle(ny, nx))].push_back((uint32_t)i);
    }
    const double e11 = P.F[0], e12 = P.F[1], e13 = P.F[2], e21 = P.F[3], e22 = P.F[4], e23 = P.F[5], e31 = P.F[6], e32 = P.F[7], e33 = P.F[8];
    for (size_t i = 0; i < nS; i++) {  // :316-409
        const double x1 = kpS[2 * i], y1 = kpS[2 * i + 1];
        const double vx = x1 - P.epipole[0], vy = y1 - P.epipole[1];
        const int bin = binOf(P, lineAngle(vx, -vy));  // normal (nx, ny) = (-vy, vx); atan2(ny, nx)
        double second = std::numeric_limits<double>::max(), best = std::numeric_limits<double>::max();
        int bestIndex = -1, countSnn = 0;
        for (uint32_t nb : bins[bin]) {
            const double x2 = kpD[2 * nb], y2 = kpD[2 * nb + 1];
            const double rxc = e11 * x2 + e21 * y2 + e31;
            const double ryc = e12 * x2 + e22 * y2 + e32;
            const double rwc = e13 * x2 + e23 * y2 + e33;
            const double r = (x1 * rxc + y1 * ryc + rwc);
            const double rx = e11 * x1 + e12 * y1 + e13;
            const double ry = e21 * x1 + e22 * y1 + e23;
            const double a1 = rxc * rxc + ryc * ryc;
            const double b1 = rx * rx + ry * ry;
            const double d2 = r * r * (a1 + b1) / (a1 * b1);  // squared symmetric epipolar distance :360
            if (d2 >= 0.75 * 0.75) continue;                   // :362
            countSnn += 1;
            double dd = 0;
            for (int m = 0; m < dim; m++) {
                const double dist = dS[i * dim + m] - dD[(size_t)nb * dim + m];  // float subtraction, then widened
                dd += dist * dist;
            }
            if (dd < best) { second = best; best = dd; bestIndex = (int)nb; }
        }
        double corr = 1.0;  // :386-399
        if (countSnn < 20) corr = 0.65 * 0.65;
        if (countSnn < 10) corr = 0.6 * 0.6;
        if (countSnn < 5) corr = 0.5 * 0.5;
        if (countSnn < 3) corr = 0.25 * 0.25;
        const double adapted = (best / second) / corr;
        if (adapted < 0.00001) continue;  // :402
        if (bestIndex > -1 && ((adapted < 0.8 * 0.8) || (countSnn == 1))) {
            matches.emplace_back((uint32_t)i, (uint32_t)bestIndex);
            ratios.push_back(adapted);
        }
    }
}

// PoseGraphBuilder::guidedMatching's selection (pose_graph_builder.h:760-782).
inline void selectMatches(const std::vector<std::pair<uint32_t, uint32_t>> &matches, const std::vector<double> &ratios,
                          size_t maxPoints, std::vector<std::tuple<uint32_t, uint32_t, double>> &out)
{
    out.clear();
    if (matches.size() > maxPoints) {
        std::priority_queue<std::pair<double, size_t>, std::vector<std::pair<double, size_t>>, std::greater<std::pair<double, size_t>>> q;
        for (size_t i = 0; i < matches.size(); ++i) q.emplace(ratios[i], i);
        while (!q.empty() && out.size() < maxPoints) {
            const auto &m = matches[q.top().second];
            out.emplace_back(m.first, m.second, q.top().first);
            q.pop();
        }
    } else
        for (const auto &m : matches) out.emplace_back(m.first, m.second, ratios[0]);  // descriptorDistances[0] (sic)
}

}  // namespace matcher
}  // namespace pgo
