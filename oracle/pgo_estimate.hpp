// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_estimate.hpp — PoseGraphBuilder::estimatePose  pose_graph_builder.h:940-1078.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "pgo_cv.hpp"
#include "pgo_eigen.hpp"
#include "pgo_fallback.hpp"
#include "pgo_geom.hpp"

namespace pgo {

struct EstimateResult {
    bool success = false;
    int branch = 0;            // 0 none, 1 path hypothesis accepted, 2 fallback accepted
    size_t inlierNumber = 0;   // pose_graph_builder.h:1022 / :1047
    size_t pathInliers = 0;    // |getInliers| of the last guess (0 if no guess)
    double E[9] = {0};         // the E handed to the decomposition (row-major, as OpenCV returns it)
    SE3 pose = se3Identity();
    size_t votes[4] = {0, 0, 0, 0};
    int fallbackIters = 0;
    int fallbackModels = 0;
};

inline EstimateResult estimatePose(const double *corr, size_t n, double thrNorm, size_t minInliers,
                                   const SE3 *guesses, size_t nGuesses, std::vector<uint8_t> &inlierMask)
{
    EstimateResult res;
    const double truncThr = (3.0 / 2.0) * thrNorm;  // :963-964
    std::vector<size_t> inliers;
    std::vector<double> tmp;
    bool success = false;
    double cvE[90];
    bool haveE = false;

    for (size_t g = 0; g < nGuesses; ++g) {  // :974
        double E[9];
        essentialFromPose(guesses[g], E);         // :980-982
        getInliers(corr, n, E, truncThr, inliers);  // :985-989 (un-squared threshold, §0.7)
        const size_t k = inliers.size();
        tmp.resize(k * 4);
        inlierMask.resize(n, 0);  // :1000 (resize, not clear)
        for (size_t i = 0; i < k; ++i) {
            std::memcpy(&tmp[4 * i], corr + 4 * inliers[i], 4 * sizeof(double));
            inlierMask[inliers[i]] = 1;
        }
        std::vector<uint8_t> tmpMask(k, 0);  // :1012
        const int nm = cvx::findEssentialMatRansacInf(tmp.data(), (int)k, cvE, tmpMask.data());  // :1013-1020
        haveE = nm > 0;
        size_t cnt = 0;
        for (size_t i = 0; i < k; ++i) cnt += tmpMask[i];
        res.inlierNumber = cnt;          // :1022-1023
        res.pathInliers = k;
        success = cnt >= minInliers;     // :1028
        if (success) res.branch = 1;
    }

    if (!success) {  // :1031
        inlierMask.resize(n, 0);
        std::vector<uint8_t> m(n, 0);
        fb::FallbackResult fr = fb::runFallback(corr, (int)n, thrNorm, m.data());  // :1037-1044
        res.fallbackIters = fr.iterations;
        res.fallbackModels = fr.models;
        for (size_t i = 0; i < n; ++i) inlierMask[i] = m[i];
        res.inlierNumber = (size_t)fr.inliers;  // :1047-1048
        if (res.inlierNumber < minInliers || !fr.ok) {  // :1053
            res.branch = 0;
            return res;
        }
        std::memcpy(cvE, fr.E, 9 * sizeof(double));
        haveE = true;
        res.branch = 2;
    }
    if (!haveE) return res;
    std::memcpy(res.E, cvE, 9 * sizeof(double));  // only the first 3x3 is used (:1057)

    double R[9], t[3];
    eig::getPoseFromEssentialMatrix(res.E, corr, n, R, t, res.votes);  // :1062-1066
    bool nan = false;
    for (int k = 0; k < 9; k++) nan |= (R[k] != R[k]);
    for (int k = 0; k < 3; k++) nan |= (t[k] != t[k]);
    if (nan) {  // :1069-1070
        res.branch = 0;
        return res;
    }
    // Sophus::SE3d(Eigen::Quaterniond(rotation), translation): SO3(quaternion) normalises  :1073-1075
    res.pose.q = quatNormalized(rotationToQuat(R));
    res.pose.t = Vec3{t[0], t[1], t[2]};
    res.success = true;
    return res;
}

}  // namespace pgo
