// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.  The product never links it.
//
// pgo_geom.hpp — pose algebra + Sampson scoring.
// All file:line citations are into /root/reference/src/pyposegraphbuilder/include/.
//
// PARITY STATUS: the reference has no tests/golden vectors and cannot be compiled here
// (Eigen, Sophus, OpenCV C++ absent).  Eigen/Sophus arithmetic below is restated from
// the published Eigen 3.4.0 / Sophus 22.x sources (scalar, non-vectorised evaluation
// order assumed) => "parity unpinned" for those stages; OpenCV-owned stages
// (pgo_cv.hpp) are pinned against the cv2 4.13.0 wheel (tests/golden/).
//
// Build flags must forbid FMA contraction (-ffp-contract=off) — the reference builds
// with -O3 and no -march (CMakeLists.txt:28-34) => plain SSE2 IEEE double.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace pgo {

struct Vec3 { double x, y, z; };

// Eigen::Quaterniond coefficient order is (x,y,z,w).
struct Quat { double x, y, z, w; };

// Sophus::SE3d = unit quaternion + translation.
struct SE3 {
    Quat q;
    Vec3 t;
};

inline SE3 se3Identity() { return SE3{Quat{0.0, 0.0, 0.0, 1.0}, Vec3{0.0, 0.0, 0.0}}; }

// Eigen redux order for fixed-size 4 vectors (Redux.h, redux_novec_unroller):
// (x0+x1)+(x2+x3); for 3-vectors x0+(x1+x2).
inline double sum4(double a, double b, double c, double d) { return (a + b) + (c + d); }
inline double sum3(double a, double b, double c) { return a + (b + c); }

// Eigen::QuaternionBase::normalize(): coeffs /= norm().
inline Quat quatNormalized(const Quat &q)
{
    const double n = std::sqrt(sum4(q.x * q.x, q.y * q.y, q.z * q.z, q.w * q.w));
    return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}

// Sophus::SO3Base::operator* — Hamilton product written out component-wise, then the
// SO3(quaternion) constructor normalises (Sophus 22.x so3.hpp).
inline Quat quatMul(const Quat &a, const Quat &b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

inline Vec3 cross(const Vec3 &a, const Vec3 &b)
{
    return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Sophus SO3 * point: uv = q.vec x p; uv += uv; p + w*uv + q.vec x uv.
inline Vec3 quatRotate(const Quat &q, const Vec3 &p)
{
    const Vec3 qv{q.x, q.y, q.z};
    Vec3 uv = cross(qv, p);
    uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
    const Vec3 c2 = cross(qv, uv);
    return Vec3{(p.x + q.w * uv.x) + c2.x, (p.y + q.w * uv.y) + c2.y, (p.z + q.w * uv.z) + c2.z};
}

// Sophus::SE3::operator*: SE3(so3*so3', t + so3*t').   graph_traversal.h:341-344
inline SE3 se3Mul(const SE3 &a, const SE3 &b)
{
    SE3 r;
    r.q = quatNormalized(quatMul(a.q, b.q));
    const Vec3 rt = quatRotate(a.q, b.t);
    r.t = Vec3{a.t.x + rt.x, a.t.y + rt.y, a.t.z + rt.z};
    return r;
}

// Sophus::SE3::inverse(): invR = so3.inverse(); SE3(invR, invR * (t * -1)).
inline SE3 se3Inverse(const SE3 &a)
{
    SE3 r;
    r.q = Quat{-a.q.x, -a.q.y, -a.q.z, a.q.w};
    const Vec3 nt{a.t.x * -1.0, a.t.y * -1.0, a.t.z * -1.0};
    r.t = quatRotate(r.q, nt);
    return r;
}

// Eigen::QuaternionBase::toRotationMatrix(), row-major 3x3 out.
inline void quatToRotation(const Quat &q, double R[9])
{
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// Eigen::Quaterniond(Matrix3d) — Shoemake.  R row-major.   pose_graph_builder.h:1073-1075
inline Quat rotationToQuat(const double R[9])
{
    Quat q;
    double t = sum3(R[0], R[4], R[8]);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double c[3];
        t = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        c[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
        c[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        c[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q.x = c[0]; q.y = c[1]; q.z = c[2];
    }
    return q;
}

// pose::getEssentialMatrixFromRelativePose  pose_utils.h:74-86 ; Pose ctor pose.h:46-55.
// E = [t]x * R, row-major; each entry is a 3-term Eigen product: a0*b0 + (a1*b1 + a2*b2).
inline void essentialFromPose(const SE3 &T, double E[9])
{
    double R[9];
    quatToRotation(T.q, R);
    const double tx = T.t.x, ty = T.t.y, tz = T.t.z;
    const double C[9] = {0.0, -tz, ty, tz, 0.0, -tx, -ty, tx, 0.0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = sum3(C[i * 3 + 0] * R[0 * 3 + j], C[i * 3 + 1] * R[1 * 3 + j], C[i * 3 + 2] * R[2 * 3 + j]);
}

// EssentialMatrixEvaluator::squaredSampsonDistance  graph_traversal.h:86-116 (literal order).
inline double sampsonSq(const double *c, const double E[9])
{
    const double x1 = c[0], y1 = c[1], x2 = c[2], y2 = c[3];
    const double e11 = E[0], e12 = E[1], e13 = E[2], e21 = E[3], e22 = E[4], e23 = E[5],
                 e31 = E[6], e32 = E[7], e33 = E[8];
    const double rxc = e11 * x2 + e21 * y2 + e31;
    const double ryc = e12 * x2 + e22 * y2 + e32;
    const double rwc = e13 * x2 + e23 * y2 + e33;
    const double r = (x1 * rxc + y1 * ryc + rwc);
    const double rx = e11 * x1 + e12 * y1 + e13;
    const double ry = e21 * x1 + e22 * y1 + e23;
    return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
}

// InTraversalPoseTester::test  graph_traversal.h:194-233 — early exit at minInliers.
inline bool inTraversalTest(const double *corr, size_t n, const SE3 &pose, double thr /*=1.5*thr_norm*/,
                            size_t minInliers, size_t &inlierNumber)
{
    const double sqThr = thr * thr;  // :184
    inlierNumber = 0;
    double E[9];
    essentialFromPose(pose, E);  // :200-202
    for (size_t i = 0; i < n; ++i) {
        if (sampsonSq(corr + 4 * i, E) < sqThr) {
            ++inlierNumber;
            if (inlierNumber >= minInliers) return true;
        }
    }
    return false;
}

// EssentialMatrixEvaluator::getInliers  graph_traversal.h:136-168 — NOTE :164 compares the
// squared residual against the UN-squared threshold; reproduced.
inline void getInliers(const double *corr, size_t n, const double E[9], double thr, std::vector<size_t> &inliers)
{
    inliers.clear();
    for (size_t i = 0; i < n; ++i)
        if (sampsonSq(corr + 4 * i, E) < thr) inliers.push_back(i);
}

// PoseGraphBuilder::createCorrespondenceMatrix  pose_graph_builder.h:864-938.
// Keypoints are cv::KeyPoint.pt (FP32 pixels).  Destination points are normalised with the
// SOURCE camera (:908-912).
inline void createCorrespondenceMatrix(const float *kpSrc, const float *kpDst, const uint32_t *matches /*n x 2*/,
                                       size_t n, double fx, double fy, double cx, double cy, double thrPx,
                                       double *corr /*n x 4*/, double &thrNorm)
{
    for (size_t i = 0; i < n; ++i) {
        const size_t s = matches[2 * i], d = matches[2 * i + 1];
        corr[4 * i + 0] = ((double)kpSrc[2 * s + 0] - cx) / fx;
        corr[4 * i + 1] = ((double)kpSrc[2 * s + 1] - cy) / fy;
        corr[4 * i + 2] = ((double)kpDst[2 * d + 0] - cx) / fx;
        corr[4 * i + 3] = ((double)kpDst[2 * d + 1] - cy) / fy;
    }
    const double normalizer = (fx + fy + fx + fy) / 4.0;  // :934-935
    thrNorm = thrPx / normalizer;
}

}  // namespace pgo
