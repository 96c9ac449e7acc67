// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_eigen.hpp — E -> (R, t): SURVEY §8a rows a9-a11.
//   pose::decomposeEssentialMatrix     pose_utils.h:144-169
//   pose::getPoseFromEssentialMatrix   pose_utils.h:172-252
//   pose::linearTriangulation          pose_utils.h:491-506
// The reference leans on Eigen::JacobiSVD<Matrix3d/Matrix4d> (un-vendored, version un-pinned:
// CMakeLists.txt:40 `find_package(Eigen3 REQUIRED)`).  The two-sided Jacobi below restates Eigen
// 3.4.0 src/SVD/JacobiSVD.h + src/Jacobi/Jacobi.h + src/misc/RealSvd2x2.h operation for operation
// (square real input => no QR preconditioner).  PARITY UNPINNED: nothing executable exists here to
// confirm Eigen's bit-level behaviour; golden vectors in tests/golden/ freeze
// THIS restatement, and mathematical property tests check it (orthogonality, U S V^T = A).
// Fixed-size dot products follow Eigen's scalar redux order: 3 terms a0+(a1+a2), 4 terms (a0+a1)+(a2+a3).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "pgo_geom.hpp"

namespace pgo {
namespace eig {

struct Rot { double c, s; };  // Eigen::JacobiRotation

// JacobiRotation::makeJacobi(x, y, z)   (Jacobi.h)
inline Rot makeJacobi(double x, double y, double z)
{
    const double deno = 2.0 * std::abs(y);
    if (deno < DBL_MIN) return Rot{1.0, 0.0};
    const double tau = (x - z) / deno;
    const double w = std::sqrt(tau * tau + 1.0);
    double t;
    if (tau > 0.0)
        t = 1.0 / (tau + w);
    else
        t = 1.0 / (tau - w);
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / std::sqrt(t * t + 1.0);
    Rot r;
    r.s = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
    r.c = n;
    return r;
}

// real_2x2_jacobi_svd  (RealSvd2x2.h)
inline void real2x2JacobiSvd(double mpp, double mpq, double mqp, double mqq, Rot &jLeft, Rot &jRight)
{
    double m00 = mpp, m01 = mpq, m10 = mqp, m11 = mqq;
    Rot rot1;
    const double t = m00 + m11;
    const double d = m10 - m01;
    if (std::abs(d) < DBL_MIN) {
        rot1.s = 0.0;
        rot1.c = 1.0;
    } else {
        const double u = t / d;
        const double tmp = std::sqrt(1.0 + u * u);
        rot1.s = 1.0 / tmp;
        rot1.c = u / tmp;
    }
    // m.applyOnTheLeft(0,1,rot1): row0' = c row0 + s row1 ; row1' = -s row0 + c row1
    if (!(rot1.c == 1.0 && rot1.s == 0.0)) {
        const double a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
        const double a10 = -rot1.s * m00 + rot1.c * m10, a11 = -rot1.s * m01 + rot1.c * m11;
        m00 = a00; m01 = a01; m10 = a10; m11 = a11;
    }
    jRight = makeJacobi(m00, m01, m11);
    // *j_left = rot1 * j_right->transpose();  transpose = (c, -s); product (c1c2 - s1s2, c1s2 + s1c2)
    const Rot tr{jRight.c, -jRight.s};
    jLeft.c = rot1.c * tr.c - rot1.s * tr.s;
    jLeft.s = rot1.c * tr.s + rot1.s * tr.c;
}

// Eigen::JacobiSVD<Matrix<double,N,N>> (column-major semantics restated on row-major storage
// M[r*N+c]).  U and V are optional (nullptr = not computed, as with ComputeFullV only).
template <int N>
inline void jacobiSvd(const double *Ain, double *U, double *V, double *S)
{
    double W[N * N];
    double scale = 0.0;
    for (int i = 0; i < N * N; i++) {
        const double a = std::abs(Ain[i]);
        if (a > scale || a != a) scale = a;  // maxCoeff<PropagateNaN>
    }
    if (!std::isfinite(scale)) {
        // Eigen: m_info = InvalidInput, matrices left unset; we emit NaNs.
        for (int i = 0; i < N * N; i++) {
            if (U) U[i] = NAN;
            if (V) V[i] = NAN;
        }
        for (int i = 0; i < N; i++) S[i] = NAN;
        return;
    }
    if (scale == 0.0) scale = 1.0;
    for (int i = 0; i < N * N; i++) W[i] = Ain[i] / scale;
    if (U)
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) U[i * N + j] = i == j ? 1.0 : 0.0;
    if (V)
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) V[i * N + j] = i == j ? 1.0 : 0.0;

    const double precision = 2.0 * DBL_EPSILON;
    const double considerAsZero = DBL_MIN;
    double maxDiagEntry = 0.0;
    for (int i = 0; i < N; i++) maxDiagEntry = std::max(maxDiagEntry, std::abs(W[i * N + i]));

    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < N; ++p)
            for (int q = 0; q < p; ++q) {
                const double threshold = std::max(considerAsZero, precision * maxDiagEntry);
                if (std::abs(W[p * N + q]) > threshold || std::abs(W[q * N + p]) > threshold) {
                    finished = false;
                    Rot jl, jr;
                    real2x2JacobiSvd(W[p * N + p], W[p * N + q], W[q * N + p], W[q * N + q], jl, jr);
                    // W.applyOnTheLeft(p,q,jl)
                    if (!(jl.c == 1.0 && jl.s == 0.0))
                        for (int k = 0; k < N; k++) {
                            const double xi = W[p * N + k], yi = W[q * N + k];
                            W[p * N + k] = jl.c * xi + jl.s * yi;
                            W[q * N + k] = -jl.s * xi + jl.c * yi;
                        }
                    // U.applyOnTheRight(p,q,jl.transpose()) -> apply_rotation_in_the_plane(col p, col q, jl)
                    if (U && !(jl.c == 1.0 && jl.s == 0.0))
                        for (int k = 0; k < N; k++) {
                            const double xi = U[k * N + p], yi = U[k * N + q];
                            U[k * N + p] = jl.c * xi + jl.s * yi;
                            U[k * N + q] = -jl.s * xi + jl.c * yi;
                        }
                    // W.applyOnTheRight(p,q,jr) -> apply_rotation_in_the_plane(col p, col q, jr.transpose())
                    if (!(jr.c == 1.0 && -jr.s == 0.0))
                        for (int k = 0; k < N; k++) {
                            const double xi = W[k * N + p], yi = W[k * N + q];
                            W[k * N + p] = jr.c * xi + (-jr.s) * yi;
                            W[k * N + q] = -(-jr.s) * xi + jr.c * yi;
                        }
                    if (V && !(jr.c == 1.0 && -jr.s == 0.0))
                        for (int k = 0; k < N; k++) {
                            const double xi = V[k * N + p], yi = V[k * N + q];
                            V[k * N + p] = jr.c * xi + (-jr.s) * yi;
                            V[k * N + q] = -(-jr.s) * xi + jr.c * yi;
                        }
                    maxDiagEntry = std::max(maxDiagEntry, std::max(std::abs(W[p * N + p]), std::abs(W[q * N + q])));
                }
            }
    }
    for (int i = 0; i < N; ++i) {
        const double a = W[i * N + i];
        S[i] = std::abs(a);
        if (U && a < 0.0)
            for (int k = 0; k < N; k++) U[k * N + i] = -U[k * N + i];
    }
    for (int i = 0; i < N; i++) S[i] *= scale;
    for (int i = 0; i < N; i++) {
        int pos = 0;
        double mx = S[i];
        for (int k = 1; k < N - i; k++)
            if (S[i + k] > mx) { mx = S[i + k]; pos = k; }
        if (mx == 0.0) break;
        if (pos) {
            pos += i;
            std::swap(S[i], S[pos]);
            if (U) for (int k = 0; k < N; k++) std::swap(U[k * N + pos], U[k * N + i]);
            if (V) for (int k = 0; k < N; k++) std::swap(V[k * N + pos], V[k * N + i]);
        }
    }
}

inline double det3(const double *M)
{
    // Eigen bruteforce_det3_helper order (Determinant.h)
    const double a = M[0] * (M[4] * M[8] - M[5] * M[7]);
    const double b = M[1] * (M[3] * M[8] - M[5] * M[6]);
    const double c = M[2] * (M[3] * M[7] - M[4] * M[6]);
    return a - b + c;
}

// decomposeEssentialMatrix  pose_utils.h:144-169.  All matrices row-major.
inline void decomposeEssentialMatrix(const double E[9], double R1[9], double R2[9], double t[3])
{
    double U[9], V[9], S[3];
    jacobiSvd<3>(E, U, V, S);
    if (det3(U) < 0) for (int k = 0; k < 3; k++) U[k * 3 + 2] *= -1.0;
    if (det3(V) < 0) for (int k = 0; k < 3; k++) V[k * 3 + 2] *= -1.0;
    // d = [0 1 0; -1 0 0; 0 0 1]:  U*d = [-U.col(1), U.col(0), U.col(2)] (exact),
    // U*d^T = [U.col(1), -U.col(0), U.col(2)]
    double Ud[9], Udt[9];
    for (int i = 0; i < 3; i++) {
        Ud[i * 3 + 0] = -U[i * 3 + 1]; Ud[i * 3 + 1] = U[i * 3 + 0]; Ud[i * 3 + 2] = U[i * 3 + 2];
        Udt[i * 3 + 0] = U[i * 3 + 1]; Udt[i * 3 + 1] = -U[i * 3 + 0]; Udt[i * 3 + 2] = U[i * 3 + 2];
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            R1[i * 3 + j] = sum3(Ud[i * 3 + 0] * V[j * 3 + 0], Ud[i * 3 + 1] * V[j * 3 + 1], Ud[i * 3 + 2] * V[j * 3 + 2]);
            R2[i * 3 + j] = sum3(Udt[i * 3 + 0] * V[j * 3 + 0], Udt[i * 3 + 1] * V[j * 3 + 1], Udt[i * 3 + 2] * V[j * 3 + 2]);
        }
    // translation = U.col(2).normalized()
    const double u0 = U[2], u1 = U[5], u2 = U[8];
    const double z = sum3(u0 * u0, u1 * u1, u2 * u2);
    if (z > 0.0) {
        const double n = std::sqrt(z);
        t[0] = u0 / n; t[1] = u1 / n; t[2] = u2 / n;
    } else {
        t[0] = u0; t[1] = u1; t[2] = u2;
    }
}

// One (candidate, correspondence) evaluation of the loop body pose_utils.h:203-231.
// P2 = [R | tc] row-major 3x4.  Returns false when a depth test fails (the `continue`s).
inline bool triangulateAndScore(const double P2[12], const double *c, double &error)
{
    // linearTriangulation pose_utils.h:497-504 with proj_1 = [I|0]
    double D[16];
    const double P1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (int k = 0; k < 4; k++) {
        D[0 * 4 + k] = c[0] * P1[8 + k] - P1[0 + k];
        D[1 * 4 + k] = c[1] * P1[8 + k] - P1[4 + k];
        D[2 * 4 + k] = c[2] * P2[8 + k] - P2[0 + k];
        D[3 * 4 + k] = c[3] * P2[8 + k] - P2[4 + k];
    }
    double V[16], S[4];
    jacobiSvd<4>(D, nullptr, V, S);
    const double X[4] = {V[0 * 4 + 3], V[1 * 4 + 3], V[2 * 4 + 3], V[3 * 4 + 3]};
    double p1[3], p2[3];
    for (int r = 0; r < 3; r++) {
        p1[r] = sum4(P1[r * 4 + 0] * X[0], P1[r * 4 + 1] * X[1], P1[r * 4 + 2] * X[2], P1[r * 4 + 3] * X[3]);
    }
    if (p1[2] < 0) return false;
    for (int r = 0; r < 3; r++) {
        p2[r] = sum4(P2[r * 4 + 0] * X[0], P2[r * 4 + 1] * X[1], P2[r * 4 + 2] * X[2], P2[r * 4 + 3] * X[3]);
    }
    if (p2[2] < 0) return false;
    const double a0 = p1[0] / p1[2] - c[0], a1 = p1[1] / p1[2] - c[1];
    const double b0 = p2[0] / p2[2] - c[2], b1 = p2[1] / p2[2] - c[3];
    error = (a0 * a0 + a1 * a1) + (b0 * b0 + b1 * b1);
    return true;
}

// getPoseFromEssentialMatrix  pose_utils.h:172-252.  votes[4] optional out.
inline int getPoseFromEssentialMatrix(const double E[9], const double *corr, size_t n, double R[9], double t[3],
                                      size_t votesOut[4] = nullptr)
{
    double R1[9], R2[9], tr[3];
    decomposeEssentialMatrix(E, R1, R2, tr);
    const double *rots[4] = {R1, R1, R2, R2};
    std::vector<double> best(n, DBL_MAX);
    std::vector<int> bestPose(n, 5);
    for (int i = 0; i < 4; i++) {
        const double sgn = (i % 2 ? -1 : 1);
        double P2[12];
        for (int r = 0; r < 3; r++) {
            P2[r * 4 + 0] = rots[i][r * 3 + 0];
            P2[r * 4 + 1] = rots[i][r * 3 + 1];
            P2[r * 4 + 2] = rots[i][r * 3 + 2];
            P2[r * 4 + 3] = sgn * tr[r];
        }
        for (size_t p = 0; p < n; ++p) {
            double err;
            if (!triangulateAndScore(P2, corr + 4 * p, err)) continue;
            if (err < best[p]) {
                best[p] = err;
                bestPose[p] = i;
            }
        }
    }
    size_t votes[4] = {0, 0, 0, 0};
    for (size_t p = 0; p < n; ++p)
        if (bestPose[p] < 5) ++votes[bestPose[p]];
    int maxIdx = 0;
    for (int i = 1; i < 4; i++)
        if (votes[i] > votes[maxIdx]) maxIdx = i;  // std::max_element: first maximum
    for (int k = 0; k < 9; k++) R[k] = rots[maxIdx][k];
    const double sgn = (maxIdx % 2 ? -1 : 1);
    for (int k = 0; k < 3; k++) t[k] = sgn * tr[k];
    if (votesOut) for (int i = 0; i < 4; i++) votesOut[i] = votes[i];
    return (int)votes[maxIdx];
}

}  // namespace eig
}  // namespace pgo
