// pgi_math.cuh — FP64 device arithmetic of the hypothesis-verification path (sm_100a).
//
// Everything here is decision-bearing: inlier counts feed edge scores that steer later A* searches
// (SURVEY §7 hard part 4) and the E->(R,t) vote is decided at ulp level.  The
// translation unit is therefore compiled with -fmad=false (no FMA contraction; the reference host
// build has none, CMakeLists.txt:28-34) and every expression keeps the reference's evaluation order.
// IEEE double div/sqrt are exact-rounded on the device by default.
//
// Reference call sites (into /root/reference/src/pyposegraphbuilder/include/):
//   squaredSampsonDistance      graph_traversal.h:86-116
//   getEssentialMatrixFromRelativePose  pose_utils.h:74-86
//   cv::findEssentialMat(RANSAC) five-point kernel   called at pose_graph_builder.h:1013-1020
//   decomposeEssentialMatrix / getPoseFromEssentialMatrix / linearTriangulation
//                               pose_utils.h:144-169, :172-252, :491-506
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace pgi {

#define PGI_DEV __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Pose algebra
// ---------------------------------------------------------------------------------------------
PGI_DEV double sum3(double a, double b, double c) { return a + (b + c); }
PGI_DEV double sum4(double a, double b, double c, double d) { return (a + b) + (c + d); }

// Eigen::Quaternion::toRotationMatrix, row-major.
PGI_DEV void quatToRotation(const double q[4] /*x y z w*/, double R[9])
{
    const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// E = [t]x R  (pose_utils.h:74-86), 3-term products in Eigen's redux order.
PGI_DEV void essentialFromPose(const double qt[7], double E[9])
{
    double R[9];
    quatToRotation(qt, R);
    const double tx = qt[4], ty = qt[5], tz = qt[6];
    const double C[9] = {0.0, -tz, ty, tz, 0.0, -tx, -ty, tx, 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = sum3(C[i * 3 + 0] * R[0 * 3 + j], C[i * 3 + 1] * R[1 * 3 + j], C[i * 3 + 2] * R[2 * 3 + j]);
}

// graph_traversal.h:107-115, literal evaluation order.
PGI_DEV double sampsonSq(double x1, double y1, double x2, double y2, const double E[9])
{
    const double rxc = E[0] * x2 + E[3] * y2 + E[6];
    const double ryc = E[1] * x2 + E[4] * y2 + E[7];
    const double rwc = E[2] * x2 + E[5] * y2 + E[8];
    const double r = (x1 * rxc + y1 * ryc + rwc);
    const double rx = E[0] * x1 + E[1] * y1 + E[2];
    const double ry = E[3] * x1 + E[4] * y1 + E[5];
    return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
}

// Eigen::Quaterniond(Matrix3d) (Shoemake) followed by Sophus' normalisation (pose_graph_builder.h:1073-1075).
PGI_DEV void rotationToUnitQuat(const double R[9], double q[4])
{
    double t = sum3(R[0], R[4], R[8]);
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
    const double n = sqrt(sum4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]));
    q[0] = q[0] / n; q[1] = q[1] / n; q[2] = q[2] / n; q[3] = q[3] / n;
}

// ---------------------------------------------------------------------------------------------
// OpenCV-owned arithmetic (core/src/lapack.cpp, mathfuncs.cpp; calib3d/src/five-point.cpp)
// ---------------------------------------------------------------------------------------------
struct CvRng {  // cv::RNG multiply-with-carry
    uint64_t state;
    PGI_DEV explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffULL) {}
    PGI_DEV uint32_t next()
    {
        state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
        return (uint32_t)state;
    }
    PGI_DEV int uniform(int a, int b) { return a == b ? a : (int)(next() % (uint32_t)(b - a) + a); }
};

PGI_DEV double cvHypot(double a, double b)  // the file-local hypot of lapack.cpp
{
    a = fabs(a);
    b = fabs(b);
    if (a > b) {
        b /= a;
        return a * sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * sqrt(1 + a * a);
    }
    return 0;
}

// JacobiSVDImpl_<double> on the N rows (length M) of At, optional Vt (N x N), completion of rows N..N1-1.
template <int M, int N, int N1, bool WITH_VT>
__device__ void cvJacobiSVD(double *At /*N1 x M*/, double *W /*N*/, double *Vt /*N x N or null*/)
{
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    const int max_iter = M > 30 ? M : 30;
    double c, s, sd;
    for (int i = 0; i < N; i++) {
        sd = 0;
        for (int k = 0; k < M; k++) {
            const double t = At[i * M + k];
            sd += t * t;
        }
        W[i] = sd;
        if (WITH_VT) {
            for (int k = 0; k < N; k++) Vt[i * N + k] = 0;
            Vt[i * N + i] = 1;
        }
    }
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < N - 1; i++)
            for (int j = i + 1; j < N; j++) {
                double *Ai = At + i * M, *Aj = At + j * M;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < M; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cvHypot(p, beta);
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < M; k++) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                if (WITH_VT) {
                    double *Vi = Vt + i * N, *Vj = Vt + j * N;
                    for (int k = 0; k < N; k++) {
                        const double t0 = c * Vi[k] + s * Vj[k];
                        const double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; i++) {
        sd = 0;
        for (int k = 0; k < M; k++) {
            const double t = At[i * M + k];
            sd += t * t;
        }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; i++) {
        int j = i;
        for (int k = i + 1; k < N; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < M; k++) { const double t = At[i * M + k]; At[i * M + k] = At[j * M + k]; At[j * M + k] = t; }
            if (WITH_VT)
                for (int k = 0; k < N; k++) { const double t = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = t; }
        }
    }
    CvRng rng(0x12345678);
    for (int i = 0; i < N1; i++) {
        sd = i < N ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / M;
            for (int k = 0; k < M; k++) At[i * M + k] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0;
                    for (int k = 0; k < M; k++) sd += At[i * M + k] * At[j * M + k];
                    double asum = 0;
                    for (int k = 0; k < M; k++) {
                        const double t = At[i * M + k] - sd * At[j * M + k];
                        At[i * M + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < M; k++) At[i * M + k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < M; k++) {
                const double t = At[i * M + k];
                sd += t * t;
            }
            sd = sqrt(sd);
        }
        s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < M; k++) At[i * M + k] *= s;
    }
}

// cv::SVD::solveZ on a 3x3.
__device__ inline void cvSolveZ3(const double Bz[9], double out[3])
{
    double At[9], W[3], Vt[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) At[i * 3 + j] = Bz[j * 3 + i];
    cvJacobiSVD<3, 3, 3, true>(At, W, Vt);
    out[0] = Vt[6]; out[1] = Vt[7]; out[2] = Vt[8];
}

// LUImpl with a 10x10 identity right-hand side (cv::invert, DECOMP_LU).  A destroyed.  false if singular.
__device__ inline bool cvInvert10(double *A /*10x10*/, double *b /*10x10 out*/)
{
    const int m = 10, n = 10;
    const double eps = DBL_EPSILON * 100;
    for (int i = 0; i < 100; i++) b[i] = 0.0;
    for (int i = 0; i < 10; i++) b[i * 11] = 1.0;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabs(A[j * m + i]) > fabs(A[k * m + i])) k = j;
        if (fabs(A[k * m + i]) < eps) {
            for (int q = 0; q < 100; q++) b[q] = 0.0;
            return false;
        }
        if (k != i) {
            for (int j = i; j < m; j++) { const double t = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = t; }
            for (int j = 0; j < n; j++) { const double t = b[i * n + j]; b[i * n + j] = b[k * n + j]; b[k * n + j] = t; }
        }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            const double alpha = A[j * m + i] * d;
            for (int q = i + 1; q < m; q++) A[j * m + q] += alpha * A[i * m + q];
            for (int q = 0; q < n; q++) b[j * n + q] += alpha * b[i * n + q];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            double s = b[i * n + j];
            for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * n + j];
            b[i * n + j] = s / A[i * m + i];
        }
    return true;
}

struct Cx { double re, im; };
PGI_DEV Cx cmul(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
PGI_DEV Cx cdiv(Cx a, Cx b)
{
    const double t = 1. / (b.re * b.re + b.im * b.im);
    return Cx{(a.re * b.re + a.im * b.im) * t, (-a.re * b.im + a.im * b.re) * t};
}

// One Gauss-Seidel Durand-Kerner sweep of cv::solvePoly on register-resident roots; returns max |correction|^2.
template <int NDEG>
__device__ __forceinline__ double dkSweep(const double (&co)[NDEG + 1], Cx (&roots)[NDEG])
{
    double maxDiffSq = 0;
#pragma unroll
    for (int i = 0; i < NDEG; i++) {
        const Cx p = roots[i];
        Cx num{co[NDEG], 0.0}, denom{co[NDEG], 0.0};
#pragma unroll
        for (int j = 0; j < NDEG; j++) {
            num = cmul(num, p);
            num.re = num.re + co[NDEG - j - 1];
            num.im = num.im + 0.0;
            if (j != i) {
                const Cx d{p.re - roots[j].re, p.im - roots[j].im};
                if (d.re != 0 || d.im != 0) denom = cmul(denom, d);
            }
        }
        num = cdiv(num, denom);
        roots[i].re = p.re - num.re;
        roots[i].im = p.im - num.im;
        const double m = num.re * num.re + num.im * num.im;
        maxDiffSq = maxDiffSq < m ? m : maxDiffSq;
    }
    return maxDiffSq;
}

template <int NDEG>
__device__ __forceinline__ bool sameRoots(const Cx (&a)[NDEG], const Cx (&b)[NDEG])
{
    bool same = true;
#pragma unroll
    for (int i = 0; i < NDEG; i++)
        same &= (__double_as_longlong(a[i].re) == __double_as_longlong(b[i].re)) &&
                (__double_as_longlong(a[i].im) == __double_as_longlong(b[i].im));
    return same;
}

// cv::solvePoly (Durand-Kerner, Gauss-Seidel order), degree exactly NDEG (registers, fully unrolled).
// cv::solvePoly always burns maxIters = 1000 sweeps (it only stops on a correction of exactly 0), but a sweep is a
// deterministic map of the root vector, and in floating point that map falls into a short cycle (period 1-12
// in > 97 % of cases) after ~30-50 sweeps.  CYCLE_JUMP detects the cycle exactly (Brent, bitwise state equality)
// and jumps to the state sweep number maxIters would have produced: bit-identical to running every sweep.
template <int NDEG, bool CYCLE_JUMP>
__device__ __forceinline__ void dkSolveFixed(const double *c /*ascending, NDEG+1*/, Cx *rootsOut, int maxIters, double tolSq)
{
    Cx roots[NDEG];
    double co[NDEG + 1];
#pragma unroll
    for (int i = 0; i <= NDEG; i++) co[i] = c[i];
    {
        Cx p{1, 0};
        const Cx r{1, 1};
#pragma unroll
        for (int i = 0; i < NDEG; i++) {
            roots[i] = p;
            p = cmul(p, r);
        }
    }
    if (CYCLE_JUMP) {
        Cx tort[NDEG];
#pragma unroll
        for (int i = 0; i < NDEG; i++) tort[i] = roots[i];
        int power = 1, lam = 0, done = 0;  // `done` sweeps applied to `roots`; tort = state after (done - lam) sweeps
        bool stopped = false;
        while (done < maxIters) {
            const double md = dkSweep<NDEG>(co, roots);
            ++done; ++lam;
            if (md <= tolSq) { stopped = true; break; }
            if (sameRoots<NDEG>(roots, tort)) {
                // state(done) == state(done - lam): period lam.  state(maxIters) == state(done + ((maxIters - done) % lam))
                int rest = (maxIters - done) % lam;
                for (; rest > 0; --rest) dkSweep<NDEG>(co, roots);  // no early stop possible inside a cycle whose sweeps all had md > tolSq
                stopped = true;
                break;
            }
            if (power == lam) {
#pragma unroll
                for (int i = 0; i < NDEG; i++) tort[i] = roots[i];
                power *= 2;
                lam = 0;
            }
        }
        (void)stopped;
    } else {
        for (int iter = 0; iter < maxIters; iter++) {
            const double md = dkSweep<NDEG>(co, roots);
            if (md <= tolSq) break;
        }
    }
#pragma unroll
    for (int i = 0; i < NDEG; i++) {
        if (fabs(roots[i].im) < 1e-100) roots[i].im = 0;
        rootsOut[i] = roots[i];
    }
}

// Generic-degree version (only reached when leading coefficients are ~0; local memory).
__device__ inline void dkSolveGeneric(const double *c, int n, Cx *roots, int maxIters, double tolSq)
{
    Cx p{1, 0};
    const Cx r{1, 1};
    for (int i = 0; i < n; i++) {
        roots[i] = p;
        p = cmul(p, r);
    }
    for (int iter = 0; iter < maxIters; iter++) {
        double maxDiffSq = 0;
        for (int i = 0; i < n; i++) {
            p = roots[i];
            Cx num{c[n], 0.0}, denom{c[n], 0.0};
            for (int j = 0; j < n; j++) {
                num = cmul(num, p);
                num.re = num.re + c[n - j - 1];
                num.im = num.im + 0.0;
                if (j != i) {
                    const Cx d{p.re - roots[j].re, p.im - roots[j].im};
                    if (d.re != 0 || d.im != 0) denom = cmul(denom, d);
                }
            }
            num = cdiv(num, denom);
            roots[i].re = p.re - num.re;
            roots[i].im = p.im - num.im;
            const double m = num.re * num.re + num.im * num.im;
            maxDiffSq = maxDiffSq < m ? m : maxDiffSq;
        }
        if (maxDiffSq <= tolSq) break;
    }
    for (int i = 0; i < n; i++)
        if (fabs(roots[i].im) < 1e-100) roots[i].im = 0;
}

// polynomial index tables: deg1 [x y z 1], deg2 [x2 xy xz x y2 yz y z2 z 1],
// deg3 (Nister order) [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
__device__ __constant__ const unsigned char kT2[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
__device__ __constant__ const unsigned char kT3[10][4] = {{0, 2, 4, 5},     {2, 3, 8, 9},     {4, 8, 10, 11},  {5, 9, 11, 12},
                                                        {3, 1, 6, 7},     {8, 6, 13, 14},   {9, 7, 14, 15},  {10, 13, 16, 17},
                                                        {11, 14, 17, 18}, {12, 15, 18, 19}};

__device__ inline void mul11(const double a[4], const double b[4], double out[10])
{
    for (int k = 0; k < 10; k++) out[k] = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[kT2[i][j]] += a[i] * b[j];
}
__device__ inline void mulAcc21(const double a[10], const double b[4], double *out /*20*/)
{
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 4; j++) out[kT3[i][j]] += a[i] * b[j];
}

// 10x20 constraint matrix: row 0 det(E), rows 1..9 (E E^T E - 1/2 tr(E E^T) E)_{ij}.
__device__ inline void buildConstraints(const double *EE /*4 x 9*/, double *A /*200*/)
{
    double e[9][4];
    for (int k = 0; k < 9; k++)
        for (int b = 0; b < 4; b++) e[k][b] = EE[b * 9 + k];
    for (int k = 0; k < 200; k++) A[k] = 0.0;
    {
        double m1[10], m2[10], d[10];
        mul11(e[4], e[8], m1); mul11(e[5], e[7], m2);
        for (int k = 0; k < 10; k++) d[k] = m1[k] - m2[k];
        mulAcc21(d, e[0], A);
        mul11(e[3], e[8], m1); mul11(e[5], e[6], m2);
        for (int k = 0; k < 10; k++) d[k] = m2[k] - m1[k];
        mulAcc21(d, e[1], A);
        mul11(e[3], e[7], m1); mul11(e[4], e[6], m2);
        for (int k = 0; k < 10; k++) d[k] = m1[k] - m2[k];
        mulAcc21(d, e[2], A);
    }
    double EEt[6][10];  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    {
        int idx = 0;
        for (int i = 0; i < 3; i++)
            for (int j = i; j < 3; j++, idx++) {
                double m[10];
                mul11(e[i * 3 + 0], e[j * 3 + 0], EEt[idx]);
                mul11(e[i * 3 + 1], e[j * 3 + 1], m);
                for (int k = 0; k < 10; k++) EEt[idx][k] += m[k];
                mul11(e[i * 3 + 2], e[j * 3 + 2], m);
                for (int k = 0; k < 10; k++) EEt[idx][k] += m[k];
            }
    }
    const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    double htr[10];
    for (int k = 0; k < 10; k++) htr[k] = 0.5 * ((EEt[0][k] + EEt[3][k]) + EEt[5][k]);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double *row = A + (1 + i * 3 + j) * 20;
            for (int k = 0; k < 3; k++) {
                double L[10];
                for (int q = 0; q < 10; q++) L[q] = EEt[sym[i][k]][q] - (i == k ? htr[q] : 0.0);
                mulAcc21(L, e[k * 3 + j], row);
            }
        }
}

__device__ inline void pmulz(const double *a, int la, const double *b, int lb, double *out)
{
    for (int k = 0; k < la + lb - 1; k++) out[k] = 0.0;
    for (int i = 0; i < la; i++)
        for (int j = 0; j < lb; j++) out[i + j] += a[i] * b[j];
}

// EMEstimatorCallback::runKernel for exactly 5 points.  Writes up to maxOut (<=10) row-major unit-norm
// essential matrices to Eout and returns the TOTAL number of solutions found (<= 10) when maxOut == 10,
// or min(total, maxOut) when the caller only needs the first ones.
// EXT_WS: the large work arrays (kFivePointWs doubles) live in caller-provided memory — shared memory when one warp
// lane solves alone (K2 on small waves): a lone lane's local memory is spread over 32-lane-interleaved lines and
// thrashes L1, shared memory does not.  Same operations on the same operands either way.
constexpr int kFivePointWs = 81 + 200 + 100 + 100 + 39 + 60;

// Front half of the kernel: null space of the 5 epipolar constraints (Vt rows 5..8 = EE), the 10x20 constraint
// matrix, its elimination, the three 13-coefficient rows b and the degree-10 determinant polynomial c (ascending).
// Returns the polynomial's degree n (10 unless leading coefficients vanish).
template <bool EXT_WS>
__device__ __forceinline__ int fivePointFront(const double x1[10], const double x2[10], double *Vt /*81*/, double *A /*200*/,
                                              double *A1 /*100*/, double *inv /*100*/, double *b /*39*/, double *R6 /*60*/,
                                              double c[11])
{
    for (int i = 0; i < 81; i++) Vt[i] = 0.0;
    for (int i = 0; i < 5; i++) {
        const double a = x1[2 * i], bb = x1[2 * i + 1], cc = x2[2 * i], d = x2[2 * i + 1];
        double *q = Vt + 9 * i;
        q[0] = a * cc; q[1] = bb * cc; q[2] = cc; q[3] = a * d; q[4] = bb * d; q[5] = d; q[6] = a; q[7] = bb; q[8] = 1.0;
    }
    double W[5];
    cvJacobiSVD<9, 5, 9, false>(Vt, W, nullptr);
    const double *EE = Vt + 45;

    buildConstraints(EE, A);

    // A1 = A[:,0:10], A2 = A[:,10:20]; R = inv(A1) * A2 (rows 4..9 are the ones consumed)
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) A1[i * 10 + j] = A[i * 20 + j];
    cvInvert10(A1, inv);
    {
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 10; j++) {
                double s = 0;
                for (int k = 0; k < 10; k++) s += inv[(i + 4) * 10 + k] * A[k * 20 + 10 + j];
                R6[i * 10 + j] = s;
            }
        for (int i = 0; i < 3; i++) {
            const double *r1 = R6 + (i * 2) * 10, *r2 = R6 + (i * 2 + 1) * 10;
            double row1[13], row2[13];
            for (int k = 0; k < 13; k++) { row1[k] = 0.0; row2[k] = 0.0; }
            for (int k = 0; k < 3; k++) { row1[1 + k] = r1[k]; row1[5 + k] = r1[3 + k]; }
            for (int k = 0; k < 4; k++) row1[9 + k] = r1[6 + k];
            for (int k = 0; k < 3; k++) { row2[0 + k] = r2[k]; row2[4 + k] = r2[3 + k]; }
            for (int k = 0; k < 4; k++) row2[8 + k] = r2[6 + k];
            for (int k = 0; k < 13; k++) b[i * 13 + k] = row1[k] - row2[k];
        }
    }
    {
        double det[11], m1[7], m2[7], mn[7], t[11];
        for (int k = 0; k < 11; k++) det[k] = 0.0;
        pmulz(b + 13 + 0, 4, b + 26 + 4, 4, m1); pmulz(b + 13 + 4, 4, b + 26 + 0, 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m1[k] - m2[k];
        pmulz(mn, 7, b + 0 + 8, 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
        pmulz(b + 0 + 0, 4, b + 26 + 4, 4, m1); pmulz(b + 0 + 4, 4, b + 26 + 0, 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m2[k] - m1[k];
        pmulz(mn, 7, b + 13 + 8, 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
        pmulz(b + 0 + 0, 4, b + 13 + 4, 4, m1); pmulz(b + 0 + 4, 4, b + 13 + 0, 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m1[k] - m2[k];
        pmulz(mn, 7, b + 26 + 8, 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
        for (int k = 0; k < 11; k++) c[k] = det[10 - k];
    }
    int n = 10;
    for (; n > 1; n--)
        if (fabs(c[n]) + 0.0 > DBL_EPSILON) break;
    return n;
}

// Back half: every real root z gives (x, y) from the 3x3 system b(z) and the essential matrix x E0 + y E1 + z E2 + E3.
// firstUsable (the legacy RANSAC of K2): a solution with a NaN entry is skipped.  cv's RANSAC counts, per solution, the
// correspondences whose error is <= threshold^2 = +inf (ptsetreg.cpp findInliers): every error of a solution with a NaN
// entry is NaN (all nine entries enter the numerator x2^T E x1), so it has no inliers and can never be kept, while a
// finite unit-norm solution has them all and ends the loop — "the first solution of the first sample" is really the
// first FINITE solution (seen once in 44,850 pairs of the sparse benchmark scene).
__device__ __forceinline__ int fivePointFinish(const double *EE /*4 x 9*/, const double *b /*39*/, const Cx *roots, int n,
                                               double *Eout, int maxOut, bool firstUsable = false)
{
    int count = 0;
    for (int i = 0; i < n; i++) {
        if (fabs(roots[i].im) > 1e-10) continue;
        const double z1 = roots[i].re, z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
        double bz[9];
        for (int j = 0; j < 3; j++) {
            const double *br = b + j * 13;
            bz[j * 3 + 0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
            bz[j * 3 + 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
            bz[j * 3 + 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
        }
        double xy1[3];
        cvSolveZ3(bz, xy1);
        if (fabs(xy1[2]) < 1e-10) continue;
        const double x = xy1[0] / xy1[2], y = xy1[1] / xy1[2];
        double Ev[9], nrm = 0;
        for (int k = 0; k < 9; k++) {
            Ev[k] = EE[0 * 9 + k] * x + EE[1 * 9 + k] * y + EE[2 * 9 + k] * z1 + EE[3 * 9 + k];
            nrm += Ev[k] * Ev[k];
        }
        nrm = sqrt(nrm);
        if (firstUsable) {
            bool finite = true;
            for (int k = 0; k < 9; k++) {
                const double e = Ev[k] / nrm;
                finite &= (e == e);
            }
            if (!finite) continue;
        }
        if (count < maxOut)
            for (int k = 0; k < 9; k++) Eout[count * 9 + k] = Ev[k] / nrm;
        count++;
        if (count >= maxOut && maxOut < 10) break;
    }
    return count;
}

template <bool CYCLE_JUMP, bool EXT_WS = false>
__device__ inline int fivePoint(const double x1[10], const double x2[10], double *Eout, int maxOut, int dkMaxIters,
                                double dkTolSq, double *ws = nullptr, bool firstUsable = false)
{
    double VtL[EXT_WS ? 1 : 81], AL[EXT_WS ? 1 : 200], A1L[EXT_WS ? 1 : 100], invL[EXT_WS ? 1 : 100], bL[EXT_WS ? 1 : 39],
        R6L[EXT_WS ? 1 : 60];
    double *const Vt = EXT_WS ? ws : VtL;
    double *const A = EXT_WS ? ws + 81 : AL;
    double *const A1 = EXT_WS ? ws + 281 : A1L;
    double *const inv = EXT_WS ? ws + 381 : invL;
    double *const b = EXT_WS ? ws + 481 : bL;
    double *const R6 = EXT_WS ? ws + 520 : R6L;  // rows 4..9 of inv*A2
    double c[11];
    const int n = fivePointFront<EXT_WS>(x1, x2, Vt, A, A1, inv, b, R6, c);
    Cx roots[10];
    if (n == 10)
        dkSolveFixed<10, CYCLE_JUMP>(c, roots, dkMaxIters, dkTolSq);
    else
        dkSolveGeneric(c, n, roots, dkMaxIters, dkTolSq);
    return fivePointFinish(Vt + 45, b, roots, n, Eout, maxOut, firstUsable);
}

// ---------------------------------------------------------------------------------------------
// Eigen-owned arithmetic (JacobiSVD.h, Jacobi.h, RealSvd2x2.h — Eigen 3.4.0 semantics)
// ---------------------------------------------------------------------------------------------
struct Rot { double c, s; };

PGI_DEV Rot makeJacobi(double x, double y, double z)
{
    const double deno = 2.0 * fabs(y);
    if (deno < DBL_MIN) return Rot{1.0, 0.0};
    const double tau = (x - z) / deno;
    const double w = sqrt(tau * tau + 1.0);
    double t;
    if (tau > 0.0)
        t = 1.0 / (tau + w);
    else
        t = 1.0 / (tau - w);
    const double sign_t = t > 0.0 ? 1.0 : -1.0;
    const double n = 1.0 / sqrt(t * t + 1.0);
    Rot r;
    r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
    r.c = n;
    return r;
}

PGI_DEV void real2x2JacobiSvd(double m00, double m01, double m10, double m11, Rot &jLeft, Rot &jRight)
{
    Rot rot1;
    const double t = m00 + m11;
    const double d = m10 - m01;
    if (fabs(d) < DBL_MIN) {
        rot1.s = 0.0;
        rot1.c = 1.0;
    } else {
        const double u = t / d;
        const double tmp = sqrt(1.0 + u * u);
        rot1.s = 1.0 / tmp;
        rot1.c = u / tmp;
    }
    if (!(rot1.c == 1.0 && rot1.s == 0.0)) {
        const double a00 = rot1.c * m00 + rot1.s * m10, a01 = rot1.c * m01 + rot1.s * m11;
        const double a10 = -rot1.s * m00 + rot1.c * m10, a11 = -rot1.s * m01 + rot1.c * m11;
        m00 = a00; m01 = a01; m10 = a10; m11 = a11;
    }
    jRight = makeJacobi(m00, m01, m11);
    const Rot tr{jRight.c, -jRight.s};
    jLeft.c = rot1.c * tr.c - rot1.s * tr.s;
    jLeft.s = rot1.c * tr.s + rot1.s * tr.c;
}

// One (p,q) step of the two-sided Jacobi sweep on register-resident W (N x N row-major), optional U, V.
template <int N, int P, int Q, bool WITH_U, bool WITH_V>
PGI_DEV void jacobiStep(double (&W)[N * N], double (&U)[WITH_U ? N * N : 1], double (&V)[WITH_V ? N * N : 1],
                        double &maxDiagEntry, bool &finished)
{
    const double precision = 2.0 * DBL_EPSILON;
    const double considerAsZero = DBL_MIN;
    const double pm = precision * maxDiagEntry;
    const double threshold = considerAsZero < pm ? pm : considerAsZero;
    if (fabs(W[P * N + Q]) > threshold || fabs(W[Q * N + P]) > threshold) {
        finished = false;
        Rot jl, jr;
        real2x2JacobiSvd(W[P * N + P], W[P * N + Q], W[Q * N + P], W[Q * N + Q], jl, jr);
        if (!(jl.c == 1.0 && jl.s == 0.0)) {
#pragma unroll
            for (int k = 0; k < N; k++) {
                const double xi = W[P * N + k], yi = W[Q * N + k];
                W[P * N + k] = jl.c * xi + jl.s * yi;
                W[Q * N + k] = -jl.s * xi + jl.c * yi;
            }
            if (WITH_U) {
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const double xi = U[k * N + P], yi = U[k * N + Q];
                    U[k * N + P] = jl.c * xi + jl.s * yi;
                    U[k * N + Q] = -jl.s * xi + jl.c * yi;
                }
            }
        }
        const double ns = -jr.s;
        if (!(jr.c == 1.0 && ns == 0.0)) {
#pragma unroll
            for (int k = 0; k < N; k++) {
                const double xi = W[k * N + P], yi = W[k * N + Q];
                W[k * N + P] = jr.c * xi + ns * yi;
                W[k * N + Q] = -ns * xi + jr.c * yi;
            }
            if (WITH_V) {
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const double xi = V[k * N + P], yi = V[k * N + Q];
                    V[k * N + P] = jr.c * xi + ns * yi;
                    V[k * N + Q] = -ns * xi + jr.c * yi;
                }
            }
        }
        const double ap = fabs(W[P * N + P]), aq = fabs(W[Q * N + Q]);
        const double m2 = ap < aq ? aq : ap;
        maxDiagEntry = maxDiagEntry < m2 ? m2 : maxDiagEntry;
    }
}

template <int N, int P, int Q, bool WITH_U, bool WITH_V>
struct SweepUnroll {
    static PGI_DEV void run(double (&W)[N * N], double (&U)[WITH_U ? N * N : 1], double (&V)[WITH_V ? N * N : 1],
                            double &mde, bool &fin)
    {
        jacobiStep<N, P, Q, WITH_U, WITH_V>(W, U, V, mde, fin);
        if constexpr (Q + 1 < P)
            SweepUnroll<N, P, Q + 1, WITH_U, WITH_V>::run(W, U, V, mde, fin);
        else if constexpr (P + 1 < N)
            SweepUnroll<N, P + 1, 0, WITH_U, WITH_V>::run(W, U, V, mde, fin);
    }
};

// Eigen::JacobiSVD<Matrix<double,N,N>>; A, U, V row-major; S descending.  Register resident.
template <int N, bool WITH_U, bool WITH_V>
PGI_DEV void eigenJacobiSvd(const double (&A)[N * N], double (&U)[WITH_U ? N * N : 1], double (&V)[WITH_V ? N * N : 1],
                            double (&S)[N])
{
    double W[N * N];
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < N * N; i++) {
        const double a = fabs(A[i]);
        if (a > scale || a != a) scale = a;
    }
    if (!(fabs(scale) <= DBL_MAX)) {  // !isfinite
        const double nan = scale - scale;
#pragma unroll
        for (int i = 0; i < N * N; i++) {
            if (WITH_U) U[i] = nan;
            if (WITH_V) V[i] = nan;
        }
#pragma unroll
        for (int i = 0; i < N; i++) S[i] = nan;
        return;
    }
    if (scale == 0.0) scale = 1.0;
#pragma unroll
    for (int i = 0; i < N * N; i++) W[i] = A[i] / scale;
#pragma unroll
    for (int i = 0; i < N * N; i++) {
        if (WITH_U) U[i] = (i / N == i % N) ? 1.0 : 0.0;
        if (WITH_V) V[i] = (i / N == i % N) ? 1.0 : 0.0;
    }
    double maxDiagEntry = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const double a = fabs(W[i * N + i]);
        maxDiagEntry = maxDiagEntry < a ? a : maxDiagEntry;
    }
    bool finished = false;
    while (!finished) {
        finished = true;
        SweepUnroll<N, 1, 0, WITH_U, WITH_V>::run(W, U, V, maxDiagEntry, finished);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double a = W[i * N + i];
        S[i] = fabs(a);
        if (WITH_U && a < 0.0) {
#pragma unroll
            for (int k = 0; k < N; k++) U[k * N + i] = -U[k * N + i];
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) S[i] *= scale;
    bool stop = false;
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (stop) continue;
        int pos = 0;
        double mx = S[i];
#pragma unroll
        for (int k = 1; k < N; k++)
            if (k < N - i && S[(i + k) % N] > mx) { mx = S[(i + k) % N]; pos = k; }
        if (mx == 0.0) { stop = true; continue; }
        if (pos) {
            // swap entry/columns i and i+pos (select-based so indices stay compile-time)
#pragma unroll
            for (int jj = 1; jj < N; jj++) {
                if (jj == pos && i + jj < N) {
                    const int j2 = (i + jj) % N;
                    const double ts = S[i]; S[i] = S[j2]; S[j2] = ts;
                    if (WITH_U) {
#pragma unroll
                        for (int k = 0; k < N; k++) { const double t = U[k * N + i]; U[k * N + i] = U[k * N + j2]; U[k * N + j2] = t; }
                    }
                    if (WITH_V) {
#pragma unroll
                        for (int k = 0; k < N; k++) { const double t = V[k * N + i]; V[k * N + i] = V[k * N + j2]; V[k * N + j2] = t; }
                    }
                }
            }
        }
    }
}

PGI_DEV double det3(const double *M)
{
    const double a = M[0] * (M[4] * M[8] - M[5] * M[7]);
    const double b = M[1] * (M[3] * M[8] - M[5] * M[6]);
    const double c = M[2] * (M[3] * M[7] - M[4] * M[6]);
    return a - b + c;
}

// pose_utils.h:144-169
__device__ inline void decomposeEssential(const double E[9], double R1[9], double R2[9], double t[3])
{
    double A[9], U[9], V[9], S[3];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = E[i];
    eigenJacobiSvd<3, true, true>(A, U, V, S);
    if (det3(U) < 0) { U[2] *= -1.0; U[5] *= -1.0; U[8] *= -1.0; }
    if (det3(V) < 0) { V[2] *= -1.0; V[5] *= -1.0; V[8] *= -1.0; }
    double Ud[9], Udt[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Ud[i * 3 + 0] = -U[i * 3 + 1]; Ud[i * 3 + 1] = U[i * 3 + 0]; Ud[i * 3 + 2] = U[i * 3 + 2];
        Udt[i * 3 + 0] = U[i * 3 + 1]; Udt[i * 3 + 1] = -U[i * 3 + 0]; Udt[i * 3 + 2] = U[i * 3 + 2];
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            R1[i * 3 + j] = sum3(Ud[i * 3 + 0] * V[j * 3 + 0], Ud[i * 3 + 1] * V[j * 3 + 1], Ud[i * 3 + 2] * V[j * 3 + 2]);
            R2[i * 3 + j] = sum3(Udt[i * 3 + 0] * V[j * 3 + 0], Udt[i * 3 + 1] * V[j * 3 + 1], Udt[i * 3 + 2] * V[j * 3 + 2]);
        }
    const double u0 = U[2], u1 = U[5], u2 = U[8];
    const double z = sum3(u0 * u0, u1 * u1, u2 * u2);
    if (z > 0.0) {
        const double n = sqrt(z);
        t[0] = u0 / n; t[1] = u1 / n; t[2] = u2 / n;
    } else {
        t[0] = u0; t[1] = u1; t[2] = u2;
    }
}

// Loop body of pose_utils.h:203-231 for one (candidate, correspondence): P2 = [R | tc] row-major 3x4.
PGI_DEV bool triangulateAndScore(const double P2[12], double c0, double c1, double c2, double c3, double &error)
{
    // design matrix with proj_1 = [I|0]:  c*P1.row(2) - P1.row(k) evaluated entry by entry
    double D[16];
    D[0] = c0 * 0.0 - 1.0; D[1] = c0 * 0.0 - 0.0; D[2] = c0 * 1.0 - 0.0; D[3] = c0 * 0.0 - 0.0;
    D[4] = c1 * 0.0 - 0.0; D[5] = c1 * 0.0 - 1.0; D[6] = c1 * 1.0 - 0.0; D[7] = c1 * 0.0 - 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        D[8 + k] = c2 * P2[8 + k] - P2[0 + k];
        D[12 + k] = c3 * P2[8 + k] - P2[4 + k];
    }
    double Udummy[1], V[16], S[4];
    eigenJacobiSvd<4, false, true>(D, Udummy, V, S);
    const double X0 = V[3], X1 = V[7], X2 = V[11], X3 = V[15];
    const double p1x = sum4(1.0 * X0, 0.0 * X1, 0.0 * X2, 0.0 * X3);
    const double p1y = sum4(0.0 * X0, 1.0 * X1, 0.0 * X2, 0.0 * X3);
    const double p1z = sum4(0.0 * X0, 0.0 * X1, 1.0 * X2, 0.0 * X3);
    if (p1z < 0) return false;
    const double p2x = sum4(P2[0] * X0, P2[1] * X1, P2[2] * X2, P2[3] * X3);
    const double p2y = sum4(P2[4] * X0, P2[5] * X1, P2[6] * X2, P2[7] * X3);
    const double p2z = sum4(P2[8] * X0, P2[9] * X1, P2[10] * X2, P2[11] * X3);
    if (p2z < 0) return false;
    const double a0 = p1x / p1z - c0, a1 = p1y / p1z - c1;
    const double b0 = p2x / p2z - c2, b1 = p2y / p2z - c3;
    error = (a0 * a0 + a1 * a1) + (b0 * b0 + b1 * b1);
    return true;
}

}  // namespace pgi
