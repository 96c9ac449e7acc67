// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
//
// pgo_cv.hpp — restatement of the OpenCV-owned arithmetic reached from
// pose_graph_builder.h:1013-1020 (cv::findEssentialMat(..., cv::RANSAC, 0.99, DBL_MAX, mask)),
// SURVEY §8a row a7 and App. B.1-B.4.  OpenCV is an un-vendored, un-pinned dependency of the
// reference (CMakeLists.txt:26 `find_package(OpenCV 4.0)`); the algorithms are restated from the
// published OpenCV 4.x sources (core/src/lapack.cpp JacobiSVDImpl_/LUImpl, core/src/mathfuncs.cpp
// solvePoly, core/src/rand.cpp RNG, calib3d/src/five-point.cpp EMEstimatorCallback,
// calib3d/src/ptsetreg.cpp RANSACPointSetRegistrator) and PINNED against the cv2 4.13.0 wheel by
// tests/test_oracle_cv.py + tests/golden/cv_*.npz.
//
// Deviation that is tolerance-level, not bit-level, w.r.t. cv2: the 10x20 constraint matrix
// (OpenCV: generated getCoeffMat) and the degree-10 determinant polynomial (OpenCV: one generated
// expression per coefficient) are expanded by table-driven polynomial arithmetic here.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace pgo {
namespace cvx {

// ---- cv::RNG (core/include/opencv2/core/operations.hpp): 64-bit multiply-with-carry ------------
struct RNG {
    uint64_t state;
    explicit RNG(uint64_t s) : state(s ? s : 0xffffffffULL) {}
    uint32_t next()
    {
        state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
        return (uint32_t)state;
    }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (uint32_t)(b - a) + a); }
};

// the file-local hypot of core/src/lapack.cpp (NOT libm's)
inline double cvHypot(double a, double b)
{
    a = std::abs(a);
    b = std::abs(b);
    if (a > b) {
        b /= a;
        return a * std::sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * std::sqrt(1 + a * a);
    }
    return 0;
}

// ---- cv::hal::SVD64f -> JacobiSVDImpl_<double> (core/src/lapack.cpp) ----------------------------
// (bit-identical to cv2 4.13.0 SVDecomp on 5x9/3x3 inputs with plain scalar loops: the wheel's
// VBLAS<double> SIMD helpers are not active for these sizes — probed 50/50.)
// One-sided Jacobi on the n rows (length m) of At; Vt (n x n) accumulates the rotations; rows
// n..n1-1 of At are filled by the RNG-seeded Gram-Schmidt completion.
inline void jacobiSVD(double *At, int astep, double *W_, double *Vt, int vstep, int m, int n, int n1)
{
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    std::vector<double> Wb(n);
    double *W = Wb.data();
    int i, j, k, iter;
    const int max_iter = std::max(m, 30);
    double c, s, sd;

    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) {
            double t = At[i * astep + k];
            sd += t * t;
        }
        W[i] = sd;
        if (Vt) {
            for (k = 0; k < n; k++) Vt[i * vstep + k] = 0;
            Vt[i * vstep + i] = 1;
        }
    }

    for (iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * astep, *Aj = At + j * astep;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = cvHypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = std::sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = std::sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                if (Vt) {
                    double *Vi = Vt + i * vstep, *Vj = Vt + j * vstep;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }

    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) {
            double t = At[i * astep + k];
            sd += t * t;
        }
        W[i] = std::sqrt(sd);
    }

    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            std::swap(W[i], W[j]);
            if (Vt) {
                for (k = 0; k < m; k++) std::swap(At[i * astep + k], At[j * astep + k]);
                for (k = 0; k < n; k++) std::swap(Vt[i * vstep + k], Vt[j * vstep + k]);
            }
        }
    }
    for (i = 0; i < n; i++) W_[i] = W[i];
    if (!Vt) return;

    RNG rng(0x12345678);
    for (i = 0; i < n1; i++) {
        sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                double val = (rng.next() & 256) != 0 ? val0 : -val0;
                At[i * astep + k] = val;
            }
            for (iter = 0; iter < 2; iter++) {
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * astep + k] * At[j * astep + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        double t = At[i * astep + k] - sd * At[j * astep + k];
                        At[i * astep + k] = t;
                        asum += std::abs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * astep + k] *= asum;
                }
            }
            sd = 0;
            for (k = 0; k < m; k++) {
                double t = At[i * astep + k];
                sd += t * t;
            }
            sd = std::sqrt(sd);
        }
        s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * astep + k] *= s;
    }
}

// cv::SVD::compute(Q(5x9), W, U, Vt, MODIFY_A|FULL_UV) as called at five-point.cpp runKernel:
// rows<cols => "at" branch of _SVDcompute: the 5 rows of Q are rotated in a zero-initialised 9x9
// buffer (m=9, n=5, n1=9) and that buffer IS the returned Vt.
inline void svdFullVt5x9(const double Q[45], double Vt[81], double W[5])
{
    std::memset(Vt, 0, 81 * sizeof(double));
    std::memcpy(Vt, Q, 45 * sizeof(double));
    double V5[25];
    jacobiSVD(Vt, 9, W, V5, 5, 9, 5, 9);
}

// cv::SVD::solveZ(Bz(3x3)) : SVD(m, 0): temp_a = Bz^T, JacobiSVD(m=3,n=3,n1=3); result = last row of vt.
inline void solveZ3(const double Bz[9], double out[3])
{
    double At[9], W[3], Vt[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) At[i * 3 + j] = Bz[j * 3 + i];
    jacobiSVD(At, 3, W, Vt, 3, 3, 3, 3);
    out[0] = Vt[6]; out[1] = Vt[7]; out[2] = Vt[8];
}

// ---- cv::invert(DECOMP_LU) for n>3 : LUImpl on a copy with identity RHS (core/src/lapack.cpp) ---
inline int luSolve(double *A, int astep, int m, double *b, int bstep, int n)
{
    const double eps = DBL_EPSILON * 100;
    int i, j, k, p = 1;
    for (i = 0; i < m; i++) {
        k = i;
        for (j = i + 1; j < m; j++)
            if (std::abs(A[j * astep + i]) > std::abs(A[k * astep + i])) k = j;
        if (std::abs(A[k * astep + i]) < eps) return 0;
        if (k != i) {
            for (j = i; j < m; j++) std::swap(A[i * astep + j], A[k * astep + j]);
            if (b)
                for (j = 0; j < n; j++) std::swap(b[i * bstep + j], b[k * bstep + j]);
            p = -p;
        }
        double d = -1 / A[i * astep + i];
        for (j = i + 1; j < m; j++) {
            double alpha = A[j * astep + i] * d;
            for (k = i + 1; k < m; k++) A[j * astep + k] += alpha * A[i * astep + k];
            if (b)
                for (k = 0; k < n; k++) b[j * bstep + k] += alpha * b[i * bstep + k];
        }
    }
    if (b) {
        for (i = m - 1; i >= 0; i--)
            for (j = 0; j < n; j++) {
                double s = b[i * bstep + j];
                for (k = i + 1; k < m; k++) s -= A[i * astep + k] * b[k * bstep + j];
                b[i * bstep + j] = s / A[i * astep + i];
            }
    }
    return p;
}

// inv(A1) for 10x10; returns false (and zero matrix, as cv::invert does) if singular.
inline bool invert10(const double A1[100], double inv[100])
{
    double tmp[100];
    std::memcpy(tmp, A1, sizeof(tmp));
    std::memset(inv, 0, 100 * sizeof(double));
    for (int i = 0; i < 10; i++) inv[i * 11] = 1.0;
    if (luSolve(tmp, 10, 10, inv, 10, 10) == 0) {
        std::memset(inv, 0, 100 * sizeof(double));
        return false;
    }
    return true;
}

// ---- cv::solvePoly (core/src/mathfuncs.cpp) — Durand-Kerner, Gauss-Seidel order ----------------
struct Cx { double re, im; };
inline int &lastDkSweeps() { static thread_local int v = 0; return v; }  // diagnostics only
inline Cx cmul(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cx cadd(Cx a, Cx b) { return Cx{a.re + b.re, a.im + b.im}; }
inline Cx csub(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
inline Cx cdiv(Cx a, Cx b)
{
    double t = 1. / (b.re * b.re + b.im * b.im);
    return Cx{(a.re * b.re + a.im * b.im) * t, (-a.re * b.im + a.im * b.re) * t};
}

// coeffs ascending (c[0] constant ... c[n0] leading). Returns number of roots written (n after
// trimming ~0 leading coefficients). The repeated-root branch of OpenCV (num_same_root>1) is
// reproduced only as "skip the factor"; its sqrt/cubic correction is never reached on this path
//  and is flagged through *sawRepeated.
inline int solvePoly(const double *c, int n0, Cx *roots, int maxIters = 1000, double tolSq = 0.0,
                     bool *sawRepeated = nullptr)
{
    int n = n0, i, j, iter;
    std::vector<Cx> coeffs(n0 + 1);
    for (i = 0; i <= n0; i++) coeffs[i] = Cx{c[i], 0.0};
    for (; n > 1; n--)
        if (std::abs(coeffs[n].re) + std::abs(coeffs[n].im) > DBL_EPSILON) break;
    Cx p{1, 0}, r{1, 1};
    for (i = 0; i < n; i++) {
        roots[i] = p;
        p = cmul(p, r);
    }
    for (iter = 0; iter < maxIters; iter++) {
        double maxDiffSq = 0;  // cv: maxDiff = max |num|, break if maxDiff <= 0  <=>  max |num|^2 <= 0
        for (i = 0; i < n; i++) {
            p = roots[i];
            Cx num = coeffs[n], denom = coeffs[n];
            for (j = 0; j < n; j++) {
                num = cadd(cmul(num, p), coeffs[n - j - 1]);
                if (j != i) {
                    Cx d = csub(p, roots[j]);
                    if (d.re != 0 || d.im != 0)
                        denom = cmul(denom, d);
                    else if (sawRepeated)
                        *sawRepeated = true;
                }
            }
            num = cdiv(num, denom);
            roots[i] = csub(p, num);
            maxDiffSq = std::max(maxDiffSq, num.re * num.re + num.im * num.im);
        }
        if (maxDiffSq <= tolSq) { ++iter; break; }
    }
    lastDkSweeps() = iter;
    const double verySmallEps = 1e-100;
    for (i = 0; i < n; i++)
        if (std::fabs(roots[i].im) < verySmallEps) roots[i].im = 0;
    return n;
}

// ---- polynomial tables for the five-point constraints -------------------------------------------
// deg-1 monomials [x y z 1]; deg-2 (10): [x2 xy xz x y2 yz y z2 z 1];
// deg-3 (20) in Nister's column order (five-point.cpp / SURVEY App. B.3):
// [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
struct PolyTables {
    int t2[4][4];    // deg1 x deg1 -> deg2 index
    int t3[10][4];   // deg2 x deg1 -> deg3 index
    PolyTables()
    {
        const int e1[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
        const int e2[10][3] = {{2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 2, 0},
                               {0, 1, 1}, {0, 1, 0}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
        const int e3[20][3] = {{3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1},
                               {0, 2, 0}, {1, 1, 1}, {1, 1, 0}, {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2},
                               {0, 1, 1}, {0, 1, 0}, {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                int ex[3] = {e1[i][0] + e1[j][0], e1[i][1] + e1[j][1], e1[i][2] + e1[j][2]};
                for (int k = 0; k < 10; k++)
                    if (e2[k][0] == ex[0] && e2[k][1] == ex[1] && e2[k][2] == ex[2]) t2[i][j] = k;
            }
        for (int i = 0; i < 10; i++)
            for (int j = 0; j < 4; j++) {
                int ex[3] = {e2[i][0] + e1[j][0], e2[i][1] + e1[j][1], e2[i][2] + e1[j][2]};
                for (int k = 0; k < 20; k++)
                    if (e3[k][0] == ex[0] && e3[k][1] == ex[1] && e3[k][2] == ex[2]) t3[i][j] = k;
            }
    }
};
inline const PolyTables &polyTables()
{
    static const PolyTables t;
    return t;
}

// out(10) = a(4) * b(4); accumulation order: i outer, j inner, out zero-initialised.
inline void mul11(const double a[4], const double b[4], double out[10])
{
    const PolyTables &T = polyTables();
    for (int k = 0; k < 10; k++) out[k] = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[T.t2[i][j]] += a[i] * b[j];
}
// out(20) += a(10) * b(4)
inline void mulAcc21(const double a[10], const double b[4], double out[20])
{
    const PolyTables &T = polyTables();
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 4; j++) out[T.t3[i][j]] += a[i] * b[j];
}

// 10x20 constraint matrix from the null-space basis EE (4 x 9, row k = vec(E_k) row-major),
// E(x,y,z) = x E0 + y E1 + z E2 + E3.  Row 0: det(E); rows 1..9: (E E^T E - 1/2 tr(E E^T) E)_{ij}.
inline void buildConstraints(const double EE[36], double A[200])
{
    double e[9][4];  // entry (r*3+c) as linear poly [x y z 1]
    for (int k = 0; k < 9; k++)
        for (int b = 0; b < 4; b++) e[k][b] = EE[b * 9 + k];
    std::memset(A, 0, 200 * sizeof(double));

    // det(E) = e0 (e4 e8 - e5 e7) - e1 (e3 e8 - e5 e6) + e2 (e3 e7 - e4 e6)
    {
        double m1[10], m2[10], d[10], neg[10];
        mul11(e[4], e[8], m1); mul11(e[5], e[7], m2);
        for (int k = 0; k < 10; k++) d[k] = m1[k] - m2[k];
        mulAcc21(d, e[0], A);
        mul11(e[3], e[8], m1); mul11(e[5], e[6], m2);
        for (int k = 0; k < 10; k++) neg[k] = m2[k] - m1[k];
        mulAcc21(neg, e[1], A);
        mul11(e[3], e[7], m1); mul11(e[4], e[6], m2);
        for (int k = 0; k < 10; k++) d[k] = m1[k] - m2[k];
        mulAcc21(d, e[2], A);
    }
    // EEt (symmetric 3x3 of deg-2 polys)
    double EEt[3][3][10];
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) {
            double acc[10], m[10];
            mul11(e[i * 3 + 0], e[j * 3 + 0], acc);
            mul11(e[i * 3 + 1], e[j * 3 + 1], m);
            for (int k = 0; k < 10; k++) acc[k] += m[k];
            mul11(e[i * 3 + 2], e[j * 3 + 2], m);
            for (int k = 0; k < 10; k++) {
                acc[k] += m[k];
                EEt[i][j][k] = acc[k];
                EEt[j][i][k] = acc[k];
            }
        }
    // Lambda = EEt - 1/2 trace I
    double L[3][3][10];
    for (int k = 0; k < 10; k++) {
        const double htr = 0.5 * ((EEt[0][0][k] + EEt[1][1][k]) + EEt[2][2][k]);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) L[i][j][k] = EEt[i][j][k] - (i == j ? htr : 0.0);
    }
    // rows 1..9 : (Lambda E)_{ij} = sum_k Lambda_ik e_kj
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double *row = A + (1 + i * 3 + j) * 20;
            for (int k = 0; k < 3; k++) mulAcc21(L[i][k], e[k * 3 + j], row);
        }
}

// polynomial (in z) helpers, highest power first like B's rows in five-point.cpp
// out(len la+lb-1) = a * b
inline void pmulz(const double *a, int la, const double *b, int lb, double *out)
{
    for (int k = 0; k < la + lb - 1; k++) out[k] = 0.0;
    for (int i = 0; i < la; i++)
        for (int j = 0; j < lb; j++) out[i + j] += a[i] * b[j];
}

// EMEstimatorCallback::runKernel (calib3d/src/five-point.cpp).  x1,x2: 5 points each (x,y).
// Writes up to 10 essential matrices (row-major, unit Frobenius norm) and returns their count.
// dkMaxIters/dkTolSq = (1000, 0) reproduce cv::solvePoly; the fallback (pgo_fallback.hpp) passes its own.
inline int fivePointKernel(const double *x1, const double *x2, int npts, double *Eout /*10x9*/,
                           int dkMaxIters = 1000, double dkTolSq = 0.0)
{
    if (npts != 5) return 0;
    double Q[45];
    for (int i = 0; i < 5; i++) {
        const double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        double *q = Q + 9 * i;
        // row = [x1x2, y1x2, x2, x1y2, y1y2, y2, x1, y1, 1]  => row-major e with x2^T E x1 = 0 (probe-verified vs cv2)
        q[0] = a * c; q[1] = b * c; q[2] = c; q[3] = a * d; q[4] = b * d; q[5] = d; q[6] = a; q[7] = b; q[8] = 1.0;
    }
    double Vt[81], W[5];
    svdFullVt5x9(Q, Vt, W);
    const double *EE = Vt + 45;  // rows 5..8

    double A[200];
    buildConstraints(EE, A);

    // A <- inv(A[:,0:10]) * A[:,10:20]   (cv::invert LU + cv::gemm: sequential k accumulation)
    double A1[100], A2[100], inv[100], R[100];
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) {
            A1[i * 10 + j] = A[i * 20 + j];
            A2[i * 10 + j] = A[i * 20 + 10 + j];
        }
    invert10(A1, inv);
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) {
            double s = 0;
            for (int k = 0; k < 10; k++) s += inv[i * 10 + k] * A2[k * 10 + j];
            R[i * 10 + j] = s;
        }

    double b[39];
    for (int i = 0; i < 3; i++) {
        const double *r1 = R + (i * 2 + 4) * 10, *r2 = R + (i * 2 + 5) * 10;
        double row1[13] = {0}, row2[13] = {0};
        for (int k = 0; k < 3; k++) { row1[1 + k] = r1[k]; row1[5 + k] = r1[3 + k]; }
        for (int k = 0; k < 4; k++) row1[9 + k] = r1[6 + k];
        for (int k = 0; k < 3; k++) { row2[0 + k] = r2[k]; row2[4 + k] = r2[3 + k]; }
        for (int k = 0; k < 4; k++) row2[8 + k] = r2[6 + k];
        for (int k = 0; k < 13; k++) b[i * 13 + k] = row1[k] - row2[k];
    }

    // det B(z): cofactor expansion along column 2 (the degree-4 column)
    // det = B02 (B10 B21 - B11 B20) - B12 (B00 B21 - B01 B20) + B22 (B00 B11 - B01 B10)
    double det[11];  // highest power first: z^10 ... 1
    {
        auto P = [&](int r, int col) { return b + r * 13 + col * 4; };  // col 0,1: 4 coeffs; col 2: 5 coeffs at +8
        double m1[7], m2[7], mn[7], t[11];
        for (int k = 0; k < 11; k++) det[k] = 0.0;
        pmulz(P(1, 0), 4, P(2, 1), 4, m1); pmulz(P(1, 1), 4, P(2, 0), 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m1[k] - m2[k];
        pmulz(mn, 7, P(0, 2), 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
        pmulz(P(0, 0), 4, P(2, 1), 4, m1); pmulz(P(0, 1), 4, P(2, 0), 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m2[k] - m1[k];
        pmulz(mn, 7, P(1, 2), 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
        pmulz(P(0, 0), 4, P(1, 1), 4, m1); pmulz(P(0, 1), 4, P(1, 0), 4, m2);
        for (int k = 0; k < 7; k++) mn[k] = m1[k] - m2[k];
        pmulz(mn, 7, P(2, 2), 5, t);
        for (int k = 0; k < 11; k++) det[k] += t[k];
    }
    double c[11];  // ascending for solvePoly
    for (int k = 0; k < 11; k++) c[k] = det[10 - k];

    Cx roots[10];
    const int nroots = solvePoly(c, 10, roots, dkMaxIters, dkTolSq);

    int count = 0;
    for (int i = 0; i < nroots; i++) {
        if (std::fabs(roots[i].im) > 1e-10) continue;
        const double z1 = roots[i].re, z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
        double bz[9];
        for (int j = 0; j < 3; j++) {
            const double *br = b + j * 13;
            bz[j * 3 + 0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
            bz[j * 3 + 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
            bz[j * 3 + 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
        }
        double xy1[3];
        solveZ3(bz, xy1);
        if (std::fabs(xy1[2]) < 1e-10) continue;
        const double x = xy1[0] / xy1[2], y = xy1[1] / xy1[2];
        double Ev[9], nrm = 0;
        for (int k = 0; k < 9; k++) {
            Ev[k] = EE[0 * 9 + k] * x + EE[1 * 9 + k] * y + EE[2 * 9 + k] * z1 + EE[3 * 9 + k];
            nrm += Ev[k] * Ev[k];
        }
        nrm = std::sqrt(nrm);
        for (int k = 0; k < 9; k++) Eout[count * 9 + k] = Ev[k] / nrm;
        count++;
    }
    return count;
}

// cv::findEssentialMat(x1, x2, I, RANSAC, 0.99, DBL_MAX, mask)  — five-point.cpp +
// ptsetreg.cpp RANSACPointSetRegistrator::run, specialised to threshold = DBL_MAX (every finite
// error is an inlier) as called at pose_graph_builder.h:1013-1020.  pts: count x 4 [x1 y1 x2 y2].
// Returns number of 3x3 models written to E (0 = failure, >1 only when count == 5);
// mask (count bytes) is written only on success, as OpenCV does.
// Durand-Kerner stopping rule used on this path.  cv::solvePoly burns maxIters = 1000 sweeps unless a correction is
// exactly 0; the sweeps after convergence only jitter the last bit.  The restatement stops at max |correction|^2 <=
// kLegacyDkTolSq = 1e-22 (the next error is then ~1e-22 by quadratic convergence; or cv's 1000-sweep cap): same Gauss-Seidel trajectory, hence the same root -> index assignment
// and the same "first admissible root"; values agree with cv2 to ~1e-13 like the rest of the kernel
// (tests/test_oracle_golden.py pins count, order and values against cv2.findEssentialMat).
constexpr double kLegacyDkTolSq = 1e-22;

struct LegacyRansacInfo {
    int iterations = 0;       // samples drawn
    int sample[5] = {0, 0, 0, 0, 0};
};
inline int findEssentialMatRansacInf(const double *pts, int count, double *E /*up to 10x9*/, uint8_t *mask,
                                     LegacyRansacInfo *info = nullptr)
{
    const int modelPoints = 5;
    if (count < modelPoints) return 0;
    double x1[10], x2[10];
    if (count == modelPoints) {
        for (int i = 0; i < 5; i++) {
            x1[2 * i] = pts[4 * i]; x1[2 * i + 1] = pts[4 * i + 1];
            x2[2 * i] = pts[4 * i + 2]; x2[2 * i + 1] = pts[4 * i + 3];
        }
        int n = fivePointKernel(x1, x2, 5, E, 1000, kLegacyDkTolSq);
        if (n <= 0) return 0;
        if (mask) std::memset(mask, 1, count);
        return n;
    }
    RNG rng((uint64_t)-1);
    int niters = 1000, maxGoodCount = 0;
    double models[90], best[9];
    for (int iter = 0; iter < niters; iter++) {
        int idx[5];
        for (int i = 0; i < modelPoints; ++i) {
            int idx_i;
            for (idx_i = rng.uniform(0, count); std::find(idx, idx + i, idx_i) != idx + i;
                 idx_i = rng.uniform(0, count)) {}
            idx[i] = idx_i;
            x1[2 * i] = pts[4 * idx_i]; x1[2 * i + 1] = pts[4 * idx_i + 1];
            x2[2 * i] = pts[4 * idx_i + 2]; x2[2 * i + 1] = pts[4 * idx_i + 3];
        }
        if (info) {
            info->iterations = iter + 1;
            for (int i = 0; i < 5; i++) info->sample[i] = idx[i];
        }
        const int nmodels = fivePointKernel(x1, x2, 5, models, 1000, kLegacyDkTolSq);
        if (nmodels <= 0) continue;
        for (int mi = 0; mi < nmodels; mi++) {
            // findInliers with thresh^2 = +inf: err <= inf unless NaN (EMEstimatorCallback::computeError)
            const double *M = models + 9 * mi;
            int good = 0;
            for (int i = 0; i < count; i++) {
                const double *p = pts + 4 * i;
                const double Ex0 = M[0] * p[0] + M[1] * p[1] + M[2], Ex1 = M[3] * p[0] + M[4] * p[1] + M[5],
                             Ex2 = M[6] * p[0] + M[7] * p[1] + M[8];
                const double Et0 = M[0] * p[2] + M[3] * p[3] + M[6], Et1 = M[1] * p[2] + M[4] * p[3] + M[7];
                const double x2tEx1 = p[2] * Ex0 + p[3] * Ex1 + Ex2;
                const float err = (float)(x2tEx1 * x2tEx1 / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1));
                good += (err <= INFINITY) ? 1 : 0;
            }
            if (good > std::max(maxGoodCount, modelPoints - 1)) {
                std::memcpy(best, M, sizeof(best));
                maxGoodCount = good;
                // RANSACUpdateNumIters(0.99, (count-good)/count, 5, niters)
                const double ep = (double)(count - good) / count;
                double num = std::max(1. - 0.99, DBL_MIN);
                double denom = 1. - std::pow(1. - ep, modelPoints);
                if (denom < DBL_MIN)
                    niters = 0;
                else {
                    num = std::log(num); denom = std::log(denom);
                    niters = (denom >= 0 || -num >= niters * (-denom)) ? niters : (int)std::lrint(num / denom);
                }
                if (mask) {
                    // mask of the winning model (all ones unless NaN errors)
                    for (int i = 0; i < count; i++) mask[i] = 1;
                }
            }
        }
    }
    if (maxGoodCount > 0) {
        std::memcpy(E, best, sizeof(best));
        return 1;
    }
    return 0;
}

}  // namespace cvx
}  // namespace pgo
