// ORACLE — TEST INFRASTRUCTURE ONLY (see pgo_geom.hpp header).
// extern "C" surface of the CPU oracle, loaded through ctypes by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include <atomic>
#include <thread>

#include "pgo_cv.hpp"
#include "pgo_eigen.hpp"
#include "pgo_estimate.hpp"
#include "pgo_fallback.hpp"
#include "pgo_geom.hpp"
#include "pgo_host.hpp"
#include "pgo_matcher.hpp"

using namespace pgo;

static SE3 se3FromArray(const double *qt)
{
    return SE3{Quat{qt[0], qt[1], qt[2], qt[3]}, Vec3{qt[4], qt[5], qt[6]}};
}
static void se3ToArray(const SE3 &T, double *qt)
{
    qt[0] = T.q.x; qt[1] = T.q.y; qt[2] = T.q.z; qt[3] = T.q.w;
    qt[4] = T.t.x; qt[5] = T.t.y; qt[6] = T.t.z;
}

extern "C" {

// ---- geometry ------------------------------------------------------------------------------------
void pgo_sampson_sq(const double *corr, uint64_t n, const double *E, double *out)
{
    for (uint64_t i = 0; i < n; i++) out[i] = sampsonSq(corr + 4 * i, E);
}
void pgo_essential_from_pose(const double *qt, double *E) { essentialFromPose(se3FromArray(qt), E); }
void pgo_se3_mul(const double *a, const double *b, double *out) { se3ToArray(se3Mul(se3FromArray(a), se3FromArray(b)), out); }
void pgo_se3_inverse(const double *a, double *out) { se3ToArray(se3Inverse(se3FromArray(a)), out); }
void pgo_rotation_to_quat(const double *R, double *q)
{
    Quat r = rotationToQuat(R);
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
void pgo_quat_to_rotation(const double *q, double *R) { quatToRotation(Quat{q[0], q[1], q[2], q[3]}, R); }
int pgo_test_pose(const double *corr, uint64_t n, const double *qt, double thr, uint64_t minInliers, uint64_t *count)
{
    size_t c = 0;
    const bool ok = inTraversalTest(corr, n, se3FromArray(qt), thr, minInliers, c);
    *count = c;
    return ok ? 1 : 0;
}
uint64_t pgo_get_inliers(const double *corr, uint64_t n, const double *E, double thr, uint64_t *idx)
{
    std::vector<size_t> in;
    getInliers(corr, n, E, thr, in);
    for (size_t i = 0; i < in.size(); i++) idx[i] = in[i];
    return in.size();
}
void pgo_create_correspondences(const float *kpSrc, const float *kpDst, const uint32_t *matches, uint64_t n, double fx,
                                double fy, double cx, double cy, double thrPx, double *corr, double *thrNorm)
{
    createCorrespondenceMatrix(kpSrc, kpDst, matches, n, fx, fy, cx, cy, thrPx, corr, *thrNorm);
}

// ---- OpenCV-owned stages -------------------------------------------------------------------------
void pgo_cv_rng(uint64_t seed, uint64_t n, uint32_t *out)
{
    cvx::RNG r(seed);
    for (uint64_t i = 0; i < n; i++) out[i] = r.next();
}
void pgo_cv_svd_5x9(const double *Q, double *Vt, double *W) { cvx::svdFullVt5x9(Q, Vt, W); }
void pgo_cv_solvez3(const double *B, double *out) { cvx::solveZ3(B, out); }
int pgo_cv_invert10(const double *A, double *inv) { return cvx::invert10(A, inv) ? 1 : 0; }
int pgo_cv_solve_poly(const double *c, int n0, double *rootsReIm, int maxIters, double tolSq)
{
    std::vector<cvx::Cx> r(n0);
    const int n = cvx::solvePoly(c, n0, r.data(), maxIters, tolSq);
    for (int i = 0; i < n; i++) {
        rootsReIm[2 * i] = r[i].re;
        rootsReIm[2 * i + 1] = r[i].im;
    }
    return n;
}
int pgo_five_point(const double *x1, const double *x2, double *E, int dkMaxIters, double dkTolSq)
{
    return cvx::fivePointKernel(x1, x2, 5, E, dkMaxIters, dkTolSq);
}
int pgo_find_essential_ransac_inf(const double *pts, int count, double *E, uint8_t *mask, int *sample, int *iterations)
{
    cvx::LegacyRansacInfo info;
    const int n = cvx::findEssentialMatRansacInf(pts, count, E, mask, &info);
    if (sample) std::memcpy(sample, info.sample, sizeof(info.sample));
    if (iterations) *iterations = info.iterations;
    return n;
}

int pgo_dbg_last_dk_sweeps() { return cvx::lastDkSweeps(); }

// ---- Eigen-owned stages --------------------------------------------------------------------------
void pgo_eigen_svd3(const double *A, double *U, double *V, double *S) { eig::jacobiSvd<3>(A, U, V, S); }
void pgo_eigen_svd4(const double *A, double *U, double *V, double *S) { eig::jacobiSvd<4>(A, U, V, S); }
void pgo_decompose_essential(const double *E, double *R1, double *R2, double *t) { eig::decomposeEssentialMatrix(E, R1, R2, t); }
int pgo_pose_from_essential(const double *E, const double *corr, uint64_t n, double *R, double *t, uint64_t *votes)
{
    size_t v[4];
    const int r = eig::getPoseFromEssentialMatrix(E, corr, n, R, t, v);
    if (votes) for (int i = 0; i < 4; i++) votes[i] = v[i];
    return r;
}

// ---- fallback ------------------------------------------------------------------------------------
void pgo_sampler_table(int N, int iters, uint32_t *out)
{
    std::vector<uint32_t> t;
    fb::samplerTable(N, iters, t);
    std::memcpy(out, t.data(), t.size() * sizeof(uint32_t));
}
void pgo_iters_table(int N, uint16_t *out)
{
    std::vector<uint16_t> t;
    fb::itersTable(N, t);
    std::memcpy(out, t.data(), t.size() * sizeof(uint16_t));
}
void pgo_score_model(const double *corr, int N, const double *E, double thr, uint64_t *cost, int *inliers)
{
    fb::scoreModel(corr, N, E, thr * thr, (1.5 * thr) * (1.5 * thr), *cost, *inliers);
}
int pgo_ls_refit(const double *corr, int N, const double *Ecur, double thr, double *Eout)
{
    return fb::lsRefit(corr, N, Ecur, thr * thr, Eout) ? 1 : 0;
}
// out: E[9], info[5] = {ok, inliers, iterations, models, loRuns}
void pgo_fallback(const double *corr, int N, double thr, double *E, uint8_t *mask, int *info)
{
    fb::FallbackResult r = fb::runFallback(corr, N, thr, mask);
    std::memcpy(E, r.E, sizeof(r.E));
    info[0] = r.ok; info[1] = r.inliers; info[2] = r.iterations; info[3] = r.models; info[4] = r.loRuns;
}

// ---- estimatePose (pose_graph_builder.h:940-1078) ------------------------------------------------
// out: E[9], pose qt[7], info[8] = {success, branch, inlierNumber, pathInliers, votes0..3}
void pgo_estimate_pose(const double *corr, uint64_t n, double thrNorm, uint64_t minInliers, const double *guesses,
                       uint64_t nGuesses, double *E, double *qt, uint8_t *mask, int64_t *info)
{
    std::vector<SE3> g(nGuesses);
    for (uint64_t i = 0; i < nGuesses; i++) g[i] = se3FromArray(guesses + 7 * i);
    std::vector<uint8_t> m;
    EstimateResult r = estimatePose(corr, n, thrNorm, minInliers, g.data(), g.size(), m);
    std::memcpy(E, r.E, sizeof(r.E));
    se3ToArray(r.pose, qt);
    if (mask) {
        std::memset(mask, 0, n);
        std::memcpy(mask, m.data(), std::min<size_t>(m.size(), n));
    }
    info[0] = r.success; info[1] = r.branch; info[2] = (int64_t)r.inlierNumber; info[3] = (int64_t)r.pathInliers;
    for (int i = 0; i < 4; i++) info[4 + i] = (int64_t)r.votes[i];
}

// Batched, multi-threaded estimatePose over independent pairs — the CPU-baseline leg of bench.py
// (the reference's own parallelism is one OpenMP worker per pair, pose_graph_builder.h:391-397;
// libgomp is absent in this image, so std::thread workers pull pairs from an atomic counter instead).
// hasGuess[p] selects 0 or 1 guess (guesses + 7p).  Returns the thread count used.
int pgo_estimate_pose_batch(const double *corr, const uint64_t *offset, uint64_t nPairs, const double *thrNorm,
                            uint64_t minInliers, const double *guesses, const uint8_t *hasGuess, double *E, double *qt,
                            int64_t *info, int threads)
{
    int used = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (used < 1) used = 1;
    if ((uint64_t)used > nPairs) used = nPairs ? (int)nPairs : 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int64_t p = next.fetch_add(1);
            if (p >= (int64_t)nPairs) break;
            const uint64_t n = offset[p + 1] - offset[p];
            pgo_estimate_pose(corr + 4 * offset[p], n, thrNorm[p], minInliers, guesses + 7 * p, hasGuess[p] ? 1 : 0,
                              E + 9 * p, qt + 7 * p, nullptr, info + 8 * p);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < used; t++) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
    return used;
}

// The per-pair body of PoseGraphBuilder::processImages (pose_graph_builder.h:553-627) over a list of
// (pair, path hypothesis) tuples, straight from the compact scene: createCorrespondenceMatrix (:553-565), the
// in-traversal test of the hypothesis A* composed (graph_traversal.h:790 -> :194-233; 1.5 x threshold, 5 inliers,
// pose_graph_builder.h:798-811), estimatePose with that hypothesis as the only guess if it passed (:616-627,
// :940-1078).  Workers pull the next tuple from a shared counter — the reference's worker-pulls-queue structure
// (:391-413; libgomp is absent in this image, so std::thread).  This is what bench.py's cpu_baseline /
// --impl reference legs time and what --verify compares the GPU verdicts with.
// out: testPassed[n], testCount[n], info[n x 8] = {success, branch, inlierNumber, pathInliers, votes0..3}, E[n x 9], qt[n x 7]
int pgo_scene_pipeline_batch(uint64_t V, const double *focal, const double *size, const uint64_t *kpOffset, const float *kp,
                             const uint32_t *pairViews, const uint64_t *mOffset, const uint32_t *matches, double thrPx,
                             uint64_t minInliers, uint64_t nItems, const uint32_t *pairIds, const double *hyp,
                             const uint8_t *hasHyp, int threads, uint8_t *testPassed, uint32_t *testCount, int64_t *info,
                             double *E, double *qt)
{
    (void)V;
    int used = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (used < 1) used = 1;
    if ((uint64_t)used > nItems) used = nItems ? (int)nItems : 1;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        std::vector<double> corr;
        std::vector<uint8_t> mask;
        for (;;) {
            const int64_t i = next.fetch_add(1);  // queue pop under the mutex, pose_graph_builder.h:400-413
            if (i >= (int64_t)nItems) break;
            const uint32_t p = pairIds[i];
            const uint32_t src = pairViews[2 * p], dst = pairViews[2 * p + 1];
            const size_t n = (size_t)(mOffset[p + 1] - mOffset[p]);
            corr.resize(n * 4);
            double thrNorm = 0.0;
            const double f = focal[src], cx = size[2 * src] / 2.0, cy = size[2 * src + 1] / 2.0;
            createCorrespondenceMatrix(kp + 2 * kpOffset[src], kp + 2 * kpOffset[dst], matches + 2 * mOffset[p], n, f, f, cx, cy,
                                       thrPx, corr.data(), thrNorm);
            std::vector<SE3> poses;
            testPassed[i] = 0;
            testCount[i] = 0;
            if (hasHyp[i]) {
                const SE3 h = se3FromArray(hyp + 7 * i);
                size_t cnt = 0;
                const bool ok = inTraversalTest(corr.data(), n, h, 1.5 * thrNorm, 5, cnt);
                testPassed[i] = ok;
                testCount[i] = (uint32_t)cnt;
                if (ok) poses.push_back(h);
            }
            mask.clear();
            EstimateResult r = estimatePose(corr.data(), n, thrNorm, minInliers, poses.data(), poses.size(), mask);
            std::memcpy(E + 9 * i, r.E, sizeof(r.E));
            se3ToArray(r.pose, qt + 7 * i);
            int64_t *o = info + 8 * i;
            o[0] = r.success; o[1] = r.branch; o[2] = (int64_t)r.inlierNumber; o[3] = (int64_t)r.pathInliers;
            for (int k = 0; k < 4; k++) o[4 + k] = (int64_t)r.votes[k];
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < used; t++) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
    return used;
}

// ---- epipolar-hashing guided matcher (matcher.h:199-405, pose_graph_builder.h:717-783) -----------------------
// prep_out[16] = F[9], epipole[2], minAngle, angularRange, binNumber.  Returns the number of matches of match();
// sel_* receive guidedMatching's selection (at most max_points entries), *n_sel its size.
uint64_t pgo_guided_match(const float *kpS, uint64_t nS, const float *dS, const float *kpD, uint64_t nD, const float *dD,
                          int dim, const double *poseQt, const double *Ks, const double *Kd, int wS, int hS, int wD, int hD,
                          int binNumber, uint64_t maxPoints, uint32_t *matches, double *ratios, uint64_t cap,
                          uint32_t *selMatches, double *selRatios, uint64_t *nSel, double *prepOut)
{
    double E[9];
    essentialFromPose(se3FromArray(poseQt), E);
    const matcher::Prepared P = matcher::prepare(E, Ks, Kd, wS, hS, wD, hD, binNumber);
    if (prepOut) {
        for (int k = 0; k < 9; k++) prepOut[k] = P.F[k];
        prepOut[9] = P.epipole[0]; prepOut[10] = P.epipole[1]; prepOut[11] = P.minAngle; prepOut[12] = P.angularRange;
        prepOut[13] = P.binNumber;
    }
    std::vector<std::pair<uint32_t, uint32_t>> m;
    std::vector<double> r;
    matcher::match(kpS, nS, dS, kpD, nD, dD, dim, P, m, r);
    for (size_t i = 0; i < m.size() && i < cap; i++) {
        matches[2 * i] = m[i].first; matches[2 * i + 1] = m[i].second; ratios[i] = r[i];
    }
    if (nSel) {
        std::vector<std::tuple<uint32_t, uint32_t, double>> sel;
        matcher::selectMatches(m, r, maxPoints, sel);
        *nSel = sel.size();
        for (size_t i = 0; i < sel.size(); i++) {
            selMatches[2 * i] = std::get<0>(sel[i]); selMatches[2 * i + 1] = std::get<1>(sel[i]); selRatios[i] = std::get<2>(sel[i]);
        }
    }
    return m.size();
}

// ---- host loop -----------------------------------------------------------------------------------
uint64_t pgo_pairlog_size() { return sizeof(host::PairLog); }

struct PgoRun {
    host::RunResult r;
};

void *pgo_run_scene(uint64_t V, const double *focal, const double *size, const double *sim, const uint64_t *kpOffset,
                    const float *kp, uint64_t P, const uint32_t *pairViews, const uint64_t *mOffset,
                    const uint32_t *matches, double simThreshold, double thrPx, uint64_t minInliers, uint64_t minPoints,
                    uint64_t maxDepth, double weight, int usePathFinding, uint64_t maxPairs)
{
    host::Scene sc;
    sc.V = V; sc.focal = focal; sc.size = size; sc.sim = sim; sc.kpOffset = kpOffset; sc.kp = kp;
    sc.P = P; sc.pairViews = pairViews; sc.mOffset = mOffset; sc.matches = matches;
    host::Config cfg;
    cfg.similarityThreshold = simThreshold; cfg.inlierOutlierThreshold = thrPx; cfg.minimumInlierNumber = minInliers;
    cfg.minimumPointNumber = minPoints; cfg.maximumSearchDepth = maxDepth; cfg.traversalHeuristicsWeight = weight;
    cfg.usePathFinding = usePathFinding != 0;
    PgoRun *h = new PgoRun;
    h->r = host::run(sc, cfg, maxPairs);
    return h;
}
uint64_t pgo_run_log_count(void *h) { return ((PgoRun *)h)->r.log.size(); }
void pgo_run_log_copy(void *h, void *out) { std::memcpy(out, ((PgoRun *)h)->r.log.data(), ((PgoRun *)h)->r.log.size() * sizeof(host::PairLog)); }
// stats[6] = {edges, pathAccepted, fallbackAccepted, rejected, skipped, corrEvals}, stats[6] = fallbackRuns
void pgo_run_stats(void *h, uint64_t *stats)
{
    const host::RunResult &r = ((PgoRun *)h)->r;
    stats[0] = r.edges.size(); stats[1] = r.pathAccepted; stats[2] = r.fallbackAccepted; stats[3] = r.rejected;
    stats[4] = r.skipped; stats[5] = r.corrEvals; stats[6] = r.fallbackRuns;
}
void pgo_run_free(void *h) { delete (PgoRun *)h; }

// Replay a logged run through the sequential oracle host (host::replay).  out[6] = {checked, mismatches, edges,
// searches, firstBad, firstBadField}
void pgo_replay_scene(uint64_t V, const double *sim, uint64_t P, const uint32_t *pairViews, const uint64_t *mOffset,
                      double simThreshold, uint64_t minPoints, uint64_t maxDepth, double weight, int usePathFinding,
                      uint64_t maxPairs, uint64_t n, const uint32_t *src, const uint32_t *dst, const int64_t *pairIndex,
                      const uint8_t *visible, const uint8_t *hadPath, const uint8_t *committed, const uint32_t *touched,
                      const uint32_t *nCorr, const double *hyp, const double *q, const double *t, const double *score,
                      int64_t *out)
{
    host::Scene sc;
    sc.V = V; sc.sim = sim; sc.P = P; sc.pairViews = pairViews; sc.mOffset = mOffset;
    host::Config cfg;
    cfg.similarityThreshold = simThreshold; cfg.minimumPointNumber = minPoints; cfg.maximumSearchDepth = maxDepth;
    cfg.traversalHeuristicsWeight = weight; cfg.usePathFinding = usePathFinding != 0;
    host::ReplayLog lg;
    lg.n = n; lg.src = src; lg.dst = dst; lg.pairIndex = pairIndex; lg.visible = visible; lg.hadPath = hadPath;
    lg.committed = committed; lg.touchedNodes = touched; lg.nCorr = nCorr; lg.hyp = hyp; lg.q = q; lg.t = t; lg.score = score;
    const host::ReplayResult r = host::replay(sc, cfg, lg, maxPairs);
    out[0] = (int64_t)r.checked; out[1] = (int64_t)r.mismatches; out[2] = (int64_t)r.edges; out[3] = (int64_t)r.searches;
    out[4] = r.firstBad; out[5] = r.firstBadField;
}

}  // extern "C"
