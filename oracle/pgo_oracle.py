"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/_build/libpgo_oracle.so (the CPU restatement of the reference hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libpgo_oracle.so")

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_u16p = C.POINTER(C.c_uint16)
_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def build(force=False):
    """Compile the oracle with its Makefile (g++, reference flags, no FMA contraction)."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in os.listdir(_HERE)
        if f.endswith((".hpp", ".cpp")) or f == "Makefile"
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.pgo_get_inliers.restype = C.c_uint64
        _lib.pgo_pairlog_size.restype = C.c_uint64
        _lib.pgo_run_scene.restype = C.c_void_p
        _lib.pgo_run_log_count.restype = C.c_uint64
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


PAIRLOG_DTYPE = np.dtype(
    [
        ("src", np.uint32), ("dst", np.uint32), ("pairIndex", np.int64),
        ("visible", np.uint8), ("hadPath", np.uint8), ("testPassed", np.uint8), ("branch", np.uint8),
        ("committed", np.uint8),
        ("testCount", np.uint32), ("inlierNumber", np.uint32), ("nCorr", np.uint32), ("touchedNodes", np.uint32),
        ("E", np.float64, (9,)), ("q", np.float64, (4,)), ("t", np.float64, (3,)), ("score", np.float64),
    ],
    align=True,
)


# ---- geometry ---------------------------------------------------------------------------------------
def sampson_sq(corr, E):
    corr, E = _d(corr), _d(E).reshape(9)
    out = np.empty(len(corr))
    lib().pgo_sampson_sq(_p(corr, _dp), C.c_uint64(len(corr)), _p(E, _dp), _p(out, _dp))
    return out


def essential_from_pose(qt):
    qt = _d(qt)
    E = np.empty(9)
    lib().pgo_essential_from_pose(_p(qt, _dp), _p(E, _dp))
    return E.reshape(3, 3)


def se3_mul(a, b):
    a, b = _d(a), _d(b)
    out = np.empty(7)
    lib().pgo_se3_mul(_p(a, _dp), _p(b, _dp), _p(out, _dp))
    return out


def se3_inverse(a):
    a = _d(a)
    out = np.empty(7)
    lib().pgo_se3_inverse(_p(a, _dp), _p(out, _dp))
    return out


def rotation_to_quat(R):
    R = _d(R).reshape(9)
    q = np.empty(4)
    lib().pgo_rotation_to_quat(_p(R, _dp), _p(q, _dp))
    return q


def quat_to_rotation(q):
    q = _d(q)
    R = np.empty(9)
    lib().pgo_quat_to_rotation(_p(q, _dp), _p(R, _dp))
    return R.reshape(3, 3)


def test_pose(corr, qt, thr, min_inliers=5):
    corr, qt = _d(corr), _d(qt)
    cnt = C.c_uint64(0)
    ok = lib().pgo_test_pose(_p(corr, _dp), C.c_uint64(len(corr)), _p(qt, _dp), C.c_double(thr),
                             C.c_uint64(min_inliers), C.byref(cnt))
    return bool(ok), int(cnt.value)


def get_inliers(corr, E, thr):
    corr, E = _d(corr), _d(E).reshape(9)
    idx = np.empty(len(corr), dtype=np.uint64)
    n = lib().pgo_get_inliers(_p(corr, _dp), C.c_uint64(len(corr)), _p(E, _dp), C.c_double(thr), _p(idx, _u64p))
    return idx[:n].copy()


def create_correspondences(kp_src, kp_dst, matches, fx, fy, cx, cy, thr_px):
    kp_src = np.ascontiguousarray(kp_src, dtype=np.float32)
    kp_dst = np.ascontiguousarray(kp_dst, dtype=np.float32)
    matches = np.ascontiguousarray(matches, dtype=np.uint32)
    n = len(matches)
    corr = np.empty((n, 4))
    thr = C.c_double(0)
    lib().pgo_create_correspondences(_p(kp_src, _fp), _p(kp_dst, _fp), _p(matches, _u32p), C.c_uint64(n),
                                     C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                     C.c_double(thr_px), _p(corr, _dp), C.byref(thr))
    return corr, thr.value


# ---- OpenCV-owned stages ------------------------------------------------------------------------------
def cv_rng(seed, n):
    out = np.empty(n, dtype=np.uint32)
    lib().pgo_cv_rng(C.c_uint64(seed), C.c_uint64(n), _p(out, _u32p))
    return out


def cv_svd_5x9(Q):
    Q = _d(Q).reshape(45)
    Vt, W = np.empty(81), np.empty(5)
    lib().pgo_cv_svd_5x9(_p(Q, _dp), _p(Vt, _dp), _p(W, _dp))
    return W, Vt.reshape(9, 9)


def cv_solvez3(B):
    B = _d(B).reshape(9)
    out = np.empty(3)
    lib().pgo_cv_solvez3(_p(B, _dp), _p(out, _dp))
    return out


def cv_invert10(A):
    A = _d(A).reshape(100)
    inv = np.empty(100)
    ok = lib().pgo_cv_invert10(_p(A, _dp), _p(inv, _dp))
    return bool(ok), inv.reshape(10, 10)


def cv_solve_poly(c, max_iters=1000, tol_sq=0.0):
    c = _d(c)
    n0 = len(c) - 1
    r = np.empty(2 * n0)
    n = lib().pgo_cv_solve_poly(_p(c, _dp), C.c_int(n0), _p(r, _dp), C.c_int(max_iters), C.c_double(tol_sq))
    return r[: 2 * n].reshape(n, 2)


def five_point(x1, x2, dk_max_iters=1000, dk_tol_sq=0.0):
    x1, x2 = _d(x1).reshape(10), _d(x2).reshape(10)
    E = np.empty(90)
    n = lib().pgo_five_point(_p(x1, _dp), _p(x2, _dp), _p(E, _dp), C.c_int(dk_max_iters), C.c_double(dk_tol_sq))
    return E[: 9 * n].reshape(n, 3, 3)


def find_essential_ransac_inf(pts):
    pts = _d(pts)
    n = len(pts)
    E = np.empty(90)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    sample = np.zeros(5, dtype=np.int32)
    iters = C.c_int(0)
    nm = lib().pgo_find_essential_ransac_inf(_p(pts, _dp), C.c_int(n), _p(E, _dp), _p(mask, _u8p), _p(sample, _ip),
                                             C.byref(iters))
    return nm, E[: 9 * max(nm, 0)].reshape(-1, 3, 3), mask[:n], sample, iters.value


# ---- Eigen-owned stages -------------------------------------------------------------------------------
def eigen_svd(A):
    A = _d(A)
    n = A.shape[0]
    U, V, S = np.empty(n * n), np.empty(n * n), np.empty(n)
    f = lib().pgo_eigen_svd3 if n == 3 else lib().pgo_eigen_svd4
    f(_p(A.reshape(-1), _dp), _p(U, _dp), _p(V, _dp), _p(S, _dp))
    return U.reshape(n, n), S, V.reshape(n, n)


def decompose_essential(E):
    E = _d(E).reshape(9)
    R1, R2, t = np.empty(9), np.empty(9), np.empty(3)
    lib().pgo_decompose_essential(_p(E, _dp), _p(R1, _dp), _p(R2, _dp), _p(t, _dp))
    return R1.reshape(3, 3), R2.reshape(3, 3), t


def pose_from_essential(E, corr):
    E, corr = _d(E).reshape(9), _d(corr)
    R, t = np.empty(9), np.empty(3)
    votes = np.zeros(4, dtype=np.uint64)
    lib().pgo_pose_from_essential(_p(E, _dp), _p(corr, _dp), C.c_uint64(len(corr)), _p(R, _dp), _p(t, _dp),
                                  _p(votes, _u64p))
    return R.reshape(3, 3), t, votes


# ---- fallback -------------------------------------------------------------------------------------------
def sampler_table(N, iters=1000):
    out = np.empty(iters * 5, dtype=np.uint32)
    lib().pgo_sampler_table(C.c_int(N), C.c_int(iters), _p(out, _u32p))
    return out.reshape(iters, 5)


def iters_table(N):
    out = np.empty(N + 1, dtype=np.uint16)
    lib().pgo_iters_table(C.c_int(N), _p(out, _u16p))
    return out


def score_model(corr, E, thr):
    corr, E = _d(corr), _d(E).reshape(9)
    cost, inl = C.c_uint64(0), C.c_int(0)
    lib().pgo_score_model(_p(corr, _dp), C.c_int(len(corr)), _p(E, _dp), C.c_double(thr), C.byref(cost), C.byref(inl))
    return cost.value, inl.value


def ls_refit(corr, E, thr):
    corr, E = _d(corr), _d(E).reshape(9)
    out = np.empty(9)
    ok = lib().pgo_ls_refit(_p(corr, _dp), C.c_int(len(corr)), _p(E, _dp), C.c_double(thr), _p(out, _dp))
    return bool(ok), out.reshape(3, 3)


def fallback(corr, thr):
    corr = _d(corr)
    E = np.empty(9)
    mask = np.zeros(len(corr), dtype=np.uint8)
    info = np.zeros(5, dtype=np.int32)
    lib().pgo_fallback(_p(corr, _dp), C.c_int(len(corr)), C.c_double(thr), _p(E, _dp), _p(mask, _u8p), _p(info, _ip))
    return dict(ok=int(info[0]), inliers=int(info[1]), iterations=int(info[2]), models=int(info[3]),
                lo_runs=int(info[4]), E=E.reshape(3, 3), mask=mask)


# ---- estimatePose -----------------------------------------------------------------------------------------
def estimate_pose(corr, thr_norm, guesses=(), min_inliers=20):
    corr = _d(corr)
    g = _d(np.asarray(guesses, dtype=np.float64).reshape(-1, 7)) if len(guesses) else np.zeros((0, 7))
    E, qt = np.empty(9), np.empty(7)
    mask = np.zeros(len(corr), dtype=np.uint8)
    info = np.zeros(8, dtype=np.int64)
    lib().pgo_estimate_pose(_p(corr, _dp), C.c_uint64(len(corr)), C.c_double(thr_norm), C.c_uint64(min_inliers),
                            _p(g, _dp), C.c_uint64(len(g)), _p(E, _dp), _p(qt, _dp), _p(mask, _u8p), _p(info, _i64p))
    return dict(success=bool(info[0]), branch=int(info[1]), inlier_number=int(info[2]), path_inliers=int(info[3]),
                votes=info[4:8].copy(), E=E.reshape(3, 3), pose=qt, mask=mask)


def estimate_pose_batch(corr, offset, thr_norm, guesses, has_guess, min_inliers=20, threads=0):
    corr = _d(corr)
    offset = np.ascontiguousarray(offset, dtype=np.uint64)
    thr_norm = _d(thr_norm)
    guesses = _d(guesses)
    has_guess = np.ascontiguousarray(has_guess, dtype=np.uint8)
    n = len(offset) - 1
    E, qt = np.empty((n, 9)), np.empty((n, 7))
    info = np.zeros((n, 8), dtype=np.int64)
    used = lib().pgo_estimate_pose_batch(_p(corr, _dp), _p(offset, _u64p), C.c_uint64(n), _p(thr_norm, _dp),
                                         C.c_uint64(min_inliers), _p(guesses, _dp), _p(has_guess, _u8p), _p(E, _dp),
                                         _p(qt, _dp), _p(info, _i64p), C.c_int(threads))
    return dict(E=E, pose=qt, info=info, threads=used)


# ---- host loop ----------------------------------------------------------------------------------------------
def run_scene(scene, sim_threshold=0.0, thr_px=0.4, min_inliers=20, min_points=50, max_depth=5, weight=0.8,
              use_path_finding=True, max_pairs=0):
    """scene: dict with focal[V], size[V,2], sim[V,V], kp_offset[V+1], kp[K,2] f32, pair_views[P,2] u32,
    m_offset[P+1] u64, matches[M,2] u32 (the container of SURVEY App. D)."""
    L = lib()
    assert L.pgo_pairlog_size() == PAIRLOG_DTYPE.itemsize, (L.pgo_pairlog_size(), PAIRLOG_DTYPE.itemsize)
    focal = _d(scene["focal"]); size = _d(scene["size"]); sim = _d(scene["sim"])
    kpo = np.ascontiguousarray(scene["kp_offset"], dtype=np.uint64)
    kp = np.ascontiguousarray(scene["kp"], dtype=np.float32)
    pv = np.ascontiguousarray(scene["pair_views"], dtype=np.uint32)
    mo = np.ascontiguousarray(scene["m_offset"], dtype=np.uint64)
    mt = np.ascontiguousarray(scene["matches"], dtype=np.uint32)
    h = L.pgo_run_scene(C.c_uint64(len(focal)), _p(focal, _dp), _p(size, _dp), _p(sim, _dp), _p(kpo, _u64p),
                        _p(kp, _fp), C.c_uint64(len(pv)), _p(pv, _u32p), _p(mo, _u64p), _p(mt, _u32p),
                        C.c_double(sim_threshold), C.c_double(thr_px), C.c_uint64(min_inliers), C.c_uint64(min_points),
                        C.c_uint64(max_depth), C.c_double(weight), C.c_int(1 if use_path_finding else 0),
                        C.c_uint64(max_pairs))
    h = C.c_void_p(h)
    n = L.pgo_run_log_count(h)
    log = np.zeros(n, dtype=PAIRLOG_DTYPE)
    if n:
        L.pgo_run_log_copy(h, log.ctypes.data_as(C.c_void_p))
    stats = np.zeros(7, dtype=np.uint64)
    L.pgo_run_stats(h, _p(stats, _u64p))
    L.pgo_run_free(h)
    keys = ["edges", "path_accepted", "fallback_accepted", "rejected", "skipped", "corr_evals", "fallback_runs"]
    return log, dict(zip(keys, (int(s) for s in stats)))


def scene_pipeline_batch(scene, pair_ids, hyp, has_hyp, thr_px=0.4, min_inliers=20, threads=0):
    """Per-pair body of processImages (createCorrespondenceMatrix -> in-traversal test -> estimatePose) over
    (pair, hypothesis) tuples, worker-pulls-queue on `threads` host threads (0 = all)."""
    L = lib()
    focal = _d(scene["focal"]); size = _d(scene["size"])
    kpo = np.ascontiguousarray(scene["kp_offset"], dtype=np.uint64)
    kp = np.ascontiguousarray(scene["kp"], dtype=np.float32)
    pv = np.ascontiguousarray(scene["pair_views"], dtype=np.uint32)
    mo = np.ascontiguousarray(scene["m_offset"], dtype=np.uint64)
    mt = np.ascontiguousarray(scene["matches"], dtype=np.uint32)
    ids = np.ascontiguousarray(pair_ids, dtype=np.uint32)
    n = len(ids)
    hyp = _d(np.asarray(hyp, dtype=np.float64).reshape(n, 7))
    has = np.ascontiguousarray(has_hyp, dtype=np.uint8)
    tp, tc = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint32)
    info = np.zeros((n, 8), dtype=np.int64)
    E, qt = np.zeros((n, 9)), np.zeros((n, 7))
    L.pgo_scene_pipeline_batch.restype = C.c_int
    used = L.pgo_scene_pipeline_batch(C.c_uint64(len(focal)), _p(focal, _dp), _p(size, _dp), _p(kpo, _u64p), _p(kp, _fp),
                                      _p(pv, _u32p), _p(mo, _u64p), _p(mt, _u32p), C.c_double(thr_px), C.c_uint64(min_inliers),
                                      C.c_uint64(n), _p(ids, _u32p), _p(hyp, _dp), _p(has, _u8p), C.c_int(threads),
                                      _p(tp, _u8p), _p(tc, _u32p), _p(info, _i64p), _p(E, _dp), _p(qt, _dp))
    return dict(test_passed=tp, test_count=tc, info=info, E=E, pose=qt, threads=int(used))


def replay_scene(scene, log, sim_threshold=0.0, min_points=50, max_depth=5, weight=0.8, use_path_finding=True, max_pairs=0):
    """Walk the queue with the sequential oracle host, taking the verdicts from a product log (builder.LOG_DTYPE
    records incl. `hyp`) and comparing the host-side decisions.  Returns dict(checked, mismatches, edges, searches,
    first_bad, first_bad_field)."""
    L = lib()
    sim = _d(scene["sim"])
    pv = np.ascontiguousarray(scene["pair_views"], dtype=np.uint32)
    mo = np.ascontiguousarray(scene["m_offset"], dtype=np.uint64)
    f = {k: np.ascontiguousarray(log[k]) for k in ("src", "dst", "pairIndex", "visible", "hadPath", "committed", "touchedNodes",
                                                    "nCorr", "hyp", "q", "t", "score")}
    out = np.zeros(6, dtype=np.int64)
    L.pgo_replay_scene.restype = None
    L.pgo_replay_scene(C.c_uint64(len(sim)), _p(sim, _dp), C.c_uint64(len(pv)), _p(pv, _u32p), _p(mo, _u64p),
                       C.c_double(sim_threshold), C.c_uint64(min_points), C.c_uint64(max_depth), C.c_double(weight),
                       C.c_int(1 if use_path_finding else 0), C.c_uint64(max_pairs), C.c_uint64(len(log)),
                       _p(f["src"].astype(np.uint32), _u32p), _p(f["dst"].astype(np.uint32), _u32p),
                       _p(f["pairIndex"].astype(np.int64), _i64p), _p(f["visible"].astype(np.uint8), _u8p),
                       _p(f["hadPath"].astype(np.uint8), _u8p), _p(f["committed"].astype(np.uint8), _u8p),
                       _p(f["touchedNodes"].astype(np.uint32), _u32p), _p(f["nCorr"].astype(np.uint32), _u32p),
                       _p(_d(f["hyp"]), _dp), _p(_d(f["q"]), _dp), _p(_d(f["t"]), _dp), _p(_d(f["score"]), _dp),
                       _p(out, _i64p))
    keys = ["checked", "mismatches", "edges", "searches", "first_bad", "first_bad_field"]
    return dict(zip(keys, (int(x) for x in out)))


def guided_match(kp_src, desc_src, kp_dst, desc_dst, pose_qt, K_src, K_dst, size_src, size_dst, bin_number=45, max_points=100):
    """HashingBasedMatcherWithPose<false, 45>::match + guidedMatching's selection.  Returns dict(matches[n,2], ratios[n],
    selected[(src, dst, ratio)], prepared[14])."""
    L = lib()
    ks = np.ascontiguousarray(kp_src, dtype=np.float32); kd = np.ascontiguousarray(kp_dst, dtype=np.float32)
    ds = np.ascontiguousarray(desc_src, dtype=np.float32); dd = np.ascontiguousarray(desc_dst, dtype=np.float32)
    cap = len(ks)
    m = np.zeros((max(cap, 1), 2), dtype=np.uint32); r = np.zeros(max(cap, 1))
    sm = np.zeros((max(cap, 1), 2), dtype=np.uint32); sr = np.zeros(max(cap, 1))
    nsel = C.c_uint64(0)
    prep = np.zeros(16)
    L.pgo_guided_match.restype = C.c_uint64
    n = L.pgo_guided_match(_p(ks, _fp), C.c_uint64(len(ks)), _p(ds, _fp), _p(kd, _fp), C.c_uint64(len(kd)), _p(dd, _fp),
                           C.c_int(ds.shape[1]), _p(_d(pose_qt), _dp), _p(_d(K_src), _dp), _p(_d(K_dst), _dp),
                           C.c_int(int(size_src[0])), C.c_int(int(size_src[1])), C.c_int(int(size_dst[0])), C.c_int(int(size_dst[1])),
                           C.c_int(bin_number), C.c_uint64(max_points), _p(m, _u32p), _p(r, _dp), C.c_uint64(cap),
                           _p(sm, _u32p), _p(sr, _dp), C.byref(nsel), _p(prep, _dp))
    n = int(n)
    k = int(nsel.value)
    return dict(matches=m[:n].copy(), ratios=r[:n].copy(), selected_matches=sm[:k].copy(), selected_ratios=sr[:k].copy(), prepared=prep[:14].copy())


def match_features(desc_src, desc_dst):
    """matchFeatures (feature_utils.h:103-210) restated with numpy: BRUTEFORCE_SL2 2-NN in both directions (:151-163), the
    0.90 ratio test on the squared distances and the mutual-nearest-neighbour check (:172-184), survivors sorted by the
    ratio (:188) -> (matches [n, 2] (queryIdx, trainIdx), ratios [n]).
    The squared distance is accumulated in float32 in dimension order, multiply then add, and the 2-NN scan keeps the
    first of equal distances first (cv::batchDistance's insertion rule) — OpenCV's own SIMD accumulation order depends
    on the build, so its distances agree to ~1e-6 only (tests/test_oracle_features.py pins that against the cv2 wheel)."""
    a = np.ascontiguousarray(desc_src, dtype=np.float32)
    b = np.ascontiguousarray(desc_dst, dtype=np.float32)

    def knn2(q, t):
        nq, nt = len(q), len(t)
        d1 = np.full(nq, np.finfo(np.float32).max, dtype=np.float32)
        d2 = d1.copy()
        i1 = np.full(nq, -1, dtype=np.int64)
        if nq == 0 or nt == 0:
            return d1, d2, i1
        step = max(1, (1 << 22) // max(nq, 1))
        for t0 in range(0, nt, step):
            tt = t[t0:t0 + step]
            acc = np.zeros((nq, len(tt)), dtype=np.float32)
            for k in range(q.shape[1]):
                d = q[:, k:k + 1] - tt[None, :, k]
                acc = acc + d * d  # float32 multiply, then float32 add
            for j in range(len(tt)):  # ascending train index, strict comparisons
                c = acc[:, j]
                better = c < d1
                second = ~better & (c < d2)
                d2 = np.where(better, d1, np.where(second, c, d2))
                i1 = np.where(better, t0 + j, i1)
                d1 = np.where(better, c, d1)
        return d1, d2, i1

    f1, f2, fi = knn2(a, b)
    g1, g2, gi = knn2(b, a)
    out = []
    if len(b) >= 2 and len(a) >= 2:  # matches[i].size() < 2 || matches_opposite[...].size() < 2 -> continue  (:174-176)
        for i in range(len(a)):
            if float(f1[i]) < 0.90 * float(f2[i]) and gi[fi[i]] == i:  # (:178-179) float < double * float
                out.append((float(np.float32(f1[i]) / np.float32(f2[i])), i, int(fi[i])))
    out.sort(key=lambda r: (r[0], r[1]))  # std::sort on (ratio, &matches[i])  (:188)
    m = np.array([(i, j) for _, i, j in out], dtype=np.uint32).reshape(-1, 2)
    return m, np.array([r for r, _, _ in out], dtype=np.float64)
