"""Property tests of the Eigen/Sophus-owned stages of the oracle (no executable Eigen exists here: SURVEY §8c Tier A)
and of the host restatements."""
import numpy as np

from helpers import pose_qt, rodrigues, two_view


def test_eigen_jacobi_svd_properties(oracle):
    rng = np.random.default_rng(1)
    for n in (3, 4):
        for _ in range(100):
            A = rng.standard_normal((n, n)) * 10 ** rng.uniform(-3, 3)
            U, S, V = oracle.eigen_svd(A)
            assert np.all(np.diff(S) <= 0) and np.all(S >= 0)
            assert np.abs(U @ np.diag(S) @ V.T - A).max() <= 1e-13 * max(1, np.abs(A).max())
            assert np.abs(V.T @ V - np.eye(n)).max() < 1e-14 and np.abs(U.T @ U - np.eye(n)).max() < 1e-14
    # rank-deficient and zero inputs
    U, S, V = oracle.eigen_svd(np.zeros((4, 4)))
    assert np.all(S == 0)
    A = np.outer(rng.standard_normal(4), rng.standard_normal(4))
    U, S, V = oracle.eigen_svd(A)
    assert S[1] < 1e-14 * S[0]


def test_se3_algebra(oracle):
    rng = np.random.default_rng(2)
    for _ in range(50):
        a = pose_qt(rodrigues(rng.standard_normal(3)), rng.standard_normal(3))
        b = pose_qt(rodrigues(rng.standard_normal(3)), rng.standard_normal(3))
        ab = oracle.se3_mul(a, b)
        Ra, Rb = oracle.quat_to_rotation(a[:4]), oracle.quat_to_rotation(b[:4])
        assert np.allclose(oracle.quat_to_rotation(ab[:4]), Ra @ Rb, atol=1e-14)
        assert np.allclose(ab[4:], Ra @ b[4:] + a[4:], atol=1e-14)
        ident = oracle.se3_mul(a, oracle.se3_inverse(a))
        assert np.allclose(np.abs(ident[3]), 1, atol=1e-15) and np.allclose(ident[4:], 0, atol=1e-14)
        q = oracle.rotation_to_quat(Ra)
        assert np.allclose(np.abs(q @ a[:4]), 1, atol=1e-14)


def test_essential_from_pose_and_sampson(oracle):
    rng = np.random.default_rng(3)
    corr, R, t = two_view(200, 0.0, rng, noise_px=0.0)
    E = oracle.essential_from_pose(pose_qt(R, t))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    assert np.allclose(E, tx @ R, atol=1e-15)
    s = oracle.sampson_sq(corr, E)
    assert s.max() < 1e-25  # noise-free correspondences lie on their epipolar lines
    # literal formula check on one row (graph_traversal.h:107-115)
    x1, y1, x2, y2 = corr[0]
    e = E
    rxc = e[0, 0] * x2 + e[1, 0] * y2 + e[2, 0]; ryc = e[0, 1] * x2 + e[1, 1] * y2 + e[2, 1]; rwc = e[0, 2] * x2 + e[1, 2] * y2 + e[2, 2]
    r = x1 * rxc + y1 * ryc + rwc
    rx = e[0, 0] * x1 + e[0, 1] * y1 + e[0, 2]; ry = e[1, 0] * x1 + e[1, 1] * y1 + e[1, 2]
    assert s[0] == r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry)


def test_get_inliers_uses_unsquared_threshold(oracle):
    # getInliers compares the SQUARED residual with the UN-squared threshold; test() squares it
    rng = np.random.default_rng(4)
    corr, R, t = two_view(500, 0.5, rng)
    E = oracle.essential_from_pose(pose_qt(R, t))
    s = oracle.sampson_sq(corr, E)
    thr = 1.5 * 0.4 / 800
    assert np.array_equal(oracle.get_inliers(corr, E, thr), np.nonzero(s < thr)[0].astype(np.uint64))
    ok, cnt = oracle.test_pose(corr, pose_qt(R, t), thr, 5)
    assert ok and cnt == 5  # early exit leaves inlierNumber_ at the minimum (graph_traversal.h:221-225)
    ok, cnt = oracle.test_pose(corr, pose_qt(rodrigues(np.array([0.5, 0.2, -0.4])) @ R, t), thr, 5)
    assert (not ok) and cnt == np.count_nonzero(oracle.sampson_sq(corr, oracle.essential_from_pose(
        pose_qt(rodrigues(np.array([0.5, 0.2, -0.4])) @ R, t))) < thr * thr)


def test_decomposition_vote_recovers_rotation(oracle):
    rng = np.random.default_rng(5)
    corr, R, t = two_view(400, 0.0, rng, noise_px=0.05)
    E = oracle.essential_from_pose(pose_qt(R, t))
    Ro, to, votes = oracle.pose_from_essential(E / np.linalg.norm(E), corr)
    ang = np.arccos(np.clip((np.trace(Ro.T @ R) - 1) / 2, -1, 1))
    assert ang < 1e-6
    assert abs(abs(to @ t) - 1) < 1e-9  # direction up to the noise-decided sign
    assert votes.sum() == len(corr)
    # the twisted-pair rotation never collects votes; +t / -t split them
    R1, R2, tt = oracle.decompose_essential(E)
    assert abs(np.linalg.det(R1) - 1) < 1e-12 and abs(np.linalg.det(R2) - 1) < 1e-12 and abs(np.linalg.norm(tt) - 1) < 1e-14


def test_create_correspondences_uses_source_intrinsics_for_both(oracle):
    # SURVEY §0.8 / pose_graph_builder.h:908-912
    kp_s = np.array([[100.5, 200.25], [800, 600]], dtype=np.float32)
    kp_d = np.array([[50, 60], [1000.125, 900]], dtype=np.float32)
    corr, thr = oracle.create_correspondences(kp_s, kp_d, np.array([[0, 1], [1, 0]]), 800.0, 800.0, 800.0, 600.0, 0.4)
    assert np.array_equal(corr[0], [(100.5 - 800) / 800, (200.25 - 600) / 800, (float(np.float32(1000.125)) - 800) / 800, (900 - 600) / 800])
    assert thr == 0.4 / 800


def test_sampler_and_iters_tables(oracle):
    t = oracle.sampler_table(57, 200)
    assert t.shape == (200, 5) and t.max() < 57
    assert all(len(set(row)) == 5 for row in t.tolist())  # 5 distinct indices per sample
    it = oracle.iters_table(100)
    assert it[0] == 1000 and it[100] == 1 and np.all(np.diff(it.astype(int)) <= 0)
