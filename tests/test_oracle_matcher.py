"""Oracle restatement of the epipolar-hashing guided matcher (matcher.h:199-405): behavioural checks on synthetic
two-view keypoints with descriptors (the reference has no tests or fixtures for it)."""
import numpy as np

from helpers import two_view_keypoints


def test_guided_matches_are_true_correspondences_and_close_to_their_epipolar_lines(oracle):
    d = two_view_keypoints(1500, np.random.default_rng(0))
    r = oracle.guided_match(d["kp_src"], d["desc_src"], d["kp_dst"], d["desc_dst"], d["pose"], d["K"], d["K"], d["size"], d["size"])
    m = r["matches"]
    assert len(m) > 50
    assert (d["truth"][m[:, 0]] == m[:, 1]).mean() > 0.97
    assert np.all(np.diff(m[:, 0].astype(np.int64)) > 0)  # one match per source keypoint, in source order
    assert np.all(r["ratios"] >= 0.00001) and np.all((r["ratios"] < 0.64) | (r["ratios"] >= 0.64))
    F = r["prepared"][:9].reshape(3, 3)
    x1 = np.hstack([d["kp_src"][m[:, 0]].astype(np.float64), np.ones((len(m), 1))])
    x2 = np.hstack([d["kp_dst"][m[:, 1]].astype(np.float64), np.ones((len(m), 1))])
    l1, l2 = x2 @ F, x1 @ F.T  # epipolar lines in image 1 / image 2
    res = np.einsum("ij,ij->i", x1, l1)
    d2 = res ** 2 * (1.0 / (l1[:, 0] ** 2 + l1[:, 1] ** 2) + 1.0 / (l2[:, 0] ** 2 + l2[:, 1] ** 2))
    assert np.all(d2 < 0.75 ** 2)
    # the epipole is the right null vector of F
    e = np.array([r["prepared"][9], r["prepared"][10], 1.0])
    assert np.linalg.norm(F @ e) / np.linalg.norm(F) < 1e-6


def test_selection_keeps_the_smallest_ratios_or_repeats_the_first(oracle):
    d = two_view_keypoints(3000, np.random.default_rng(1))
    args = (d["kp_src"], d["desc_src"], d["kp_dst"], d["desc_dst"], d["pose"], d["K"], d["K"], d["size"], d["size"])
    r = oracle.guided_match(*args, max_points=20)
    assert len(r["matches"]) > 20 and len(r["selected_matches"]) == 20
    assert np.allclose(np.sort(r["ratios"])[:20], r["selected_ratios"])
    r = oracle.guided_match(*args, max_points=100000)
    assert len(r["selected_matches"]) == len(r["matches"]) and np.all(r["selected_ratios"] == r["ratios"][0])  # :777-779 (sic)
