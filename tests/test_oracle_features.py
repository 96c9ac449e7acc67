"""Oracle restatement of matchFeatures (feature_utils.h:103-210) against the cv2 wheel's BFMatcher(NORM_L2SQR) — the
library call the reference makes: same 2-NN indices, squared distances to 1e-5 relative (OpenCV accumulates with SIMD
lanes, the restatement in dimension order), and therefore the same surviving matches with ratios to 1e-6."""
import numpy as np
import pytest

from helpers import two_view_keypoints


def _cv2_match(d1, d2):
    cv2 = pytest.importorskip("cv2")
    fw = cv2.BFMatcher(cv2.NORM_L2SQR).knnMatch(d1, d2, 2)
    bw = cv2.BFMatcher(cv2.NORM_L2SQR).knnMatch(d2, d1, 2)
    good = []
    for i, m in enumerate(fw):
        if len(m) < 2 or len(bw[m[0].trainIdx]) < 2:
            continue
        if m[0].distance < 0.90 * m[1].distance and m[0].queryIdx == bw[m[0].trainIdx][0].trainIdx:
            good.append((m[0].distance / m[1].distance, i, m[0].trainIdx))
    good.sort()
    return good


@pytest.mark.parametrize("n,seed", [(200, 0), (900, 1)])
def test_match_features_agrees_with_cv2(oracle, n, seed):
    d = two_view_keypoints(n, np.random.default_rng(seed))
    m, r = oracle.match_features(d["desc_src"], d["desc_dst"])
    ref = _cv2_match(d["desc_src"], d["desc_dst"])
    assert len(m) > n // 2
    assert [(int(a), int(b)) for a, b in m] == [(i, j) for _, i, j in ref] or \
        sorted((int(a), int(b)) for a, b in m) == sorted((i, j) for _, i, j in ref)  # (order may differ between equal-to-1e-7 ratios)
    by_pair = {(i, j): x for x, i, j in ref}
    for (a, b), x in zip(m, r):
        assert abs(x - by_pair[(int(a), int(b))]) <= 1e-6 * max(1.0, x)
    assert np.all(np.diff(r) >= 0)
    assert (d["truth"][m[:, 0]] == m[:, 1]).mean() > 0.98


def test_match_features_degenerate_inputs(oracle):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((5, 128)).astype(np.float32)
    m, r = oracle.match_features(a, a[:1])          # one train descriptor: no second neighbour, nothing survives (:174-176)
    assert len(m) == 0 and len(r) == 0
    m, r = oracle.match_features(a[:0], a)
    assert len(m) == 0
    b = np.vstack([a, a[2:3]])                      # a duplicated train row: best == second, ratio 1 -> rejected
    m, r = oracle.match_features(a, b)
    assert 2 not in m[:, 0]
