"""K8 (pgi_match_features) against the oracle's restatement of matchFeatures (feature_utils.h:103-210): identical
matches in identical order, bit-identical ratios (same float32 accumulation order on both sides)."""
import numpy as np
import pytest

from helpers import two_view_keypoints

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,seed,dim", [(40, 0, 128), (700, 1, 128), (2500, 2, 128), (1300, 3, 64)])
def test_match_features_equals_oracle(engine, oracle, n, seed, dim):
    d = two_view_keypoints(n, np.random.default_rng(seed))
    a, b = d["desc_src"][:, :dim].copy(), d["desc_dst"][:, :dim].copy()
    m, r = engine.match_features(a, b)
    mo, ro = oracle.match_features(a, b)
    assert np.array_equal(m, mo)
    assert np.array_equal(r.view(np.uint64), ro.view(np.uint64))
    if dim == 128 and n >= 700:
        assert len(m) > n // 2 and (d["truth"][m[:, 0]] == m[:, 1]).mean() > 0.98


def test_match_features_ties_and_degenerate_inputs(engine, oracle):
    rng = np.random.default_rng(5)
    a = rng.integers(-3, 4, (300, 128)).astype(np.float32)   # small integers: many exactly equal distances
    b = np.vstack([a[rng.permutation(300)[:200]] + rng.integers(0, 2, (200, 128)).astype(np.float32), a[:50]])
    m, r = engine.match_features(a, b)
    mo, ro = oracle.match_features(a, b)
    assert np.array_equal(m, mo) and np.array_equal(r.view(np.uint64), ro.view(np.uint64))
    for x, y in ((a[:5], a[:1]), (a[:0], a), (a, a[:0])):
        m, r = engine.match_features(x, y)
        assert len(m) == 0 and len(r) == 0
    with pytest.raises(ValueError):
        engine.match_features(a[:, :100], b[:, :100])
