"""pgb_tracklets_* (the product's Tracklets, pgb_tracklets.cpp) against the statement-by-statement restatement of
point_track.h:541-712 in oracle/pgo_tracklets.py: same correspondences in the same order after random add() sequences,
including the reference's quirks (pair number 0, maximum + 1 matches, several points of one view in a track)."""
import numpy as np
import pytest

from pose_graph_initialization_b200 import builder as B
from oracle.pgo_tracklets import Tracklets as OracleTracklets


def _random_history(rng, views, points, adds, per_add):
    for _ in range(adds):
        s, d = rng.choice(views, 2, replace=False)
        n = int(rng.integers(1, per_add + 1))
        m = np.stack([rng.integers(0, points, n), rng.integers(0, points, n), rng.random(n)], axis=1)
        mask = (rng.random(n) < 0.8).astype(np.uint8)
        yield int(s), int(d), m, mask


@pytest.mark.parametrize("seed,views,points,adds,per_add", [(0, 6, 12, 60, 8), (1, 12, 40, 200, 30), (2, 4, 5, 80, 6), (3, 30, 400, 300, 120)])
def test_tracklets_match_the_restated_reference(seed, views, points, adds, per_add):
    rng = np.random.default_rng(seed)
    mine, ref = B.Tracklets(views), OracleTracklets(views)
    for k, (s, d, m, mask) in enumerate(_random_history(rng, views, points, adds, per_add)):
        mine.add(s, d, m, mask)
        ref.add(s, d, [tuple(r) for r in m], list(mask))
        if k % 7 == 0:
            a, b = (int(x) for x in rng.choice(views, 2, replace=False))
            for maximum in (3, 5000):
                assert mine.getCorrespondences(a, b, maximum) == ref.getCorrespondences(a, b, maximum)
    assert mine.track_count() == len(ref.tmpTracks)
    for a in range(views):
        for b in range(views):
            if a != b:
                assert mine.getCorrespondences(a, b, 5000) == ref.getCorrespondences(a, b, 5000)
    mine.close()


def test_tracklets_quirks_of_the_reference():
    t = B.Tracklets(3)
    assert t.getCorrespondences(0, 1, 10) == []                     # unknown views: nothing (point_track.h:583-596)
    t.add(0, 1, np.array([[5, 7, 0.1], [6, 8, 0.2], [9, 9, 0.3]]), [1, 1, 0])
    assert t.getCorrespondences(0, 1, 10) == [(5, 7, 0.0), (6, 8, 0.0)]   # masked-out match ignored; value member is 0.0
    assert t.getCorrespondences(1, 0, 10) == [(7, 5, 0.0), (8, 6, 0.0)]
    assert len(t.getCorrespondences(0, 1, 0)) == 1                  # stops only AFTER exceeding the maximum (:631-632)
    # (0,5) was the first pair ever: its number is 0 = "not numbered", so its next sight starts a new track instead of
    # extending the old one (point_track.h:655-661); (0,6) extends its track normally
    t.add(0, 2, np.array([[5, 1, 0.0], [6, 2, 0.0]]), [1, 1])
    assert t.track_count() == 3
    assert t.getCorrespondences(1, 2, 10) == [(8, 2, 0.0)]
    with pytest.raises(ValueError):
        t.add(0, 1, np.array([[1, 2, 0.0]]), [1, 1])
    t.close()
