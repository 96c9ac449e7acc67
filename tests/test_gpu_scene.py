"""End-to-end graph identity on the GPU: PoseGraphBuilder.run (speculative waves + sm_100a engine through the
C-ABI) must commit the pose graph the sequential CPU oracle commits on the same synthetic scene: same edges in
the same order, bit-identical poses, inlier counts and scores (BASELINE.json north_star correctness bar)."""
import numpy as np
import pytest

from pose_graph_initialization_b200 import builder as B
from pose_graph_initialization_b200 import scene as S

pytestmark = pytest.mark.gpu


def run_and_compare(oracle, sc, **kw):
    olog, ostats = oracle.run_scene(sc, sim_threshold=0.0)
    pgb = B.PoseGraphBuilder(kCoreNumber_=4, kSimilarityThreshold_=0.0, scene=sc, **kw)
    graph = pgb.run()
    plog = pgb.log
    assert len(plog) == len(olog)
    for f in ("src", "dst", "pairIndex", "visible", "hadPath", "testPassed", "branch", "committed", "testCount",
              "inlierNumber", "nCorr", "touchedNodes", "E", "q", "t", "score"):
        assert np.array_equal(plog[f], olog[f]), f
    assert graph.numEdges() == ostats["edges"]
    committed = olog[olog["committed"] > 0]
    assert np.array_equal(graph.edges["src"], committed["src"]) and np.array_equal(graph.edges["dst"], committed["dst"])
    assert np.array_equal(graph.edges["q"], committed["q"]) and np.array_equal(graph.edges["t"], committed["t"])
    assert np.array_equal(graph.edges["score"], committed["score"])
    pgb.close()
    return pgb, ostats


@pytest.mark.parametrize("prefetch,wave,overlap", [(True, 64, True), (True, 1000, False), (False, 16, False)])
def test_small_scene_graph_identity(oracle, prefetch, wave, overlap):
    sc = S.make_scene(n_views=14, n_corr=300, outlier_ratio=0.3, seed=21, n_points=900)
    pgb, ostats = run_and_compare(oracle, sc, prefetch_fallback=prefetch, wave_size=wave, overlap_fallback=overlap, fallback_wave=32)
    assert ostats["edges"] > 0


def test_sparse_fallback_heavy_scene(oracle):
    # cfg-4 style: only ring neighbours share points; distant pairs are 100 % outliers and must be rejected
    sc = S.make_scene(n_views=12, n_corr=250, outlier_ratio=0.5, seed=22, n_points=800, overlap_knn=2,
                      overlap_outlier_ratio=0.5)
    pgb, ostats = run_and_compare(oracle, sc, prefetch_fallback=True, wave_size=32)
    assert ostats["rejected"] > 0


def test_accurate_low_noise_scene_takes_path_branch(oracle):
    # tiny noise => composed path hypotheses pass the in-traversal test and the path branch is exercised
    sc = S.make_scene(n_views=10, n_corr=300, outlier_ratio=0.2, seed=23, n_points=900, noise_px=0.01)
    pgb, ostats = run_and_compare(oracle, sc, prefetch_fallback=False, wave_size=8)
    assert ostats["path_accepted"] + ostats["fallback_accepted"] == ostats["edges"]


@pytest.mark.parametrize("native", [True, False])
def test_a_pass_driven_in_batches_commits_the_same_graph(native):
    """bench.py's steps are batches of one pass (PoseGraphBuilder.run_in_batches): through the native wave driver and
    through the Python loop, the batched pass must commit byte for byte what run() commits, and yield once per batch."""
    sc = S.make_scene(n_views=16, n_corr=300, outlier_ratio=0.3, seed=24, n_points=900)
    P = len(sc["pair_views"])
    ref = B.PoseGraphBuilder(kCoreNumber_=4, kSimilarityThreshold_=0.0, scene=sc, wave_size=16, fallback_wave=32, native_loop=native)
    g_ref = ref.run()
    pgb = B.PoseGraphBuilder(kCoreNumber_=4, kSimilarityThreshold_=0.0, scene=sc, wave_size=16, fallback_wave=32, native_loop=native)
    steps = 5
    gen = pgb.run_in_batches(-(-P // steps))
    marks, graph = [], None
    while graph is None:
        try:
            marks.append(next(gen))
        except StopIteration as done:
            graph = done.value
    assert graph.edges.tobytes() == g_ref.edges.tobytes() and pgb.log.tobytes() == ref.log.tobytes()
    assert len(marks) == steps - 1 and all(b > a for a, b in zip(marks, marks[1:]))
    ref.close()
    pgb.close()
