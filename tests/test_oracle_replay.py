"""The two checks `bench.py --verify` runs at benchmark scale, exercised here on a small scene (CPU only):
(1) tuple replay — the logged (pair, hypothesis) tuples through the oracle's per-pair pipeline reproduce the logged
verdicts; (2) host replay — the oracle's sequential host, fed the logged verdicts, reproduces every host-side decision
(queue order, hasLink, A* result and touched nodes, the composed hypothesis bit for bit) and detects tampering."""
import numpy as np
import pytest

from oracle_engine import OracleEngine
from pose_graph_initialization_b200 import builder as B
from pose_graph_initialization_b200 import scene as S
from pose_graph_initialization_b200.verify import compare_tuples

CFG = dict(similarity_threshold=0.0, minimum_inlier_number=20, minimum_point_number=50, maximum_search_depth=5,
           traversal_heuristics_weight=0.8, use_path_finding=True)


@pytest.fixture(scope="module")
def logged_run(oracle):
    sc = S.make_scene(n_views=12, n_corr=200, outlier_ratio=0.3, seed=5, n_points=700)
    eng = OracleEngine(oracle, sc)
    host = B.HostBuilder(sc, host_threads=2, lazy_fallback=True, **CFG)
    while host.remaining() > 0:
        items = host.next_wave(16)
        host.commit_wave(eng.run_items(items[items["need_gpu"] > 0], path=True, fallback=True))
    return sc, host.log().copy(), host.edges().copy()


def test_host_replay_accepts_the_log_and_detects_tampering(oracle, logged_run):
    sc, log, edges = logged_run
    r = oracle.replay_scene(sc, log)
    assert r["mismatches"] == 0 and r["checked"] == len(log) and r["edges"] == len(edges) and r["searches"] > 0
    k = int(np.nonzero(log["hadPath"])[0][3])
    bad = log.copy()
    bad["hyp"][k, 2] = np.nextafter(bad["hyp"][k, 2], 1.0)  # one ulp in one hypothesis
    r = oracle.replay_scene(sc, bad)
    assert r["mismatches"] >= 1 and r["first_bad"] == k and r["first_bad_field"] == 6
    bad = log.copy()
    bad["touchedNodes"][k] += 1
    assert oracle.replay_scene(sc, bad)["first_bad_field"] == 5
    bad = log.copy()
    c = int(np.nonzero(bad["committed"])[0][5])
    bad["score"][c] *= 0.5  # a different edge score steers later searches
    assert oracle.replay_scene(sc, bad)["mismatches"] >= 1


def test_tuple_replay_reproduces_logged_verdicts(oracle, logged_run):
    sc, log, _ = logged_run
    rep = compare_tuples(oracle, sc, log, np.arange(len(log)), threads=2)
    assert rep["tuples"] > 0 and rep["mismatches"] == 0
    bad = log.copy()
    k = int(np.nonzero(bad["committed"])[0][2])
    bad["inlierNumber"][k] += 1
    assert compare_tuples(oracle, sc, bad, np.arange(len(log)), threads=2)["mismatches"] == 1
