"""Host logic on CPU: the speculative-wave host (pgb.h) must commit exactly the graph the sequential oracle
host commits — same edges in the same order with bit-identical poses/scores, same per-pair log — whatever the
wave size, and with lazy or prefetched fallback verdicts.  Verdicts come from the oracle here (tests only);
the GPU engine is substituted in tests/test_gpu_scene.py."""
import numpy as np
import pytest

from oracle_engine import OracleEngine
from pose_graph_initialization_b200 import builder as B
from pose_graph_initialization_b200 import scene as S

CFG = dict(similarity_threshold=0.0, minimum_inlier_number=20, minimum_point_number=50, maximum_search_depth=5,
           traversal_heuristics_weight=0.8, use_path_finding=True)


@pytest.fixture(scope="module")
def small_scene():
    return S.make_scene(n_views=12, n_corr=200, outlier_ratio=0.3, seed=5, n_points=700)


@pytest.fixture(scope="module")
def oracle_run(oracle, small_scene):
    return oracle.run_scene(small_scene, sim_threshold=0.0)


def drive(host, eng, wave, lazy):
    while host.remaining() > 0:
        items = host.next_wave(wave)
        todo = items[items["need_gpu"] > 0]
        v = eng.run_items(todo, path=True, fallback=lazy)
        host.commit_wave(v)


def compare_logs(plog, olog):
    assert len(plog) == len(olog)
    for f in ("src", "dst", "pairIndex", "visible", "hadPath", "testPassed", "branch", "committed", "testCount",
              "inlierNumber", "nCorr", "touchedNodes"):
        assert np.array_equal(plog[f], olog[f]), f
    for f in ("E", "q", "t", "score"):
        assert np.array_equal(plog[f], olog[f]), f


@pytest.mark.parametrize("wave,lazy", [(1, True), (7, True), (64, True), (16, False), (1000, False)])
def test_wave_host_matches_sequential_oracle(oracle, small_scene, oracle_run, wave, lazy):
    olog, ostats = oracle_run
    eng = OracleEngine(oracle, small_scene)
    host = B.HostBuilder(small_scene, host_threads=4, lazy_fallback=lazy, **CFG)
    if not lazy:
        fb = np.zeros(host.n_pairs, dtype=B.VERDICT_DTYPE)
        for p in range(host.n_pairs):
            fb[p] = eng.verdict(p, None, path=False, fallback=True)
        host.set_fallback_verdicts(fb)
    drive(host, eng, wave, lazy)
    compare_logs(host.log(), olog)
    edges = host.edges()
    assert len(edges) == ostats["edges"]
    committed = olog[olog["committed"] > 0]
    assert np.array_equal(edges["src"], committed["src"]) and np.array_equal(edges["dst"], committed["dst"])
    assert np.array_equal(edges["q"], committed["q"]) and np.array_equal(edges["score"], committed["score"])
    c = host.counters()
    assert c["committed"] == ostats["edges"] and c["path_accepted"] == ostats["path_accepted"]


def test_astar_tie_breaking_matches_oracle_on_tied_similarities(oracle):
    # coarse similarities (1 decimal) => many equal costs: the heap tie order must be inherited exactly
    sc = S.make_scene(n_views=10, n_corr=120, outlier_ratio=0.2, seed=9, n_points=500)
    sc["sim"] = np.round(sc["sim"], 1)
    np.fill_diagonal(sc["sim"], 1.0)
    sc["sim"][sc["sim"] >= 1.0] = 0.9
    np.fill_diagonal(sc["sim"], 1.0)
    olog, _ = oracle.run_scene(sc, sim_threshold=0.0)
    eng = OracleEngine(oracle, sc)
    host = B.HostBuilder(sc, host_threads=2, lazy_fallback=True, **CFG)
    drive(host, eng, 5, True)
    compare_logs(host.log(), olog)


def test_pairs_without_matches_and_small_pairs_are_skipped(oracle):
    sc = S.make_scene(n_views=8, n_corr=60, outlier_ratio=0.2, seed=2, n_points=400)
    # drop some pairs from the pair list (queued by similarity but no correspondences) and shrink one below 50
    keep = np.ones(len(sc["pair_views"]), dtype=bool)
    keep[[3, 10]] = False
    n = 60
    mo = [0]
    rows = []
    for p in range(len(keep)):
        if not keep[p]:
            continue
        m = sc["matches"][p * n:(p + 1) * n]
        if p == 5:
            m = m[:40]
        rows.append(m)
        mo.append(mo[-1] + len(m))
    sc["pair_views"] = sc["pair_views"][keep]
    sc["matches"] = np.vstack(rows)
    sc["m_offset"] = np.array(mo, dtype=np.uint64)
    olog, ostats = oracle.run_scene(sc, sim_threshold=0.0)
    eng = OracleEngine(oracle, sc)
    host = B.HostBuilder(sc, host_threads=1, lazy_fallback=True, **CFG)
    drive(host, eng, 9, True)
    compare_logs(host.log(), olog)
    assert host.counters()["skipped"] == ostats["skipped"] == 3


def test_similarity_threshold_filters_queue(oracle, small_scene):
    olog, _ = oracle.run_scene(small_scene, sim_threshold=0.6, max_pairs=0)
    eng = OracleEngine(oracle, small_scene)
    cfg = dict(CFG, similarity_threshold=0.6)
    host = B.HostBuilder(small_scene, host_threads=1, lazy_fallback=True, **cfg)
    drive(host, eng, 8, True)
    compare_logs(host.log(), olog)


@pytest.mark.parametrize("wave,lazy", [(7, True), (64, False), (1000, False)])
def test_native_wave_driver_matches_sequential_oracle(oracle, small_scene, oracle_run, wave, lazy):
    """pgb_run_wave (the C++ loop: next_wave -> engine rounds -> commit) with the engine behind the two function pointers
    it takes in production (pgi_submit_wave / pgi_wait_wave); here they are ctypes callbacks over the oracle stand-in."""
    import ctypes as C

    olog, ostats = oracle_run
    eng = OracleEngine(oracle, small_scene)
    host = B.HostBuilder(small_scene, host_threads=4, lazy_fallback=lazy, **CFG)
    if not lazy:
        fb = np.zeros(host.n_pairs, dtype=B.VERDICT_DTYPE)
        for p in range(host.n_pairs):
            fb[p] = eng.verdict(p, None, path=False, fallback=True)
        host.set_fallback_verdicts(fb)
    pending = {}

    @C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_uint32)
    def submit(_ctx, n, pair_id, hyp_offset, hyp, _flags):
        out = np.zeros(n, dtype=B.VERDICT_DTYPE)
        for i in range(n):
            h = None
            if hyp_offset[i + 1] > hyp_offset[i]:
                h = np.array([hyp[7 * hyp_offset[i] + k] for k in range(7)])
            out[i] = eng.verdict(int(pair_id[i]), h, path=True, fallback=lazy)
        pending["v"] = out
        return 0

    @C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)
    def wait(_ctx, out, _masks):
        v = pending.pop("v")
        C.memmove(out, v.ctypes.data, v.nbytes)
        return 0

    st = B.DriveStats()
    while host.remaining() > 0:
        assert host.run_wave(wave, submit, wait, None, 1, st) == B.WAVE_DONE
    assert st.rounds > 0 and st.items > 0
    compare_logs(host.log(), olog)
    assert len(host.edges()) == ostats["edges"]
    assert host.counters()["path_accepted"] == ostats["path_accepted"]


def _synthetic_verdict(sc, pair_id, path_ok, inliers):
    from helpers import quat_from_rot

    v = np.zeros(1, dtype=B.VERDICT_DTYPE)[0]
    s, d = sc["pair_views"][pair_id]
    R, t = S.relative_gt(sc, int(s), int(d))
    v["pair_id"], v["accepted"], v["branch"] = pair_id, 1, 1 if path_ok else 2
    v["inlier_count"], v["n_corr"] = inliers, 50
    v["q"], v["t"] = quat_from_rot(R), t / np.linalg.norm(t)
    v["status"], v["test_passed"] = (0 if path_ok else 1), path_ok
    return v


@pytest.mark.parametrize("threads", [1, 8])
def test_committed_graph_is_independent_of_the_wave_size_at_scale(threads):
    """4,950 pairs with synthetic verdicts that are pure functions of (pair, hypothesis bits): the committed log and
    edges must be byte-identical for every wave size (wave 1 IS the sequential semantics).  Exercises the overlay
    truncation and the visibility checkpoints of rebuildOverlay on waves far longer than the oracle-backed tests."""
    sc = S.make_scene(n_views=100, n_corr=50, outlier_ratio=0.3, seed=2, n_points=500)
    P = len(sc["pair_views"])
    fb = np.zeros(P, dtype=B.VERDICT_DTYPE)
    for p in range(P):
        fb[p] = _synthetic_verdict(sc, p, False, 30 + p % 15)
    outs = {}
    for wave in (1, 37, 256, 2048):
        host = B.HostBuilder(sc, similarity_threshold=0.0, host_threads=threads, lazy_fallback=False)
        host.set_fallback_verdicts(fb)
        while host.remaining() > 0:
            items = host.next_wave(wave)
            while host.wave_status() != B.WAVE_DONE:
                todo = items[items["need_gpu"] > 0]
                v = np.zeros(len(todo), dtype=B.VERDICT_DTYPE)
                for i, it in enumerate(todo):
                    ok = bool((int(it["pair_id"]) + int(abs(it["hyp"][4]) * 1e6)) % 3 == 0) if it["has_hyp"] else False
                    p = int(it["pair_id"])
                    v[i] = _synthetic_verdict(sc, p, ok, 20 + int(abs(it["hyp"][5]) * 1e6) % 25 if ok else 30 + p % 15)
                host.commit_wave(v)
                if host.wave_status() != B.WAVE_DONE:
                    items = host.next_wave(wave)
        lg = host.log()
        outs[wave] = (lg[[n for n in lg.dtype.names if n != "touchedNodes"]].tobytes(), host.edges().tobytes(),
                      host.counters()["path_accepted"])
    assert outs[1][2] > 500  # plenty of path-accepted edges, i.e. plenty of prediction changes inside the waves
    for wave in (37, 256, 2048):
        assert outs[wave][:2] == outs[1][:2], wave


def test_both_directions_of_a_pair_queued_in_one_wave():
    """Asymmetric similarities queue (i,j) and (j,i) (imagesimilarity_graph.h:149-157); the second one is skipped by the
    duplicate-edge check (pose_graph_builder.h:438-443) when the first was committed.  The speculative overlay must
    model that skip: the wave result may not depend on the wave size (wave = 1 is the sequential run)."""
    from fake_verdicts import dense_scene, drive

    sc = dense_scene(40, n_corr=400, seed=3)
    rng = np.random.default_rng(8)
    sim = sc["sim"].copy()
    iu = np.triu_indices(40, 1)
    pick = rng.random(len(iu[0])) < 0.5
    sim[iu[1][pick], iu[0][pick]] = np.round(np.clip(sim[iu[0][pick], iu[1][pick]] - 0.013, 0, 0.999), 3)  # lower triangle differs
    sc["sim"] = sim
    rev = np.stack([iu[1][pick], iu[0][pick]], axis=1).astype(np.uint32)
    sc["pair_views"] = np.concatenate([sc["pair_views"], rev])
    sc["m_offset"] = np.arange(len(sc["pair_views"]) + 1, dtype=np.uint64) * np.uint64(400)
    logs = []
    for wave in (1, 37, 256):
        host = B.HostBuilder(sc, host_threads=3, lazy_fallback=False, **CFG)
        drive(host, wave, 400)
        logs.append((host.log().copy(), host.edges().copy(), host.counters()))
        host.close()
    assert logs[0][2]["skipped"] > 0  # the reversed duplicates were met and skipped
    for lg, ed, _ in logs[1:]:
        assert lg.tobytes() == logs[0][0].tobytes()
        assert ed.tobytes() == logs[0][1].tobytes()


@pytest.mark.parametrize("fb_reject", [0, 4])
def test_cost_floor_and_fine_staleness_keep_the_sequential_result_on_a_dense_graph(monkeypatch, fb_reject):
    """Dense ring-camera view graph with the benchmark scenes' score mix (path-branch edges score LOWER than the fallback
    edges they replace).  Two exactness claims of pgb_host.cpp are exercised at once: searches run with a cost floor
    (children below it skip the push_heap climb) and a search result survives a changed edge at an expanded vertex when
    the affected child stays below every cost the search popped.  Wave 1 is the sequential semantics."""
    import fake_verdicts
    from fake_verdicts import dense_scene, drive

    monkeypatch.setattr(fake_verdicts, "FB_SCORE", [0.29, 0.06])
    monkeypatch.setattr(fake_verdicts, "PATH_SCORE", [0.06, 0.19])
    monkeypatch.setattr(fake_verdicts, "FB_REJECT", [fb_reject])  # 4: a quarter of the fallbacks reject (edges appear / stay absent)
    sc = dense_scene(110, n_corr=2000, seed=5, ring_cameras=True)
    logs = []
    for wave, threads in ((1, 1), (64, 4), (256, 8)):
        host = B.HostBuilder(sc, host_threads=threads, lazy_fallback=False, **CFG)
        drive(host, wave, 2000)
        logs.append((host.log().copy(), host.edges().copy(), host.counters()))
        host.close()
    c = logs[2][2]
    assert (c["rejected"] > 300) == bool(fb_reject)
    assert c["path_accepted"] > 1000 and c["astar_reruns"] > 0
    assert c["stale_spared"] > 0  # the fine rule kept results the coarse rule would have searched again
    for lg, ed, _ in logs[1:]:
        assert lg.tobytes() == logs[0][0].tobytes()
        assert ed.tobytes() == logs[0][1].tobytes()
