"""PoseGraphBuilder.run_in_batches on CPU: the wave loop of builder.py (Python form: the one several ranks use) with its
prefetch worker thread, driven through a stand-in engine that answers with the pure-function verdicts of
tests/fake_verdicts.py.  A pass consumed in batches must commit exactly what run() commits, and the batches must cut the
queue where bench.py expects them (its steps are these batches)."""
import numpy as np
import pytest

import fake_verdicts as F
from pose_graph_initialization_b200 import builder as B


class FakeEngine:
    """Just enough of engine.Engine for builder.PoseGraphBuilder (no CUDA): verdicts from fake_verdicts."""

    def __init__(self, device=0, min_inliers=20, background=False):
        self.h = None
        self.n_corr = 2000
        self.pending = None
        self.n_pairs = 0

    def register_scene(self, scene, thr_px=0.4):
        self.n_pairs = len(scene["pair_views"])

    def share_pairs(self, other):
        self.n_pairs = other.n_pairs

    def _verdicts(self, pair_ids, hyp_offset, hyp, flags):
        items = np.zeros(len(pair_ids), dtype=B.ITEM_DTYPE)
        items["pair_id"] = pair_ids
        if hyp_offset is not None:
            has = np.diff(np.asarray(hyp_offset, dtype=np.int64)) > 0
            items["has_hyp"] = has
            items["hyp"][has] = np.asarray(hyp).reshape(-1, 7)
        return F.fake_verdicts(items, self.n_corr, fallback=bool(flags & B.WAVE_FALLBACK))

    def run_wave(self, pair_ids, hyp_offset=None, hyp=None, flags=3, want_masks=False):
        return self._verdicts(np.asarray(pair_ids, dtype=np.uint32), hyp_offset, hyp, flags)

    def submit_wave(self, pair_ids, hyp_offset=None, hyp=None, flags=3):
        self.pending = self._verdicts(np.asarray(pair_ids, dtype=np.uint32), hyp_offset, hyp, flags)

    def wait_wave(self):
        v, self.pending = self.pending, None
        return v

    def stats(self):
        return {}

    def reset_stats(self):
        pass

    def search_stats(self):
        return {}

    def close(self):
        pass


@pytest.mark.parametrize("steps", [1, 3, 20])
def test_a_pass_in_batches_commits_what_run_commits(monkeypatch, steps):
    monkeypatch.setattr(B._engine, "Engine", FakeEngine)
    monkeypatch.setattr(F, "FB_SCORE", [0.29, 0.06])
    monkeypatch.setattr(F, "PATH_SCORE", [0.06, 0.19])
    sc = F.dense_scene(70, n_corr=2000, seed=9, ring_cameras=True)
    P = len(sc["pair_views"])

    def make():
        return B.PoseGraphBuilder(kCoreNumber_=4, kSimilarityThreshold_=0.0, scene=sc, wave_size=64, fallback_wave=256,
                                  native_loop=False)

    ref = make()
    g_ref = ref.run()
    assert g_ref.numEdges() == P and ref.counters["path_accepted"] > 200
    pgb = make()
    batch = -(-P // steps)
    gen = pgb.run_in_batches(batch)
    marks, graph = [], None
    while graph is None:
        try:
            marks.append(next(gen))
        except StopIteration as done:
            graph = done.value
    assert graph.edges.tobytes() == g_ref.edges.tobytes()
    assert pgb.log.tobytes() == ref.log.tobytes()
    # one yield per completed batch but the last (the pass ends with the generator's return), at wave boundaries
    assert len(marks) == steps - 1
    assert all(b > a for a, b in zip(marks, marks[1:]))
    for k, m in enumerate(marks):
        assert (k + 1) * batch <= m < (k + 1) * batch + 64


def _rank_worker(rank, world, port, steps, ret):
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    import torch.distributed as dist

    import fake_verdicts as FV
    from pose_graph_initialization_b200 import builder as BB

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    FV.FB_SCORE[:] = [0.29, 0.06]
    FV.PATH_SCORE[:] = [0.06, 0.19]

    class RankEngine(FakeEngine):  # the engine of a rank is addressed with LOCAL pair ids (global id // world)
        def _verdicts(self, pair_ids, hyp_offset, hyp, flags):
            return super()._verdicts(pair_ids.astype(np.uint32) * np.uint32(world) + np.uint32(rank), hyp_offset, hyp, flags)

    BB._engine.Engine = RankEngine
    sc = FV.dense_scene(60, n_corr=2000, seed=9, ring_cameras=True)
    pgb = BB.PoseGraphBuilder(kCoreNumber_=2, kSimilarityThreshold_=0.0, scene=sc, wave_size=128, fallback_wave=128,
                              rank=rank, world_size=world)
    pgb._sub = {"pair_views": sc["pair_views"][rank::world]}  # (a host-only scene has no matches to shard; bench.py sets the shard too)
    P = len(sc["pair_views"])
    gen = pgb.run_in_batches(-(-P // steps))
    marks, graph = [], None
    while graph is None:
        try:
            marks.append(next(gen))
            dist.barrier()  # (bench.py brackets its steps with barriers)
        except StopIteration as done:
            graph = done.value
    ret[rank] = (graph.edges.tobytes(), pgb.log[[n for n in pgb.log.dtype.names if n != "touchedNodes"]].tobytes(), marks,
                 int(pgb.timing["exchanges"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_through_the_builder_loop_with_prefetch_thread_and_record_exchange(monkeypatch):
    """builder.PoseGraphBuilder on two gloo ranks (stand-in engines): sharded prefetch gathered by the worker thread over
    its own gloo group, rank-sharded wave rounds, record all-reduce — in batches, as bench.py drives it with several
    GPUs.  Both ranks must commit what a single rank commits."""
    import os

    import torch.multiprocessing as mp

    monkeypatch.setattr(B._engine, "Engine", FakeEngine)
    monkeypatch.setattr(F, "FB_SCORE", [0.29, 0.06])
    monkeypatch.setattr(F, "PATH_SCORE", [0.06, 0.19])
    sc = F.dense_scene(60, n_corr=2000, seed=9, ring_cameras=True)
    ref = B.PoseGraphBuilder(kCoreNumber_=2, kSimilarityThreshold_=0.0, scene=sc, wave_size=128, fallback_wave=128, native_loop=False)
    g_ref = ref.run()
    names = [n for n in ref.log.dtype.names if n != "touchedNodes"]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, 5, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for r in range(2):
        assert ret[r][0] == g_ref.edges.tobytes() and ret[r][1] == ref.log[names].tobytes(), r
        assert ret[r][3] > 0
    assert ret[0][2] == ret[1][2] and len(ret[0][2]) == 4  # both ranks cut their batches at the same queue positions
