"""N > 1 host path on CPU: two ranks (gloo) shard the pairs by owner; each rank runs the A* searches and produces the
verdicts (oracle stand-in for the engine — tests only) for the wave positions it owns, the per-position records are
merged with one all-reduce per round, and BOTH ranks must commit the pose graph the single-process sequential oracle
commits (SPMD host)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import torch.distributed as dist

    from oracle import pgo_oracle as O
    from oracle_engine import OracleEngine
    from pose_graph_initialization_b200 import builder as B
    from pose_graph_initialization_b200 import scene as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = S.make_scene(n_views=9, n_corr=150, outlier_ratio=0.3, seed=11, n_points=500)
    P = len(sc["pair_views"])
    sub = B.shard_scene(sc, rank, world)  # pairs rank, rank + world, ... (interleaved ownership)
    mine_ids = np.arange(rank, P, world)
    assert len(sub["pair_views"]) == len(mine_ids) and int(sub["m_offset"][-1]) == len(sub["matches"])
    assert np.array_equal(sub["pair_views"], sc["pair_views"][mine_ids])
    assert np.array_equal(sub["matches"][:150], sc["matches"][int(sc["m_offset"][rank]):int(sc["m_offset"][rank]) + 150])
    eng = OracleEngine(O, sc)  # verdicts only for owned pairs
    host = B.HostBuilder(sc, similarity_threshold=0.0, host_threads=2, lazy_fallback=False)
    # sharded prefetch + one all-gather
    mine = np.zeros(len(mine_ids), dtype=B.VERDICT_DTYPE)
    for k, p in enumerate(mine_ids):
        mine[k] = eng.verdict(int(p), None, path=False, fallback=True)
    parts = B.allgather_verdicts(mine, [len(range(r, P, world)) for r in range(world)])
    allv = np.zeros(P, dtype=B.VERDICT_DTYPE)
    for r in range(world):
        allv[r::world] = parts[r]
    host.set_fallback_verdicts(allv)
    host.set_partition(rank, world)  # each rank searches and verifies only the positions it owns
    exchanges = 0
    while host.remaining() > 0:
        items = host.next_wave(64)
        status = host.wave_status()
        while status != B.WAVE_DONE:
            if status == B.WAVE_NEED_GPU:
                todo = items[items["need_gpu"] > 0]
                assert np.all(B.owner_of(todo["pair_id"], world) == rank)  # only own positions are ever requested
                host.commit_wave(eng.run_items(todo, path=True, fallback=False))
            else:
                host.import_records(B.allreduce_records(host.export_records()))
                exchanges += 1
            status = host.wave_status()
            if status != B.WAVE_DONE:
                items = host.next_wave(64)
    assert exchanges > 0
    edges = host.edges()
    olog, ostats = O.run_scene(sc, sim_threshold=0.0)
    committed = olog[olog["committed"] > 0]
    ok = (len(edges) == ostats["edges"] and np.array_equal(edges["src"], committed["src"])
          and np.array_equal(edges["q"], committed["q"]) and np.array_equal(edges["score"], committed["score"]))
    ret[rank] = (bool(ok), eng.calls, len(edges))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_commit_the_sequential_graph():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]
    assert ret[0][2] == ret[1][2] > 0
    # the work really was sharded: neither rank verified everything
    assert ret[0][1] > 0 and ret[1][1] > 0


def test_interleaved_ownership():
    from pose_graph_initialization_b200 import builder as B

    assert list(B.owner_of([0, 1, 2, 4, 5, 9], 4)) == [0, 1, 2, 0, 1, 1]
    assert list(B.local_id([0, 1, 2, 4, 5, 9], 4)) == [0, 0, 0, 1, 1, 2]
