"""Device A* (K6) vs the host pool on a dense synthetic view graph with fake verdicts (tests/fake_verdicts.py):
python scripts/astar_bench.py VIEWS WAVE MAX_POSITIONS [host|gpu|check]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from fake_verdicts import dense_scene, drive  # noqa: E402

from pose_graph_initialization_b200 import builder as B  # noqa: E402

views, wave, maxpos = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else "gpu"
import fake_verdicts  # noqa: E402

ring = os.environ.get("RING", "1") != "0"  # similarity structure of the benchmark scenes
if ring:
    # inlier ratios of the 40 %-outlier benchmark scene (cfg3 log): fallback 0.29-0.35, path branch 0.06-0.25 (a composed
    # hypothesis passes the 5-inlier test long before it is accurate, and its edge carries the path inliers' count)
    fake_verdicts.FB_SCORE[:] = [0.29, 0.06]
    fake_verdicts.PATH_SCORE[:] = [0.06, 0.19]
sc = dense_scene(views, seed=3, ring_cameras=ring)
cfg = dict(similarity_threshold=0.0, minimum_inlier_number=20, minimum_point_number=50, maximum_search_depth=5,
           traversal_heuristics_weight=0.8, use_path_finding=True)
if mode == "check":
    os.environ["PGB_SEARCH_CHECK"] = "1"
host = B.HostBuilder(sc, host_threads=0, lazy_fallback=False, **cfg)
eng = None
if mode != "host":
    from pose_graph_initialization_b200 import Engine

    eng = Engine(device=0)
    host.set_search_backend(eng, min_batch=int(os.environ.get("MIN_BATCH", "16")))
t0 = time.perf_counter()
rounds = drive(host, wave, 2000, maxpos)
dt = time.perf_counter() - t0
c = host.counters()
out = dict(mode=mode, views=views, wave=wave, positions=int(c["pairs_popped"]), rounds=rounds, wall_s=dt,
           astar_runs=int(c["astar_runs"]), pops=int(c["astar_pops"]), pushes=int(c["astar_pushes"]), sec_astar=c["sec_astar"],
           sec_search_gpu=c["sec_search_gpu"], gpu_searches=int(c["gpu_searches"]), redo=int(c["gpu_search_redo"]),
           mismatches=int(c["search_mismatches"]), floor_retries=int(c["floor_retries"]), stale_spared=int(c["stale_spared"]), sec_visibility=c["sec_visibility"], sec_commit=c["sec_commit"])
if eng is not None:
    out["search_stats"] = eng.search_stats()
print(json.dumps(out))
