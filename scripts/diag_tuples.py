"""Diagnose verdict mismatches between a GPU run's log and the oracle: python scripts/diag_tuples.py CONFIG"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import pgo_oracle as O  # noqa: E402
from pose_graph_initialization_b200 import builder as B  # noqa: E402
from pose_graph_initialization_b200 import scene as S  # noqa: E402
from pose_graph_initialization_b200.verify import compare_tuples, verifiable_positions  # noqa: E402

name = sys.argv[1]
sc = S.make_scene(**S.CONFIGS[name])
pgb = B.PoseGraphBuilder(kCoreNumber_=16, kSimilarityThreshold_=0.0, scene=sc)
pgb.run()
log = pgb.log
rep = compare_tuples(O, sc, log, verifiable_positions(log))
print(json.dumps({k: v for k, v in rep.items()}))
eng = pgb.engine
extra = [int(x) for x in sys.argv[2:]]  # positions to dump whatever the comparison says
if extra:
    keep = {}
    for pos in extra:
        lg = log[pos]
        corr, thr = S.pair_correspondences(sc, int(lg["pairIndex"]), 0.4)
        keep["corr_%d" % pos] = corr
        keep["thr_%d" % pos] = thr
        keep["hyp_%d" % pos] = lg["hyp"].copy()
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tuples_hard_%s.npz" % name), **keep)
for pos in list(rep["bad_positions"][:6]) + extra:
    lg = log[pos]
    p = int(lg["pairIndex"])
    corr, thr = S.pair_correspondences(sc, p, 0.4)
    hyp = lg["hyp"].copy()
    print("--- position", pos, "pair", p, "hadPath", int(lg["hadPath"]), "log: testPassed", int(lg["testPassed"]), "testCount", int(lg["testCount"]),
          "branch", int(lg["branch"]), "inl", int(lg["inlierNumber"]), "committed", int(lg["committed"]))
    ok, cnt = O.test_pose(corr, hyp, 1.5 * thr, 5)
    ro = O.estimate_pose(corr, thr, [hyp] if ok else [])
    print("oracle: test", ok, cnt, "branch", ro["branch"], "inl", ro["inlier_number"], "path_inliers", ro["path_inliers"], "success", ro["success"])
    v = eng.run_wave(np.array([p], dtype=np.uint32), np.array([0, 1], dtype=np.uint32), hyp[None, :], flags=3)[0]
    print("gpu single: test", int(v["test_passed"]), int(v["test_count"]), "branch", int(v["branch"]), "inl", int(v["inlier_count"]),
          "path_inliers", int(v["path_inliers"]), "status", int(v["status"]), "accepted", int(v["accepted"]))
    v2 = eng.run_wave(np.array([p], dtype=np.uint32), np.array([0, 1], dtype=np.uint32), hyp[None, :], flags=1)[0]
    print("gpu path-only: branch", int(v2["branch"]), "inl", int(v2["inlier_count"]), "status", int(v2["status"]), "E equal oracle", np.array_equal(v2["E"].reshape(3, 3), ro["E"]))
    E = O.essential_from_pose(hyp)
    idx = O.get_inliers(corr, E, 1.5 * thr)
    print("oracle getInliers", len(idx))
    if len(idx) >= 5:
        sub = corr[idx]
        rr = O.find_essential_ransac_inf(np.ascontiguousarray(sub))
        print("oracle legacy ransac:", {k: (v if not hasattr(v, "shape") else v.shape) for k, v in rr.items()} if isinstance(rr, dict) else rr)
pgb.close()
