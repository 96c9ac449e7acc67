"""TEST-ONLY stand-in for the CUDA engine: produces pgi_verdict records with the CPU oracle so that the host
logic (speculative waves, commit, sharding, verdict exchange) can be exercised on a box without a GPU.
Lives under tests/ and is never imported by the product."""
import numpy as np

from pose_graph_initialization_b200 import scene as S
from pose_graph_initialization_b200.engine import VERDICT_DTYPE


class OracleEngine:
    def __init__(self, oracle, scene, thr_px=0.4, min_inliers=20, lo=0):
        self.o, self.scene, self.thr_px, self.min_inliers, self.lo = oracle, scene, thr_px, min_inliers, lo
        self.calls = 0

    def verdict(self, pair_id, hyp, path=True, fallback=True):
        corr, thr = S.pair_correspondences(self.scene, pair_id, self.thr_px)
        v = np.zeros(1, dtype=VERDICT_DTYPE)[0]
        v["pair_id"] = pair_id
        v["n_corr"] = len(corr)
        v["q"] = [0, 0, 0, 1]
        guesses = []
        if hyp is not None and path:
            ok, cnt = self.o.test_pose(corr, hyp, 1.5 * thr, 5)
            v["n_hypotheses"] = 1
            v["test_passed"], v["test_count"] = ok, cnt
            if ok:
                guesses = [hyp]
        if guesses or fallback:
            r = self.o.estimate_pose(corr, thr, guesses, self.min_inliers)
            if r["branch"] == 1 or fallback:
                v["accepted"], v["branch"] = r["success"], r["branch"] if r["success"] else 0
                v["inlier_count"] = r["inlier_number"]
                v["path_inliers"] = r["path_inliers"]
                if r["success"]:
                    v["E"] = r["E"].reshape(9)
                    v["q"], v["t"] = r["pose"][:4], r["pose"][4:]
                if r["branch"] != 1:
                    v["status"] = 1  # the fallback ran
            else:
                v["inlier_count"] = r["path_inliers"] if r["path_inliers"] >= 5 else 0
        self.calls += 1
        return v

    def run_items(self, items, path=True, fallback=True):
        out = np.zeros(len(items), dtype=VERDICT_DTYPE)
        for i, it in enumerate(items):
            out[i] = self.verdict(int(it["pair_id"]), it["hyp"].copy() if it["has_hyp"] else None, path, fallback)
        return out
