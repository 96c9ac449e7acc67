"""Checks of a finished run against a CPU checker, used by `bench.py --verify` and the tests.

The checker (the oracle, tests/bench only) is passed in by the caller: nothing here imports it, and nothing on
the product path imports this module."""
import numpy as np


def verifiable_positions(log):
    """Queue positions that went through the per-pair pipeline (not skipped: they have correspondences and were
    neither duplicates nor below minimum_point_number, i.e. they carry a verdict)."""
    ran = (log["pairIndex"] >= 0) & ((log["committed"] > 0) | (log["inlierNumber"] > 0) | (log["hadPath"] > 0) |
                                     (np.abs(log["E"]).sum(axis=1) > 0))
    return np.nonzero(ran)[0]


def compare_tuples(checker, scene, log, positions, thr_px=0.4, min_inliers=20, threads=0, chunk=4096):
    """Run the logged (pair, hypothesis) tuples of `positions` through checker.scene_pipeline_batch and compare the
    verdicts with the log bit for bit: test verdict and count, branch, inlier number, E, committed pose.
    Returns dict(tuples, mismatches, first_bad, seconds, threads, branch counts)."""
    import time

    positions = np.asarray(positions, dtype=np.int64)
    positions = positions[np.isin(positions, verifiable_positions(log))]
    bad, first, secs, used = 0, -1, 0.0, 0
    bad_positions = []
    mix = {"path": 0, "fallback": 0, "rejected": 0}
    for s in range(0, len(positions), chunk):
        pos = positions[s:s + chunk]
        lg = log[pos]
        t0 = time.perf_counter()
        r = checker.scene_pipeline_batch(scene, lg["pairIndex"].astype(np.uint32), lg["hyp"], lg["hadPath"], thr_px, min_inliers, threads)
        secs += time.perf_counter() - t0
        used = r["threads"]
        info = r["info"]
        ok = np.ones(len(pos), dtype=bool)
        had = lg["hadPath"] > 0
        ok &= (r["test_passed"] > 0) == (lg["testPassed"] > 0)
        ok &= ~had | (r["test_count"] == lg["testCount"])
        success = info[:, 0] > 0
        ok &= success == (lg["committed"] > 0)
        ok &= np.where(success, info[:, 1], 0) == lg["branch"]
        ok &= info[:, 2] == lg["inlierNumber"]
        ok &= (r["E"].view(np.uint64) == np.ascontiguousarray(lg["E"]).view(np.uint64)).all(axis=1)
        pose = np.concatenate([lg["q"], lg["t"]], axis=1)
        same_pose = (r["pose"].view(np.uint64) == np.ascontiguousarray(pose).view(np.uint64)).all(axis=1)
        ok &= ~success | same_pose
        if not ok.all():
            if first < 0:
                first = int(pos[np.nonzero(~ok)[0][0]])
            bad += int((~ok).sum())
            bad_positions.extend(int(x) for x in pos[np.nonzero(~ok)[0]][:64])
        mix["path"] += int((success & (info[:, 1] == 1)).sum())
        mix["fallback"] += int((success & (info[:, 1] == 2)).sum())
        mix["rejected"] += int((~success).sum())
    return dict(tuples=int(len(positions)), mismatches=bad, first_bad=first, bad_positions=bad_positions[:64], seconds=secs,
                threads=used, **mix)
