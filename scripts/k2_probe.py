"""K2 latency probe: single-pair PATH waves (hypothesis = slightly perturbed ground truth, so the in-traversal test
passes and the legacy five-point RANSAC runs) over many pairs of a cfg2-like scene; prints the distribution of the
per-wave K2 (five-point) and K3 times and of the whole round trip — what one engine round of the wave loop costs at
least."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import perturbed_pose  # noqa: E402

from pose_graph_initialization_b200 import Engine  # noqa: E402
from pose_graph_initialization_b200 import scene as S  # noqa: E402

sc = S.make_scene(n_views=40, n_corr=2000, outlier_ratio=0.4, seed=3)
eng = Engine(device=0)
eng.register_scene(sc, 0.4)
rng = np.random.default_rng(0)
P = len(sc["pair_views"])
k2, k3, rt = [], [], []
for wave in (1, 32, 256):
    k2w, k3w, rtw, acc = [], [], [], 0
    for rep in range(60 if wave > 1 else 300):
        ids = rng.choice(P, wave, replace=False).astype(np.uint32)
        hyp = np.zeros((wave, 7))
        for i, p in enumerate(ids):
            s, d = (int(x) for x in sc["pair_views"][p])
            R, t = S.relative_gt(sc, s, d)
            hyp[i] = perturbed_pose(R, t / np.linalg.norm(t), rng, 2e-3, 2e-3)
        eng.reset_stats()
        t0 = time.perf_counter()
        v = eng.run_wave(ids, np.arange(wave + 1, dtype=np.uint32), hyp, flags=1)
        rtw.append((time.perf_counter() - t0) * 1e3)
        st = eng.stats()
        k2w.append(st["ms_fivept"]); k3w.append(st["ms_decompose"]); acc += int((v["branch"] == 1).sum())
    q = lambda a: [round(float(x), 3) for x in np.quantile(a, [0.1, 0.5, 0.9, 0.99, 1.0])]
    print(json.dumps(dict(wave=wave, path_accepted_share=acc / (wave * len(k2w)), k2_ms_p10_50_90_99_max=q(k2w), k3_ms=q(k3w), round_trip_ms=q(rtw))))
