"""One fallback-only wave + one path wave on a slice of cfg2 — the short command profiled under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pose_graph_initialization_b200 import Engine, scene as S
from pose_graph_initialization_b200.engine import WAVE_FALLBACK, WAVE_PATH

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1184  # 8 x 148
sc = S.make_scene(n_views=60, n_corr=2000, outlier_ratio=0.3, seed=2, n_points=4000)
eng = Engine(device=0)
eng.register_scene(sc, 0.4)
ids = np.arange(min(n_pairs, eng.n_pairs), dtype=np.uint32)
for rep in range(2):
    out = eng.run_wave(ids, None, None, flags=WAVE_FALLBACK)
# a path wave: ground-truth-ish hypotheses (pass the test) so K2/K3 run
hyp = np.zeros((len(ids), 7))
for k, p in enumerate(ids):
    s, d = (int(x) for x in sc["pair_views"][p])
    R, t = S.relative_gt(sc, s, d)
    hyp[k, :4] = S._quat_from_rot(R); hyp[k, 4:] = t / np.linalg.norm(t)
hoff = np.arange(len(ids) + 1, dtype=np.uint32)
out2 = eng.run_wave(ids, hoff, hyp, flags=WAVE_PATH)
st = eng.stats()
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
import hashlib
print("verdict sha", hashlib.sha1(out.tobytes() + out2.tobytes()).hexdigest()[:12])
print("fallback accepted", int(out["accepted"].sum()), "path accepted", int((out2["branch"] == 1).sum()))
