"""Where does PoseGraphBuilder.prepare() (the e2e leg's extra work) spend its time?"""
import sys, time
import numpy as np
import torch
from pose_graph_initialization_b200 import builder as B, scene as S, engine as E

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_300v"
sc = S.make_scene(**S.CONFIGS[cfg])
for k in ("kp", "matches"):
    t = torch.from_numpy(np.ascontiguousarray(sc[k])).pin_memory()
    sc[k] = t.numpy()
eng = E.Engine(device=0)
for it in range(3):
    t0 = time.perf_counter(); eng.register_scene(sc, 0.4); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("register_scene %.1f ms" % ((t1 - t0) * 1e3), eng.stats() if it == 0 else "")
fb = E.Engine(device=0, background=True)
t0 = time.perf_counter(); fb.share_pairs(eng); t1 = time.perf_counter()
print("share_pairs %.1f ms" % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); h = B.HostBuilder(sc); t1 = time.perf_counter()
print("HostBuilder create %.1f ms" % ((t1 - t0) * 1e3))
