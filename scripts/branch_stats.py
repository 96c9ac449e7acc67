"""Branch statistics of one cfg2 run by queue-position decile: visible / path found / test passed / path accepted."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pose_graph_initialization_b200 import builder as B, scene as S

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_300v"
sc = S.make_scene(**S.CONFIGS[cfg])
pgb = B.PoseGraphBuilder(kCoreNumber_=os.cpu_count(), kSimilarityThreshold_=0.0, scene=sc)
pgb.run()
lg = pgb.log
n = len(lg)
print("positions", n, "visible", int(lg["visible"].sum()), "hadPath", int(lg["hadPath"].sum()), "testPassed",
      int(lg["testPassed"].sum()), "branch1", int((lg["branch"] == 1).sum()), "branch2", int((lg["branch"] == 2).sum()))
for d in range(10):
    s = slice(d * n // 10, (d + 1) * n // 10)
    print(d, "visible %.3f hadPath %.3f testPassed %.3f path-accepted %.3f" % (
        lg["visible"][s].mean(), lg["hadPath"][s].mean(), lg["testPassed"][s].mean(), (lg["branch"][s] == 1).mean()))
pgb.close()
