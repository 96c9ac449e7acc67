"""TEST-ONLY driver: runs bench.main() on a CPU-only machine by replacing the CUDA pieces with stand-ins — the engine
(tests/test_builder_batches.FakeEngine with timing statistics), torch.cuda events / pinned memory, nvidia-smi — so that
the control flow of bench.py (argument handling, warm-up, the K timed steps, the arithmetic of the JSON line) is
exercised where no B200 exists.  Nothing measured here means anything; `python tests/bench_dryrun.py <bench args>`."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import fake_verdicts as F  # noqa: E402
from pose_graph_initialization_b200 import builder as B  # noqa: E402
from pose_graph_initialization_b200 import scene as S  # noqa: E402
from test_builder_batches import FakeEngine  # noqa: E402

STAT_KEYS = ("ms_correspondences", "ms_score", "ms_fivept", "ms_fallback_solve", "ms_fallback_score", "ms_decompose", "ms_total")


class StatEngine(FakeEngine):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.reset_stats()

    def register_scene(self, scene, thr_px=0.4):
        super().register_scene(scene, thr_px)
        self.n_corr = int(scene["m_offset"][1] - scene["m_offset"][0]) if len(scene["m_offset"]) > 1 else 1
        self._s["h2d_bytes"] += 1000

    def _verdicts(self, pair_ids, hyp_offset, hyp, flags):
        v = super()._verdicts(pair_ids, hyp_offset, hyp, flags)
        v["iters"] = 1000
        s = self._s
        s["launches"] += 5; s["pairs"] += len(v); s["corr_evals"] += len(v) * self.n_corr
        s["h2d_bytes"] += 56 * len(v); s["d2h_bytes"] += 160 * len(v)
        if flags & B.WAVE_FALLBACK:
            s["fallback_pairs"] += len(v); s["fallback_models"] += 4000 * len(v)
        for k in STAT_KEYS:
            s[k] += 0.01
        return v

    def fp64_peak(self, fused=False):
        return 34.6 if fused else 18.2

    def stats(self):
        return dict(self._s)

    def reset_stats(self):
        self._s = {k: 0.0 for k in STAT_KEYS}
        self._s.update(launches=0, pairs=0, corr_evals=0, fallback_pairs=0, fallback_models=0, h2d_bytes=0, d2h_bytes=0)


class FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class NoClocks:
    def __init__(self, idx):
        pass

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["dry run"], "samples": 0}


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))

    class RankEngine(StatEngine):  # a rank's engine is addressed with LOCAL pair ids (global id // world)
        def _verdicts(self, pair_ids, hyp_offset, hyp, flags):
            return super()._verdicts(np.asarray(pair_ids).astype(np.uint32) * np.uint32(world) + np.uint32(rank), hyp_offset, hyp, flags)

    B._engine.Engine = RankEngine if world > 1 else StatEngine
    if world > 1:  # several ranks: gloo instead of NCCL, CPU tensors in the exchanges
        import torch.distributed as dist

        real_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **k: real_init("gloo", **{kk: vv for kk, vv in k.items() if kk != "device_id"})
    real_builder = B.PoseGraphBuilder
    B.PoseGraphBuilder = lambda *a, **k: real_builder(*a, **{**k, "native_loop": False})  # the native driver needs the real engine
    answers = [True]  # bench.py's "is there a B200" check; builder.py asks again later and must hear the truth (CPU tensors)
    torch.cuda.is_available = lambda: answers.pop() if answers else False
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = FakeEvent
    real_empty, real_tensor = torch.empty, torch.tensor
    torch.empty = lambda *a, **k: real_empty(*a, **{kk: vv for kk, vv in k.items() if kk != "pin_memory"})
    torch.tensor = lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    bench.ClockSampler = NoClocks
    bench.cv2_yardstick = lambda *a, **k: None
    F.FB_SCORE[:] = [0.29, 0.06]
    F.PATH_SCORE[:] = [0.06, 0.19]
    S.CONFIGS["dry_24v"] = dict(n_views=24, n_corr=2000, outlier_ratio=0.3, seed=7, n_points=2500)
    bench.main()


if __name__ == "__main__":
    main()
