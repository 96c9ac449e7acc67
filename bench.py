#!/usr/bin/env python
"""bench.py — pairs verified/s of the B200 hypothesis-verification path on BASELINE.json's configs[1]
(synthetic 300-view scene, 44,850 pairs, 2,000 correspondences/pair, 30 % outliers).

A "step" is one complete pass of the hot path over the scene: every queued pair goes through A* (host),
in-traversal test + getInliers + five-point + E->(R,t) vote or the robust fallback (GPU), and the
sequential commit, producing the pose graph.  `value` times the step with the correspondences already built
in HBM (registration outside the timed region); `e2e` times the same step through the public API from HOST
buffers (pinned-size H2D of the compact scene + on-device createCorrespondenceMatrix inside the timed region,
verdicts read back every wave).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2_300v]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ncu --set full on k5_fallback_score (profiles/): (dram__bytes_read + dram__bytes_write) / pairs of the launch
NCU_K5_DRAM_BYTES_PER_PAIR_LAUNCH = 164_500


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(name):
    from pose_graph_initialization_b200 import scene as S
    return S.make_scene(**S.CONFIGS[name])


def scene_outlier_ratio(name):
    from pose_graph_initialization_b200 import scene as S
    return S.CONFIGS.get(name, {}).get("outlier_ratio")


def dense_sample(scene, pair_ids, thr_px=0.4):
    from pose_graph_initialization_b200 import scene as S
    corr, thr, off = [], [], [0]
    for p in pair_ids:
        c, t = S.pair_correspondences(scene, int(p), thr_px)
        corr.append(c); thr.append(t); off.append(off[-1] + len(c))
    return np.vstack(corr), np.array(off, dtype=np.uint64), np.array(thr)


def cpu_leg(scene, n_pairs, threads=0, seed=0):
    """The oracle (CPU restatement of the reference path: estimatePose = fallback + E->(R,t) vote, the branch the
    synthetic scenes take for >90 % of the pairs) on a bounded, evenly spaced sample of the scene's pairs, all
    host threads.  Only place besides tests/ and smoke() where oracle/ is executed."""
    from oracle import pgo_oracle as O
    O.build()
    P = len(scene["pair_views"])
    ids = np.unique(np.linspace(0, P - 1, n_pairs).astype(np.int64))
    corr, off, thr = dense_sample(scene, ids)
    t0 = time.perf_counter()
    r = O.estimate_pose_batch(corr, off, thr, np.zeros((len(ids), 7)), np.zeros(len(ids), dtype=np.uint8), 20, threads)
    dt = time.perf_counter() - t0
    acc = int(r["info"][:, 0].sum())
    return dict(pairs=len(ids), seconds=dt, pairs_per_s=len(ids) / dt, cores=int(r["threads"]), accepted=acc,
                corr=int(off[-1]))


def _cv2_usac_one(job):
    import cv2
    cv2.setNumThreads(1)
    c, thr = job
    E, mask = cv2.findEssentialMat(c[:, :2].copy(), c[:, 2:].copy(), np.eye(3), cv2.USAC_MAGSAC, 0.99, thr)
    return 0 if mask is None else int(mask.sum())


def cv2_yardstick(scene, n_pairs, workers):
    """Third-party yard-stick of SURVEY 8(d): the fallback call site's own library call, cv2.findEssentialMat(USAC_MAGSAC)
    from the cv2 wheel of this image, over a process pool of `workers` (fallback stage only: no E->(R,t) vote, no A*).
    Runs in a fresh child process (this one holds CUDA/NCCL threads) with a hard timeout; None if unavailable."""
    import tempfile
    from pose_graph_initialization_b200 import scene as S
    P = len(scene["pair_views"])
    ids = np.unique(np.linspace(0, P - 1, n_pairs).astype(np.int64))
    jobs = [S.pair_correspondences(scene, int(p), 0.4) for p in ids]
    try:
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "jobs.npz")
            np.savez(path, corr=np.stack([c for c, _ in jobs]) if len({len(c) for c, _ in jobs}) == 1 else
                     np.array([c for c, _ in jobs], dtype=object), thr=np.array([t for _, t in jobs]))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cv2-yardstick", path, "--workers", str(workers)],
                               capture_output=True, text=True, timeout=240)
        return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else None
    except Exception:
        return None


def cv2_yardstick_child(path, workers):
    import multiprocessing as mp
    try:
        import cv2
    except Exception:
        return 1
    d = np.load(path, allow_pickle=True)
    jobs = [(np.asarray(c, dtype=np.float64), float(t)) for c, t in zip(d["corr"], d["thr"])]
    with mp.get_context("fork").Pool(workers) as pool:
        pool.map(_cv2_usac_one, jobs[:workers])  # warm the workers
        t0 = time.perf_counter()
        inl = pool.map(_cv2_usac_one, jobs)
        dt = time.perf_counter() - t0
    emit({"pairs_per_s": len(jobs) / dt, "workers": workers, "pairs": len(jobs),
          "accepted": int(sum(1 for k in inl if k >= 20)),
          "what": "cv2 %s findEssentialMat(USAC_MAGSAC) only, process pool" % cv2.__version__})
    return 0


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path (oracle port: the reference cannot be compiled in this image)
    timed on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    scene = make_scene(args.config)
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or 128 * cores  # ~5 s of all-core CPU work per step
    for _ in range(args.warmup):
        cpu_leg(scene, max(cores, 8))
    tot_p, tot_s, last = 0, 0.0, None
    for _ in range(args.steps):
        last = cpu_leg(scene, sample)
        tot_p += last["pairs"]; tot_s += last["seconds"]
    v = tot_p / tot_s
    line = {
        "impl": "reference", "metric": "image_pairs_verified_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.config, "views": int(len(scene["focal"])), "pairs": int(len(scene["pair_views"])),
                   "corr_per_pair": int(scene["m_offset"][1] - scene["m_offset"][0])},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": last["cores"], "kind": "port",
                         "sample": f"{sample} evenly spaced pairs of the scene per step, estimatePose (fallback + E->(R,t) vote), "
                                   f"{last['cores']} std::thread workers"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    """Exactly one JSON line on the real stdout (NCCL/torch banners are diverted to stderr, see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    # keep stdout clean for the one-line contract: everything else this process (or NCCL) prints goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg2_300v")
    ap.add_argument("--wave", type=int, default=0, help="queue positions per speculative wave (0 = 256 on one GPU, 512 with several ranks: every round of a wave costs a record exchange there)")
    ap.add_argument("--window", type=int, default=0, help="host re-search window (0 = library default)")
    ap.add_argument("--no-overlap", action="store_true", help="prefetch the fallback before the waves instead of concurrently")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the cpu_baseline sample (0 = 256 x cores, ~10-15 s)")
    ap.add_argument("--fb-wave", type=int, default=1024, help="pairs per fallback prefetch wave")
    ap.add_argument("--fb-streams", type=int, default=2, help="background contexts running prefetch waves concurrently")
    ap.add_argument("--cv2-yardstick", default="", help=argparse.SUPPRESS)
    ap.add_argument("--workers", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--lazy", action="store_true", help="run the fallback lazily inside the waves instead of prefetching it")
    args = ap.parse_args()
    if args.cv2_yardstick:  # child mode of cv2_yardstick(): no torch, no CUDA
        sys.exit(cv2_yardstick_child(args.cv2_yardstick, args.workers))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.wave <= 0:
        args.wave = 256 if world == 1 else 512
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pose_graph_initialization_b200 import builder as B

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU implementation (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scene = make_scene(args.config)  # same seed on every rank => identical scene
    # the step's host inputs live in PINNED memory (the e2e leg copies them host->device every step)
    for key in ("matches", "kp", "sim", "pair_views", "m_offset", "kp_offset", "focal", "size"):
        src = np.ascontiguousarray(scene[key])
        pinned = torch.empty(src.nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(src.dtype).reshape(src.shape)
        pinned[...] = src
        scene[key] = pinned
    P = len(scene["pair_views"])
    n_corr = int(scene["m_offset"][1] - scene["m_offset"][0])
    pgb = B.PoseGraphBuilder(kCoreNumber_=max(1, (os.cpu_count() or 1) // world),  # ranks share the box's cores
                              kSimilarityThreshold_=0.0, scene=scene, device=local_rank,
                             wave_size=args.wave, prefetch_fallback=not args.lazy, overlap_fallback=not args.no_overlap, research_window=args.window, fallback_wave=args.fb_wave, prefetch_streams=args.fb_streams, group=group, rank=rank,
                             world_size=world)
    pgb.prepare()
    fp64_peak = pgb.engine.fp64_peak(fused=False)
    fp64_peak_fma = pgb.engine.fp64_peak(fused=True)

    # ---- `value`: inputs resident in HBM (registration done once, outside the timed region) -------------------
    for _ in range(args.warmup):
        pgb.run()
    pgb.reset_engine_stats()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    counters = None
    step_wall = {"resident": [], "e2e": []}
    for _ in range(args.steps):
        ts = time.perf_counter()
        graph = pgb.run()
        step_wall["resident"].append(round((time.perf_counter() - ts) * 1e3, 1))
        counters = pgb.counters
    barrier()
    ev1.record()
    ev1.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    ms_resident = ev0.elapsed_time(ev1)
    st = pgb.engine_stats()
    timing = dict(pgb.timing)

    # ---- `e2e`: same step from host buffers through the public API (register_scene H2D + K0 inside) -----------
    pgb.reset_engine_stats()
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        ts = time.perf_counter()
        pgb.prepare()
        tp = time.perf_counter()
        graph = pgb.run()
        n_edges = graph.numEdges()
        step_wall["e2e"].append(round((time.perf_counter() - ts) * 1e3, 1))
        step_wall.setdefault("e2e_detail", []).append(
            {"prepare_ms": round((tp - ts) * 1e3, 1), **{k: round(v * 1e3, 1) for k, v in pgb.timing.items() if k.endswith("_s")}})
    barrier()
    ev3.record()
    ev3.synchronize()
    ms_e2e = ev2.elapsed_time(ev3)
    st_e2e = pgb.engine_stats()

    t = torch.tensor([ms_resident, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])
    K = max(args.steps, 1)
    value = P * K / (ms_resident * 1e-3)
    e2e = P * K / (ms_e2e * 1e-3)

    # ---- K1 on full-size waves (HBM roofline probe): every local pair scored against one hypothesis ---------------
    # In the pipeline K1 only sees the few hundred hypotheses of a round (launch-latency bound); its bandwidth
    # behaviour is measured here on waves of 8192 pairs x 2000 rows (0.5 GB each, inputs larger than L2).
    eng = pgb.engine
    n_local = eng.n_pairs
    ident = np.tile(np.array([0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0]), (8192, 1))
    def k1_pass():
        for s0 in range(0, n_local, 8192):
            ids = np.arange(s0, min(n_local, s0 + 8192), dtype=np.uint32)
            eng.run_wave(ids, np.arange(len(ids) + 1, dtype=np.uint32), ident[:len(ids)], flags=B.WAVE_PATH)
    k1_pass()
    eng.reset_stats()
    k1_pass()
    st_k1 = eng.stats()

    # ---- K5/K4 on one fallback wave with nothing else on the GPU: in the step three contexts overlap (wave loop + two
    # prefetch streams), so the per-stage CUDA-event brackets there contain each other's kernels; the roofline's launch
    # durations are taken from this isolated wave of the same pairs instead (same launches: 8 x (K4, K5) + K3).
    # (1776 pairs = 4 x 148 SMs x 3 resident K5 CTAs: the grid is a whole number of resident waves, as the two prefetch
    # streams of the step fill each other's partial waves)
    fb_ids = np.arange(min(n_local, 1776), dtype=np.uint32)
    eng.run_wave(fb_ids, None, None, flags=B.WAVE_FALLBACK)
    eng.reset_stats()
    v_fb = eng.run_wave(fb_ids, None, None, flags=B.WAVE_FALLBACK)
    st_fb = eng.stats()
    it = np.sort(v_fb["iters"].astype(np.int64))  # iterations of the robust loop per pair (report next to the rate)
    fb_iters = {"pairs": int(len(it)), "mean": float(it.mean()), "p10": int(it[len(it) // 10]), "median": int(it[len(it) // 2]),
                "p90": int(it[(len(it) * 9) // 10]), "max": int(it[-1]),
                "models_per_pair": st_fb["fallback_models"] / max(1, st_fb["fallback_pairs"])}

    # ---- roofline of the dominant kernel (per-stage CUDA-event times from the engine's own stream) -----------
    stages = {"k1_score_hypotheses": st["ms_score"], "k2_fivept_first_solution": st["ms_fivept"],
              "k4_fallback_solve": st["ms_fallback_solve"], "k5_fallback_score": st["ms_fallback_score"],
              "k3_decompose_vote": st["ms_decompose"]}
    dominant = max(stages, key=stages.get)
    hbm_peak, hbm_src = load_peaks()
    # K5: every scored model evaluates the Sampson residual of all N correspondences: 33 FP64 flops each (algorithmic)
    k5_flops = st_fb["fallback_models"] * n_corr * 33.0
    k5_t = st_fb["ms_fallback_score"] * 1e-3
    k5_launches = 8
    roof_k5 = {"kernel": "k5_fallback_score", "bound": "fp64", "achieved": k5_flops / k5_t / 1e12 if k5_t > 0 else 0.0,
               "peak": fp64_peak, "unit": "TFLOP/s",
               # dram__bytes_read+write per launch from profiles/ (ncu --set full, 1184 pairs x 125 iterations): ~195 MB;
               # algorithmic bytes per launch = pairs x N x 32 B read once
               "traffic": NCU_K5_DRAM_BYTES_PER_PAIR_LAUNCH * len(fb_ids),
               # FP64 rows staged once per CTA + the (72 B FP64 + 48 B FP32) model records of the chunk's iterations
               "algorithmic_bytes_per_launch": int(len(fb_ids) * (
                   n_corr * 32 + 120.0 * st_fb["fallback_models"] / max(1, st_fb["fallback_pairs"] * 8))),
               "launches": k5_launches, "ms_per_launch": st_fb["ms_fallback_score"] / k5_launches,
               "timed_on": "one isolated %d-pair fallback wave after the timed region (in the step three contexts overlap and "
                           "their CUDA-event brackets contain each other's kernels)" % len(fb_ids),
               "peak_source": "measured live: DMUL+DADD chains (parity forbids FMA); DFMA peak %.1f TFLOP/s; "
                              "achieved counts 33 algorithmic FP64 flops per model x correspondence evaluation, most of "
                              "which are certified in FP32 (see DESIGN.md section 3)" % fp64_peak_fma}
    roof_k5["frac"] = roof_k5["achieved"] / fp64_peak if fp64_peak > 0 else None
    # K1: 32 B per hypothesis x correspondence evaluation, measured on the full-size probe waves above
    k1_bytes = st_k1["corr_evals"] * 32.0
    k1_t = st_k1["ms_score"] * 1e-3
    roof_k1 = {"kernel": "k1_score_hypotheses", "bound": "hbm", "achieved": k1_bytes / k1_t / 1e9 if k1_t > 0 else 0.0,
               "peak": hbm_peak, "unit": "GB/s", "traffic": None,
               "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs); probe: %d pairs x %d rows per pass, waves of 8192 pairs" % (n_local, n_corr),
               "gcorr_evals_per_s": st_k1["corr_evals"] / k1_t / 1e9 if k1_t > 0 else 0.0}
    roof_k1["frac"] = roof_k1["achieved"] / hbm_peak
    roofline = dict(roof_k5 if dominant in ("k5_fallback_score", "k4_fallback_solve") else roof_k1)
    roofline["share_of_gpu_time"] = stages[dominant] / max(sum(stages.values()), 1e-9)
    roofline["dominant"] = dominant

    line = None
    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = cpu_leg(scene, args.cpu_sample or 256 * cores)  # ~10-15 s of all-core CPU work
        yard = cv2_yardstick(scene, 8 * cores, cores)  # ~1-2 s
        line = {
            "metric": "image_pairs_verified_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_resident / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.config, "views": int(len(scene["focal"])), "pairs": int(P), "corr_per_pair": n_corr,
                       "outlier_ratio": scene_outlier_ratio(args.config), "wave": args.wave, "fallback": "lazy" if args.lazy else "prefetched",
                       "l2": "inputs (%.1f GB of FP64 correspondences) larger than the 126 MB L2" % (P * n_corr * 32 / 1e9),
                       "parallelism": "pairs sharded over %d rank(s), verdict all-gather" % world},
            "step_wall_ms": step_wall,
            "e2e": {"value": e2e, "unit": "pairs/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] // K), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] // K)},
            "gpu_launches": int(st["launches"]),
            "gcorr_evals_per_sec": {"path_hypotheses": st["corr_evals"] / (ms_resident * 1e-3) / 1e9,
                                    "fallback_models": st["fallback_models"] * n_corr / (ms_resident * 1e-3) / 1e9},
            "roofline": roofline, "roofline_k1_scoring": roof_k1, "roofline_k5_fallback": roof_k5,
            "gpu_stage_ms_per_step": {k: v / K for k, v in stages.items()},
            "gpu_stage_ms_note": "CUDA-event brackets per context; with the prefetch overlapping the waves the brackets of "
                                 "concurrent contexts include each other's kernels (sums exceed the step's GPU time)",
            "gpu_stage_ms_isolated_fallback_wave": {"pairs": int(len(fb_ids)), "k4_fallback_solve": st_fb["ms_fallback_solve"],
                                                    "k5_fallback_score": st_fb["ms_fallback_score"],
                                                    "k3_decompose_vote": st_fb["ms_decompose"]},
            "host_s_per_step": {k: timing.get(k) for k in ("prefetch_s", "waves_s", "total_s", "engine_s", "exchange_s", "host_s",
                                                            "wait_prefetch_s", "engine_rounds", "exchanges")},
            "search_stats": getattr(pgb, "search_stats", None), "host_counters": counters, "edges": int(n_edges), "wall_s_resident": wall,
            "cpu_baseline": {"value": cpu["pairs_per_s"], "unit": "pairs/s", "cores": cpu["cores"], "kind": "port",
                             "sample": "%d evenly spaced pairs of the same scene through the oracle's estimatePose "
                                       "(fallback + E->(R,t) vote), %.1f s" % (cpu["pairs"], cpu["seconds"]),
                             "third_party_yardstick": yard},
            "fallback_iterations": fb_iters,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
