"""TEST-ONLY: synthetic pgi_verdict records that are pure functions of (pair, hypothesis) — no geometry, no oracle —
so the speculative-wave host and the device A* (K6) can be exercised on large dense view graphs in seconds.
About a third of the hypotheses are "path-accepted" with a score near 1, everything else is "fallback-accepted"
with a score near 0.6: the branch mix of the benchmark scenes."""
import numpy as np

from pose_graph_initialization_b200.engine import VERDICT_DTYPE


def _mix(x):
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xff51afd7ed558ccd)
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xc4ceb9fe1a85ec53)
    return x ^ (x >> np.uint64(33))


def _unit_pose(h):
    """Deterministic unit quaternion + translation from 64-bit hashes (n,)."""
    n = len(h)
    out = np.zeros((n, 7))
    for k in range(7):
        with np.errstate(over="ignore"):
            hk = _mix(h + np.uint64((0x9e3779b97f4a7c15 * (k + 1)) & 0xFFFFFFFFFFFFFFFF))
        out[:, k] = (hk >> np.uint64(11)).astype(np.float64) / float(1 << 53) - 0.5
    out[:, :4] /= np.linalg.norm(out[:, :4], axis=1, keepdims=True)
    return out


FB_SCORE = [0.55, 0.10]  # fallback inlier ratio: base, spread (module-level knob of the benchmark script)
PATH_SCORE = []          # path-branch inlier ratio [base, spread]; empty: ~1.0 (the benchmark script sets the measured mix)
FB_REJECT = [0]          # [k]: the fallback REJECTS every pair whose hash is 0 mod k (0: never) — sparse-scene behaviour: an edge
                         # then exists only if a path hypothesis is accepted, i.e. predictions change by edges APPEARING


def fake_verdicts(items, n_corr, path_ratio=3, fallback=True):
    """items: ITEM_DTYPE records with need_gpu set.  fallback=True (lazy-fallback waves): a failed/absent path
    hypothesis runs the fallback (status bit 0), whose verdict depends on the pair alone.  fallback=False (prefetched
    fallback verdicts, PATH-only waves): a failed hypothesis yields a non-accepted record."""
    n = len(items)
    v = np.zeros(n, dtype=VERDICT_DTYPE)
    with np.errstate(over="ignore"):
        pid = items["pair_id"].astype(np.uint64)
        hb = np.ascontiguousarray(items["hyp"]).view(np.uint64).reshape(n, 7)
        hh = _mix(pid + np.uint64(12345))
        for k in range(7):
            hh = _mix(hh ^ hb[:, k])
        hp = _mix(pid * np.uint64(0x2545F4914F6CDD1D) + np.uint64(777))
    has = items["has_hyp"] > 0
    path_ok = has & (hh % np.uint64(path_ratio) == 0)
    v["pair_id"] = items["pair_id"]
    v["n_corr"] = n_corr
    v["n_hypotheses"] = has
    v["test_passed"] = path_ok
    v["test_count"] = np.where(path_ok, 5, (hh % np.uint64(5)).astype(np.uint32))
    v["accepted"] = 1
    v["branch"] = np.where(path_ok, 1, 2)
    fb_rejected = np.zeros(n, dtype=bool)
    if FB_REJECT[0]:
        fb_rejected = ~path_ok & (hp % np.uint64(FB_REJECT[0]) == 0)
    inl_path = n_corr - (hh % np.uint64(max(1, n_corr // 100))).astype(np.uint32)
    if PATH_SCORE:
        inl_path = (PATH_SCORE[0] * n_corr + (hh % np.uint64(max(1, int(n_corr * PATH_SCORE[1]))))).astype(np.uint32)
    inl_fb = (FB_SCORE[0] * n_corr + (hp % np.uint64(max(1, int(n_corr * FB_SCORE[1]))))).astype(np.uint32)
    v["inlier_count"] = np.where(path_ok, inl_path, inl_fb)
    v["path_inliers"] = np.where(path_ok, inl_path, 0)
    v["status"] = np.where(path_ok, 0, 1)
    pose = np.where(path_ok[:, None], _unit_pose(hh), _unit_pose(hp))
    v["q"], v["t"] = pose[:, :4], pose[:, 4:]
    v["E"] = 0.0
    v["accepted"][fb_rejected] = 0
    v["branch"][fb_rejected] = 0
    v["inlier_count"][fb_rejected] = 0
    if not fallback:
        failed = ~path_ok
        v["accepted"][failed] = 0
        v["branch"][failed] = 0
        v["inlier_count"][failed] = 0
        v["status"][failed] = 0
    return v


def dense_scene(n_views, n_corr=2000, seed=0, decimals=3, ring_cameras=False):
    """A host-only scene (similarities + pair list, no keypoints): every i<j pair queued.  ring_cameras=True takes the
    similarity matrix of the benchmark scenes' generator (scene.make_scene: cosine of the viewing directions of
    cameras on a ring, half of the pairs at 0)."""
    if ring_cameras:
        from pose_graph_initialization_b200 import scene as S

        sc = S.make_scene(n_views=n_views, n_corr=1, seed=seed, n_points=64)
        pv = sc["pair_views"]
        mo = (np.arange(len(pv) + 1, dtype=np.uint64) * np.uint64(n_corr))
        return dict(sim=sc["sim"], pair_views=pv, m_offset=mo, focal=sc["focal"])
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, 2 * np.pi, n_views)
    d = np.abs(ang[:, None] - ang[None, :])
    d = np.minimum(d, 2 * np.pi - d)
    sim = np.round(np.clip(np.cos(d / 2) * 0.95 + rng.uniform(-0.02, 0.02, (n_views, n_views)), 0, 0.999), decimals)
    sim = np.minimum(sim, sim.T)
    np.fill_diagonal(sim, 1.0)
    iu = np.triu_indices(n_views, 1)
    pv = np.stack(iu, axis=1).astype(np.uint32)
    mo = (np.arange(len(pv) + 1, dtype=np.uint64) * np.uint64(n_corr))
    return dict(sim=sim, pair_views=pv, m_offset=mo, focal=np.full(n_views, 800.0))


def drive(host, wave, n_corr, max_positions=None, prefetch=True):
    """Python wave loop over fake verdicts.  prefetch=True installs every pair's fallback verdict first (the product's
    default), so waves are `wave` positions long and are iterated to their fixed point.  Returns the engine rounds."""
    from pose_graph_initialization_b200.builder import ITEM_DTYPE

    if prefetch:
        all_items = np.zeros(host.n_pairs, dtype=ITEM_DTYPE)
        all_items["pair_id"] = np.arange(host.n_pairs, dtype=np.uint32)
        host.set_fallback_verdicts(fake_verdicts(all_items, n_corr))
    rounds = 0
    total = host.remaining()
    while host.remaining() > 0:
        if max_positions is not None and total - host.remaining() >= max_positions and host.wave_status() == 0:
            break
        items = host.next_wave(wave)
        todo = items[items["need_gpu"] > 0]
        host.commit_wave(fake_verdicts(todo, n_corr, fallback=not prefetch))
        rounds += 1
    return rounds
