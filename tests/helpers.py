"""Shared synthetic two-view data for the tests (seeded; no dependency on /root/reference)."""
import numpy as np


def rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def quat_from_rot(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def two_view(n, outlier_ratio, rng, f=800.0, noise_px=0.5):
    """n normalised correspondences [x1 y1 x2 y2] of a random relative pose, `outlier_ratio` of them random.
    Returns (corr, R, t) with x2 ~ R x1 + t."""
    X = rng.standard_normal((n, 3)) + np.array([0, 0, 6.0])
    R = rodrigues(rng.uniform(-0.3, 0.3, 3))
    t = rng.standard_normal(3)
    t /= np.linalg.norm(t)
    x1 = X[:, :2] / X[:, 2:]
    X2 = X @ R.T + t
    x2 = X2[:, :2] / X2[:, 2:]
    x1 = x1 + rng.standard_normal((n, 2)) * noise_px / f
    x2 = x2 + rng.standard_normal((n, 2)) * noise_px / f
    no = int(n * outlier_ratio)
    x2[:no] = rng.uniform(-0.8, 0.8, (no, 2))
    p = rng.permutation(n)
    return np.ascontiguousarray(np.hstack([x1, x2])[p]), R, t


def pose_qt(R, t):
    return np.concatenate([quat_from_rot(R), t])


def perturbed_pose(R, t, rng, rot_sigma=1e-4, t_sigma=1e-4):
    Rp = rodrigues(rng.standard_normal(3) * rot_sigma) @ R
    tp = t + rng.standard_normal(3) * t_sigma
    return pose_qt(Rp, tp)


def two_view_keypoints(n, rng, f=800.0, w=1600, h=1200, noise_px=0.3, desc_noise=0.05, extra=0.3):
    """Pixel keypoints (float32) + 128-d float32 descriptors of n shared 3-D points seen by two cameras, plus `extra`*n
    unrelated keypoints per view, each view in its own shuffled order.  Returns dict(kp_src, desc_src, kp_dst, desc_dst,
    pose (T_dst_src as q,t), K, size, truth[src index] = dst index or -1)."""
    X = rng.standard_normal((n, 3)) * np.array([1.5, 1.0, 1.0]) + np.array([0, 0, 7.0])
    R = rodrigues(rng.uniform(-0.25, 0.25, 3))
    t = rng.standard_normal(3) * 0.8
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])
    base = rng.standard_normal((n, 128))
    base /= np.linalg.norm(base, axis=1, keepdims=True)

    def view(Xc):
        uv = (Xc[:, :2] / Xc[:, 2:]) * f + np.array([w / 2.0, h / 2.0]) + rng.standard_normal((n, 2)) * noise_px
        d = base + rng.standard_normal((n, 128)) * desc_noise
        m = int(extra * n)
        uv = np.vstack([uv, rng.uniform([0, 0], [w, h], (m, 2))])
        de = rng.standard_normal((m, 128))
        d = np.vstack([d, de / np.linalg.norm(de, axis=1, keepdims=True)])
        p = rng.permutation(n + m)
        inv = np.empty(n + m, dtype=np.int64)
        inv[p] = np.arange(n + m)
        return uv[p].astype(np.float32), d[p].astype(np.float32), inv  # inv[original] = position in the view

    kp_s, d_s, inv_s = view(X)
    kp_d, d_d, inv_d = view(X @ R.T + t)
    truth = np.full(len(kp_s), -1, dtype=np.int64)
    truth[inv_s[:n]] = inv_d[:n]
    return dict(kp_src=kp_s, desc_src=d_s, kp_dst=kp_d, desc_dst=d_d, pose=pose_qt(R, t), K=K, size=(w, h), truth=truth)
