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
