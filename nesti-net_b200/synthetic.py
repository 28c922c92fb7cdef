"""Synthetic workloads of the benchmark and the tests: deterministic clouds of the shapes the
BASELINE configs name.  Pure numpy; no dataset or network access is needed anywhere in this repository."""
import numpy as np

F32 = np.float32


def synthetic_cloud(n_points, cloud_id=0, kind="pcpnet", noise=0.0):
    """Deterministic synthetic cloud, float32 [N,3].

    kind='pcpnet': closed smooth bumpy surface sampled ~uniformly by area
    (PCPNet-shape); kind='scan': range-scanner-like non-uniform density
    (density ~ 1/range^2) of the same surface.  Seed 1000+cloud_id."""
    rng = np.random.RandomState(1000 + cloud_id)
    if kind == "pcpnet":
        v = rng.normal(size=(n_points, 3))
    elif kind == "scan":
        # concentrate directions around +z with a heavy tail: >20x density spread
        v = rng.normal(size=(n_points, 3)) * np.array([1.0, 1.0, 0.35]) + np.array([0.0, 0.0, 0.9])
    else:
        raise ValueError("Unknown cloud kind: %s" % kind)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    th = np.arctan2(v[:, 1], v[:, 0])
    ph = np.arccos(np.clip(v[:, 2], -1, 1))
    k = 2 + (cloud_id % 5)
    rad = 1.0 + 0.18 * np.sin(k * th) * np.sin(ph) ** 2 + 0.12 * np.cos((k + 1) * ph)
    pts = v * rad[:, None] * np.array([1.0, 0.8, 0.6])
    if noise > 0:
        diag = np.linalg.norm(pts.max(0) - pts.min(0))
        pts = pts + rng.normal(size=pts.shape) * noise * diag
    return np.ascontiguousarray(pts, dtype=F32)
