"""Synthetic point clouds for parity tests and benchmarks.

There are no datasets in the container, so inputs are generated: surface-like clouds (noisy
union of random planes and spheres, the occupancy pattern of a ScanNet crop) or uniform-in-ball
clouds (worst-case occupancy), normalised to the unit ball the way the reference's loaders do
(utils/utils.py:47-61 ``normalize_point_cloud``), float32, 4th channel w = 1.0.
"""
import numpy as np


def normalize_point_cloud(pc):
    """Mean-shift, then scale so that the farthest point has norm 1 (utils/utils.py:47-61)."""
    pc = pc - np.mean(pc, axis=0)
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    return pc / np.maximum(1e-5, m)


def _surface_points(rng, n):
    """n points on 6 random planes + 2 spheres with small normal noise."""
    parts = []
    per = [n // 8] * 8
    per[0] += n - sum(per)
    for s in range(6):
        normal = rng.normal(size=3)
        normal /= np.linalg.norm(normal)
        a = np.cross(normal, rng.normal(size=3))
        a /= np.linalg.norm(a)
        b = np.cross(normal, a)
        origin = rng.uniform(-0.5, 0.5, size=3)
        uv = rng.uniform(-0.8, 0.8, size=(per[s], 2))
        pts = origin + uv[:, :1] * a + uv[:, 1:] * b + rng.normal(scale=0.004, size=(per[s], 1)) * normal
        parts.append(pts)
    for s in range(6, 8):
        c = rng.uniform(-0.4, 0.4, size=3)
        rad = rng.uniform(0.15, 0.35)
        d = rng.normal(size=(per[s], 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        parts.append(c + d * (rad + rng.normal(scale=0.004, size=(per[s], 1))))
    return np.concatenate(parts, axis=0)


def _ball_points(rng, n):
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d * rng.uniform(0, 1, size=(n, 1)) ** (1.0 / 3.0)


def make_cloud(n, seed, kind="surface", face_eps=1e-4, voxels=(0.05,), shift=1.0):
    """One (n, 4) float32 cloud.  Points closer than ``face_eps`` (in voxel units) to a voxel
    face of any of ``voxels`` are resampled so that floor() cannot depend on the compiler
    (SURVEY.md s8d); duplicates are removed."""
    rng = np.random.default_rng(seed)
    gen = _surface_points if kind == "surface" else _ball_points
    # normalise a generous sample once so that the kept subset stays inside the unit ball
    pool = normalize_point_cloud(gen(rng, int(n * 1.3) + 64)).astype(np.float32)
    keep = np.ones(len(pool), bool)
    for v in voxels:
        q = (pool.astype(np.float32) + np.float32(shift)) / np.float32(v)
        frac = q - np.floor(q)
        keep &= np.all((frac > face_eps) & (frac < 1 - face_eps), axis=1)
    pool = pool[keep]
    _, first = np.unique(pool, axis=0, return_index=True)
    pool = pool[np.sort(first)]
    assert len(pool) >= n, "synthetic pool too small"
    xyz = pool[:n]
    return np.concatenate([xyz, np.ones((n, 1), np.float32)], axis=1).astype(np.float32)


def make_batch(batch, n, seed0=0, kind="surface", voxels=(0.05,), shift=1.0):
    """(batch, n, 4) float32, cloud b generated with seed ``seed0 + b``; plus (batch, 1) int32
    ``actual_numpoints = n`` (data_loader/ggcn_gpu_scannet_loader.py:72-74,288)."""
    data = np.stack([make_cloud(n, seed0 + b, kind, voxels=voxels, shift=shift) for b in range(batch)])
    return data, np.full((batch, 1), n, np.int32)
