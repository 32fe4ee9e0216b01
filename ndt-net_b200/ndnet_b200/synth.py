"""Synthetic point clouds for tests and bench: fp32, seeded per cloud index."""
from __future__ import annotations

import numpy as np


def lidar_cloud(n_points: int, seed: int = 0, with_labels: bool = False, num_classes: int = 28):
    """LiDAR-like scan: az~U[0,2pi), el~U[-25deg,+3deg], range=min(1.5+Gamma(2,12), ground hit, 100 m)."""
    rng = np.random.default_rng(seed)
    az = rng.uniform(0.0, 2.0 * np.pi, n_points)
    el = np.deg2rad(rng.uniform(-25.0, 3.0, n_points))
    rg = 1.5 + rng.gamma(2.0, 12.0, n_points)
    ground = 1.8 / np.maximum(np.sin(-el), 1e-3)
    rg = np.minimum(np.minimum(rg, ground), 100.0)
    xyz = np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el)], axis=1)
    xyz = xyz.astype(np.float32)
    if with_labels:
        labels = rng.integers(0, num_classes + 1, n_points).astype(np.uint16)
        return xyz, labels
    return xyz


def modelnet_cloud(n_points: int = 2048, seed: int = 0):
    """ModelNet-like object: random unit directions scaled by U[0.3, 1]."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_points, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.uniform(0.3, 1.0, (n_points, 1))
    return (d * r).astype(np.float32)


def lidar_batch(batch: int, n_points: int, seed0: int = 0, with_labels: bool = False, num_classes: int = 28):
    clouds, labels = [], []
    for b in range(batch):
        out = lidar_cloud(n_points, seed0 + b, with_labels, num_classes)
        if with_labels:
            clouds.append(out[0]); labels.append(out[1])
        else:
            clouds.append(out)
    pts = np.stack(clouds)
    return (pts, np.stack(labels)) if with_labels else pts
