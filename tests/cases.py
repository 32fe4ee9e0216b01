"""Seeded input clouds for the parity tests.  Every case is (name, points f32/f64 [N,3], labels u16|None,
num_classes, num_desired).  Used by tests/golden/make_golden.py (against the reference build, here) and by
the CPU/GPU tests (against the oracle / the CUDA path)."""
from __future__ import annotations

import numpy as np

from ndnet_b200.synth import lidar_cloud, modelnet_cloud

CUBE16 = np.array([
    [-1, 1, -1], [1, -1, -1], [1, 1, -1], [-1, -1, -1], [-1, 1, 1], [1, -1, 1], [1, 1, 1], [-1, -1, 1],
    [-.5, .5, -.5], [.5, -.5, -.5], [.5, .5, -.5], [-.5, -.5, -.5], [-.5, .5, .5], [.5, -.5, .5], [.5, .5, .5],
    [-.5, -.5, .5]], dtype=np.float64)   # core_legacy/tests/test_ndt.cpp:7-24


def _labels(n, seed, k=28):
    return np.random.default_rng(1000 + seed).integers(0, k + 1, n).astype(np.uint16)


def small_cases():
    """Cases the CPU oracle finishes in well under a second each."""
    c = []
    c.append(("lidar16k_d1000", lidar_cloud(16000, 0), None, 0, 1000))
    c.append(("lidar16k_d1000_labels", lidar_cloud(16000, 1), _labels(16000, 1), 28, 1000))
    c.append(("modelnet2048_d512", modelnet_cloud(2048, 0), None, 0, 512))
    c.append(("modelnet2048_d512_s1", modelnet_cloud(2048, 1), None, 0, 512))
    c.append(("modelnet4099_d100_tail", modelnet_cloud(4099, 2), None, 0, 100))          # N % 8 != 0 (A5)
    neg = lidar_cloud(8001, 3).copy(); neg[:, 1] = -np.abs(neg[:, 1]) - 0.25              # all-negative axis (A2)
    c.append(("lidar8001_negative_axis", neg, _labels(8001, 3), 28, 300))
    dup = modelnet_cloud(3000, 4); dup[1000:2000] = dup[:1000]                            # duplicate points
    c.append(("modelnet3000_duplicates", dup, None, 0, 200))
    flat = lidar_cloud(6000, 5).copy(); flat[:, 2] = np.round(flat[:, 2] * 0.1) * 0.05     # a few thin z layers
    c.append(("lidar6000_thin_z", flat, None, 0, 400))
    c.append(("cube16_d8", CUBE16.copy(), None, 0, 8))
    c.append(("cube16_d4", CUBE16.copy(), None, 0, 4))
    c.append(("cube16_d3", CUBE16.copy(), None, 0, 3))
    degenerate = modelnet_cloud(512, 6).copy(); degenerate[:, 2] = 0.125                   # len_z == 0 -> -3 (A3/A16)
    c.append(("plane512_degenerate", degenerate, None, 0, 64))
    c.append(("tiny7_fewer_than_workers", modelnet_cloud(7, 7), None, 0, 2))             # N < 8: nothing voxelised
    c.append(("modelnet300_d300_unreachable", modelnet_cloud(300, 8), None, 0, 300))     # needs every point alone
    lattice = (np.stack(np.meshgrid(np.arange(6), np.arange(5), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
               * 100.0).astype(np.float64)[:112]                                           # isolated voxels: walk stops (-2)
    c.append(("lattice112_isolated_prune_stop", lattice, None, 0, 100))
    # exact-integer extent on x at the first guess 14.995 (A4: points on the max face leave the grid)
    rng = np.random.default_rng(9)
    risky = rng.uniform(0.0, 1.0, (4000, 3)) * np.array([29.99, 20.0, 6.0])
    risky[::97, 0] = 29.99; risky[5, 0] = 0.0; risky[6] = [29.99, 20.0, 6.0]; risky[7] = [0.0, 0.0, 0.0]
    c.append(("risky4000_exact_extent", risky.astype(np.float64), _labels(4000, 9), 28, 30))
    # ... and one where the ACCEPTED grid is the risky one (first guess, 2x3x1 cells): every worker chunk is
    # cut at its first point on the x-max face, so membership and statistics depend on the early return
    acc = rng.uniform(0.0, 1.0, (2000, 3)) * np.array([29.99, 44.9, 14.0])
    acc[3] = [29.99, 44.9, 14.0]; acc[4] = [0.0, 0.0, 0.0]; acc[260, 0] = 29.99; acc[900, 0] = 29.99; acc[1999, 0] = 29.99
    c.append(("risky2000_accepted_grid", acc.astype(np.float64), _labels(2000, 10), 28, 5))
    return c


def medium_cases():
    """BASELINE.json-sized single clouds (oracle ~0.1-0.3 s each)."""
    return [
        ("lidar120k_d1000", lidar_cloud(120000, 10), None, 0, 1000),
        ("lidar120k_d1000_labels", lidar_cloud(120000, 11), _labels(120000, 11), 28, 1000),
        ("lidar100k_d1000", lidar_cloud(100000, 12), None, 0, 1000),
        ("lidar120003_d4096", lidar_cloud(120003, 13), _labels(120003, 13), 28, 4096),
        ("lidar120k_d256", lidar_cloud(120000, 14), None, 0, 256),
        ("lidar120k_d1024", lidar_cloud(120000, 15), None, 0, 1024),
    ]


# (cloud name, n_desired_nds, further prune targets): chains during which no prune walk passes over an already-removed
# entry, i.e. the regime in which the reference's own retained-handle continuation is well defined.
CLEAN_PRUNE_CHAINS = [
    ("lidar16k_d1000", 881, [877, 873]),
    ("lidar16k_d1000_labels", 888, [880, 866]),
    ("modelnet2048_d512", 449, [430, 405]),
    ("lidar8001_negative_axis", 265, [260, 256]),
    ("modelnet3000_duplicates", 130, [120, 110]),
]
