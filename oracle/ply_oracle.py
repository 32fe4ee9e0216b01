"""TEST INFRASTRUCTURE — CPU restatement of the reference's ASCII-PLY reader, used only as the checker in tests/
and by tools/bench_ply.py's CPU leg.  Never imported by the product (ndt-net_b200/).

Follows /root/reference/ndnet/datasets/CARLA_Seg.py:96-183 (`CARLA_Seg.get_data_pcl`) line by line; the only change
is that the random subsample's indexes (`np.random.choice(N, n_samples, replace=False)`, :141) are an argument, so
that both sides of a parity test use the same draw.  Pinned against the reference's own method by
tests/golden/make_ply_golden.py (fixture tests/golden/ply_ref_golden.npz).
"""
from __future__ import annotations

import io

import numpy as np


def read_lines(raw: bytes):
    """`open(path, 'r').readlines()` (:111-112): text mode, universal newlines."""
    return io.TextIOWrapper(io.BytesIO(raw), encoding="utf-8", newline=None).readlines()


def parse(raw: bytes, n_classes: int, num_header_lines: int = 10):
    """:115-146 — every data line: float(data[0..2]), int(data[-1]), bound check; float64 points, uint16 tags."""
    points, classes = [], []
    for point in read_lines(raw)[num_header_lines:]:
        data = point.strip().split()                  # :118
        x = float(data[0])                            # :120
        y = float(data[1])
        z = float(data[2])
        class_tag = int(data[-1])                     # :123
        if class_tag > n_classes:                     # :127
            raise ValueError(f"Class tag {class_tag} out of bounds")
        points.append(np.array([x, y, z]))            # :131
        classes.append(class_tag)                     # :134
    np_points = np.asarray(points)                    # :138
    np_classes = np.asarray(classes, dtype=np.uint16)  # :146
    return np_points, np_classes


def get_data_pcl(raw: bytes, n_classes: int, indexes, num_header_lines: int = 10):
    """:138-183 — subsample, float32 points, one-hot ground truth.  Returns (points f32 [n,3], gt f32 [n,C+1], tags)."""
    np_points, np_classes = parse(raw, n_classes, num_header_lines)
    indexes = np.asarray(indexes)
    np_points = np_points[indexes]                    # :142
    np_classes = np_classes[indexes]                  # :147
    points = np_points.astype(np.float32)             # torch.tensor(np_points).float(), :173
    gt = np.zeros((np_classes.shape[0], n_classes + 1), np.float32)   # :176
    for i in range(np_classes.shape[0]):              # :177-178
        gt[i, int(np_classes[i])] = 1
    return points, gt, np_classes
