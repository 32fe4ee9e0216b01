"""TEST INFRASTRUCTURE ONLY - CPU restatement of the library's optional "textbook" mode (NDNET_B200_TEXTBOOK_KL,
include/ndnet_b200.h): the algorithm the reference's README documents (/root/reference/README.md:6 - "the
Kullback-Leibler divergence is computed between neighboring distributions in all directions.  The distributions with
the least divergence ... are the most redundant and are hence removed"), as opposed to what the compiled reference does
(per-step-divided off-diagonals, covariances factorised in place, a trace-only pseudo divergence,
the LARGEST divergences removed first).

PARITY UNPINNED: the reference holds no implementation, test or golden vector of this variant; this file IS its
definition.  What it shares with the legacy path - bounding box, voxel-size search, voxel assignment, the sequential mean
and m2 recurrences (/root/reference/core_legacy/src/normal_distributions.c:76-89), the 6-neighbourhood in enum order
(voxel.c:116-175), head-first removal of first occurrences (ndt.c:45-72), ascending compaction (ndt.c:75-117) - comes from
the pinned oracle (oracle/ndt_oracle.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

from . import ndt_oracle


class TextbookResult:
    pass


def _spd_det_inv(S):
    a, b, c, d, e, f = S[0], S[1], S[2], S[4], S[5], S[8]
    c00, c01, c02 = d * f - e * e, c * e - b * f, b * e - c * d
    det = a * c00 + b * c01 + c * c02
    if not (a > 0.0) or not (a * d - b * b > 0.0) or not (det > 0.0):
        return None
    r = 1.0 / det
    inv = [c00 * r, c01 * r, c02 * r, 0.0, (a * f - c * c) * r, (b * c - a * e) * r, 0.0, 0.0, (a * d - b * b) * r]
    inv[3], inv[6], inv[7] = inv[1], inv[2], inv[5]
    return det, inv


def run(cloud: np.ndarray, num_desired: int) -> TextbookResult:
    """cloud [N,3] (any float dtype, widened exactly to f64)."""
    base = ndt_oracle.run(cloud, num_desired)
    out = TextbookResult()
    out.ret, out.lens, out.voxel_size = base.ret, base.lens, base.voxel_size
    if base.ret != 0:
        return out
    x = np.ascontiguousarray(cloud, np.float64)
    pv = base.point_voxel
    cells = np.unique(pv[pv >= 0])
    slot_of = {int(c): i for i, c in enumerate(cells)}
    V = len(cells)
    n = np.zeros(V, np.int64)
    mean = np.zeros((V, 3))
    cov = np.zeros((V, 9))
    # sequential statistics per voxel, points in ascending index order
    order = np.argsort(pv, kind="stable")
    order = order[pv[order] >= 0]
    start = 0
    for s, cell in enumerate(cells):
        cnt = int(np.count_nonzero(pv == cell))
        idx = order[start:start + cnt]
        start += cnt
        mu = [0.0, 0.0, 0.0]
        m2 = [0.0, 0.0, 0.0]
        c01 = c02 = c12 = 0.0
        for k, i in enumerate(idx):
            p = [float(x[i, 0]), float(x[i, 1]), float(x[i, 2])]
            kk = float(k + 1)
            d0 = p[0] - mu[0]
            mu[0] = mu[0] + d0 / kk
            e0 = p[0] - mu[0]
            m2[0] = m2[0] + d0 * e0
            f1, f2 = p[1] - mu[1], p[2] - mu[2]
            c01 = c01 + e0 * f1
            c02 = c02 + e0 * f2
            mu[1] = mu[1] + f1 / kk
            e1 = p[1] - mu[1]
            m2[1] = m2[1] + f1 * e1
            c12 = c12 + e1 * f2
            mu[2] = mu[2] + f2 / kk
            m2[2] = m2[2] + f2 * (p[2] - mu[2])
        n[s] = cnt
        mean[s] = mu
        cn = float(cnt)
        cov[s] = [m2[0] / cn, c01 / cn, c02 / cn, c01 / cn, m2[1] / cn, c12 / cn, c02 / cn, c12 / cn, m2[2] / cn]
    lx, ly, lz = base.lens
    # divergences in insertion order: voxels ascending, directions X+,X-,Y+,Y-,Z+,Z-
    entries = []
    inv_cache = {}

    def spd(s):
        if s not in inv_cache:
            inv_cache[s] = _spd_det_inv(cov[s]) if n[s] > 1 else None
        return inv_cache[s]

    for s, cell in enumerate(cells):
        P = spd(s)
        for d in range(6):
            r, nb = ndt_oracle.neighbor(int(cell), lx, ly, lz, d)
            if r != 0 or int(nb) not in slot_of or P is None:
                continue
            q = slot_of[int(nb)]
            Q = spd(q)
            if Q is None:
                continue
            qdet, qi = Q
            Pm = cov[s]
            tr = 0.0
            for i in range(3):
                for k in range(3):
                    tr += qi[i * 3 + k] * Pm[k * 3 + i]
            dm = [mean[q][0] - mean[s][0], mean[q][1] - mean[s][1], mean[q][2] - mean[s][2]]
            maha = sum(dm[i] * (qi[i * 3 + 0] * dm[0] + qi[i * 3 + 1] * dm[1] + qi[i * 3 + 2] * dm[2]) for i in range(3))
            div = 0.5 * (tr + maha - 3.0 + math.log(qdet / P[0]))
            if math.isfinite(div):
                entries.append((div + 0.0, s * 6 + d, s, q))
    entries.sort(key=lambda t: (t[0], t[1]))           # ascending divergence, insertion order among equals
    out.kl_div = np.array([t[0] for t in entries])
    out.kl_p = np.array([int(cells[t[2]]) for t in entries], np.int64)
    out.kl_q = np.array([int(cells[t[3]]) for t in entries], np.int64)
    # head-first removal of the first V - D distinct p's
    to_remove = V - int(num_desired)
    removed = np.zeros(V, bool)
    got = 0
    for t in entries:
        if got >= to_remove:
            break
        if not removed[t[2]]:
            removed[t[2]] = True
            got += 1
    keep = [s for s in range(V) if not removed[s]][:int(num_desired)]
    out.num_voxels, out.num_removed = V, got
    out.out_voxel = np.array([int(cells[s]) for s in keep], np.int64)
    out.out_pts = mean[keep]
    out.out_cov = cov[keep]
    return out
