"""CPU checks of oracle/textbook_oracle.py (the statement of the optional NDNET_B200_TEXTBOOK_KL mode; parity unpinned -
the reference has no implementation of its README's algorithm): its pieces against numpy's linear algebra and the closed
form of the Gaussian Kullback-Leibler divergence, and the selection rule on a cloud small enough to follow by hand."""
import numpy as np

from oracle import ndt_oracle, textbook_oracle
from ndnet_b200.synth import lidar_cloud


def test_inverse_and_determinant_against_numpy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.normal(size=(3, 3))
        S = a @ a.T + 1e-3 * np.eye(3)
        det, inv = textbook_oracle._spd_det_inv(S.reshape(-1))
        assert abs(det - np.linalg.det(S)) <= 1e-10 * abs(det)
        assert np.allclose(np.array(inv).reshape(3, 3), np.linalg.inv(S), rtol=1e-8, atol=1e-12)
    assert textbook_oracle._spd_det_inv(np.diag([1.0, 0.0, 1.0]).reshape(-1)) is None        # singular
    assert textbook_oracle._spd_det_inv(np.diag([1.0, -1.0, -1.0]).reshape(-1)) is None       # det > 0 but indefinite


def test_statistics_and_divergences_against_numpy():
    pts = lidar_cloud(6000, 3)
    r = textbook_oracle.run(pts, 200)
    base = ndt_oracle.run(pts, 200)
    x = pts.astype(np.float64)
    cells = np.unique(base.point_voxel[base.point_voxel >= 0])
    stats = {}
    for c in cells:
        m = x[base.point_voxel == c]
        stats[int(c)] = (m.mean(0), np.cov(m.T, bias=True) if len(m) > 1 else np.zeros((3, 3)))
    # the exported rows are population means / covariances of the voxel's points
    for v, mu, cov in zip(r.out_voxel, r.out_pts, r.out_cov):
        m, S = stats[int(v)]
        assert np.allclose(mu, m, rtol=1e-9, atol=1e-12) and np.allclose(cov.reshape(3, 3), S, rtol=1e-6, atol=1e-12)
    # closed-form KL on a sample of well-conditioned pairs
    checked = 0
    for a, b, d in zip(r.kl_p, r.kl_q, r.kl_div):
        (mp, Sp), (mq, Sq) = stats[int(a)], stats[int(b)]
        if np.linalg.cond(Sq) > 1e6 or np.linalg.cond(Sp) > 1e6:
            continue
        qi = np.linalg.inv(Sq)
        want = 0.5 * (np.trace(qi @ Sp) + (mq - mp) @ qi @ (mq - mp) - 3.0 + np.log(np.linalg.det(Sq) / np.linalg.det(Sp)))
        assert abs(d - want) <= 1e-6 * max(abs(want), 1.0), (a, b, d, want)
        assert d >= -1e-9                                     # a Kullback-Leibler divergence is non-negative
        checked += 1
    assert checked >= 20
    # selection: ascending list, the first occurrences of the first V - D distinct p's are gone, rows ascending
    assert np.all(np.diff(r.kl_div) >= 0)
    gone = []
    for a in r.kl_p:
        if int(a) not in gone:
            gone.append(int(a))
    gone = set(gone[:r.num_voxels - 200])
    assert set(r.out_voxel.tolist()) == set(int(c) for c in cells) - gone
    assert np.all(np.diff(r.out_voxel) > 0)
