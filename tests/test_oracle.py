"""CPU tests of the oracle (test infrastructure) itself: the reference's own known answers, the golden
vectors produced by the reference's C sources (tests/golden/make_golden.py), and - when oracle/_ref is
present - the reference build run live."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ndt_oracle, ref_ctypes
from tests import cases
from tests.helpers import same_bits

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ndt_ref_golden.npz")


def test_limits_known_answer():
    # core_legacy/tests/test_pointclouds.cpp:5-23
    pc = np.array([[0, 1, 0], [1, 0, 0], [0, -1, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -2]], np.float64)
    lim = ndt_oracle.limits(pc)
    assert lim.tolist() == [1.0, 1.0, 1.0, -1.0, -1.0, -2.0]


def test_limits_dbl_min_quirk():
    # pointclouds.c:44-46: max starts at DBL_MIN, so an all-negative axis reports 2.2e-308
    pc = np.array([[-1, -2, 3], [-4, -5, 6]], np.float64)
    lim = ndt_oracle.limits(pc)
    assert lim[0] == np.finfo(np.float64).tiny and lim[1] == np.finfo(np.float64).tiny and lim[2] == 6.0


@pytest.mark.parametrize("lens,direction,expected", [((5, 3, 2), 4, 22), ((5, 3, 1), 2, 12), ((5, 3, 1), 0, 8)])
def test_neighbor_known_answers(lens, direction, expected):
    # core_legacy/tests/test_voxel.cpp:152-180 (index 7; Z+ -> 22, Y+ -> 12, X+ -> 8)
    r, idx = ndt_oracle.neighbor(7, *lens, direction)
    assert r == 0 and idx == expected


def test_neighbor_out_of_grid():
    assert ndt_oracle.neighbor(0, 5, 3, 2, 1)[0] == -4      # x-1 at x == 0 wraps (voxel.c:157-166)
    assert ndt_oracle.neighbor(4, 5, 3, 2, 0)[0] == -4
    assert ndt_oracle.neighbor(29, 5, 3, 2, 4)[0] == -4


def test_cube16_count():
    # core_legacy/tests/test_ndt.cpp:28-31 expects 8 distributions from the 16-point double cube
    r = ndt_oracle.run(cases.CUBE16, 8)
    assert r.ret == 0 and r.num_out == 8


def test_smoke_shape_90000_to_24():
    # core_legacy/tests/ndt_downsample.c:19-60: 90 000 uniform [0,1]^3 points -> 24 NDs must succeed
    rng = np.random.default_rng(0)
    r = ndt_oracle.run(rng.random((90000, 3)), 24)
    assert r.ret == 0 and r.num_out == 24


def _all_cases():
    return cases.small_cases() + cases.medium_cases()[:3]


def test_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    names = [str(n) for n in g["names"]]
    by_name = {c[0]: c for c in _all_cases()}
    assert set(names) == set(by_name)
    for name in names:
        _, pts, labels, ncls, d = by_name[name]
        sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest(), np.uint8)
        assert np.array_equal(sha, g[name + "/sha"]), f"{name}: seeded input differs from the one the fixture was made with"
        o = ndt_oracle.run(pts, d, labels, ncls)
        hdr = g[name + "/hdr"]
        assert o.ret == hdr[0], name
        assert o.lens == tuple(int(x) for x in hdr[1:4]), name
        assert o.voxel_size == g[name + "/vs"][0], name
        if o.ret != 0:
            continue
        assert np.array_equal(o.offsets, g[name + "/vs"][1:]), name
        assert (o.num_out, o.num_valid) == (hdr[4], hdr[5]), name
        assert o.num_kl == hdr[6], name
        assert same_bits(o.out_pts, g[name + "/pts"]), name
        assert same_bits(o.out_cov, g[name + "/cov"]), name
        if labels is not None:
            assert np.array_equal(o.out_cls, g[name + "/cls"]), name


@pytest.mark.skipif(not ref_ctypes.have_ref("det"), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_live():
    lib = ref_ctypes.load(ref_ctypes.ref_lib_path("det"))
    rng = np.random.default_rng(5)
    for k in range(24):
        n = int(rng.integers(500, 9000)) if k < 12 else int(rng.integers(9, 400))
        d = int(rng.integers(20, 400)) if k < 12 else int(rng.integers(2, max(3, n // 3)))
        pts = (rng.normal(size=(n, 3)) * rng.uniform(0.5, 20, 3)).astype(np.float32)
        if k % 5 == 4:
            pts[n // 2:] = pts[: n - n // 2]                          # exact duplicates
        if k % 7 == 6:
            pts = np.round(pts * 4) / 4                               # lattice points: cells on exact boundaries
        labels = rng.integers(0, 7, n).astype(np.uint16) if k % 2 else None
        r = ref_ctypes.downsample(lib, pts.astype(np.float64), d, labels, 6 if labels is not None else 0, introspect=True)
        o = ndt_oracle.run(pts, d, labels, 6 if labels is not None else 0)
        assert r.ret == o.ret and r.lens == o.lens and r.voxel_size == o.voxel_size
        if r.ret != 0:
            continue
        assert r.num_out == o.num_out and r.num_valid == o.num_valid and r.num_kl == o.num_kl
        assert same_bits(r.points[: r.num_out], o.out_pts) and same_bits(r.covs[: r.num_out], o.out_cov)
        if labels is not None:
            assert np.array_equal(r.classes[: r.num_out], o.out_cls)
        # internals: the reference's nd_array after pruning
        occupied = r.nd["num_samples"] > 0
        assert np.array_equal(occupied, o.num_samples > 0)
        assert same_bits(r.nd["mean"][occupied], o.mean[occupied])


@pytest.mark.skipif(not ref_ctypes.have_ref("det"), reason="reference build absent")
@pytest.mark.parametrize("name,d,chain", cases.CLEAN_PRUNE_CHAINS, ids=[c[0] for c in cases.CLEAN_PRUNE_CHAINS])
def test_prune_continuation_model_matches_reference_live(name, d, chain):
    """Row f1: the list-walk restatement the GPU test uses (tests/prune_model.py::ListWalk on the oracle's sorted list)
    against the reference build's own prune_nds / to_point_cloud on retained handles.  The reference's continuation is
    only defined while no walk skipped an already-removed entry (its shifted list then reads past its end, SURVEY.md
    A15); cases.CLEAN_PRUNE_CHAINS are (cloud, D, further targets) picked so that the whole chain stays in that regime."""
    from tests.prune_model import ListWalk, RefHandles
    _, pts, labels, ncls, _ = [c for c in cases.small_cases() if c[0] == name][0]
    o = ndt_oracle.run(pts, d, labels, ncls)
    walk = ListWalk(o)
    assert walk.prune(d) == o.prune_ret == 0 and not walk.skipped
    ref = RefHandles(pts, labels, ncls, d)
    assert ref.ret == 0 and ref.n_valid.value == walk.n_valid and ref.n_kl.value == walk.num_kl
    for d2 in chain:
        assert walk.prune(d2) == 0 and not walk.skipped
        rp, rc, rk = ref.prune(d2)
        rows = walk.rows()
        assert ref.n_valid.value == walk.n_valid == len(rows) == d2 and ref.n_kl.value == walk.num_kl
        assert same_bits(rp, o.mean[rows]) and same_bits(rc, o.cov[rows])
        if labels is not None:
            assert np.array_equal(rk, o.cls[rows])
    ref.close()
