"""GPU parity of the retained-handle continuation: NDT_Sampler.downsample(D) followed by
prune(D2), prune(D3) through the legacy ABI (ndt_legacy.py:173-240 -> core_legacy/src/ndt.c:28-117).

What is checked, bit for bit:
  * the surviving voxel set after every call equals "remove the next first-occurrence p's of the oracle's
    pre-prune list kl_p0", walked with the reference's shrinking-length stop (ndt.c:53) on the WELL-FORMED list
    (the entries behind the previous walk; the reference shifts by the walk length but keeps a larger count, so
    after a walk that skipped an entry it reads uninitialised memory);
  * rows come out in ascending voxel index with the oracle's means, (LU-mangled) covariances and labels;
  * num_valid_nds / num_kl_divergences / the -2 return follow the reference's counters;
  * whenever no walk so far skipped an entry (the case in which the reference's own continuation is well defined)
    the same call sequence on the reference build (oracle/_ref, deterministic schedule) gives the same rows.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import ndt_oracle, ref_ctypes
from tests import cases
from tests.helpers import same_bits
from tests.prune_model import ListWalk, RefHandles

pytestmark = pytest.mark.gpu


def _case(name):
    for c in cases.small_cases() + cases.medium_cases():
        if c[0] == name:
            return c
    raise KeyError(name)


CHAINS = [(n, d, ch, True) for n, d, ch in cases.CLEAN_PRUNE_CHAINS] + [(n, None, ch, False) for n, ch in [
    ("lidar16k_d1000", [900, 640]),
    ("lidar16k_d1000_labels", [990, 700]),
    ("modelnet2048_d512", [500, 300]),
    ("lidar8001_negative_axis", [250, 100]),
    ("modelnet3000_duplicates", [150, 20]),
    ("lidar120k_d1000_labels", [800, 256]),
    ("lattice112_isolated_prune_stop", [90, 50]),            # the first walk already stops (-2): every voxel is isolated
    ("modelnet4099_d100_tail", [60, 1]),                     # the continuation runs off the list -> -2
]]


@pytest.mark.parametrize("name,d_override,chain,clean", CHAINS, ids=[c[0] + ("_clean" if c[3] else "") for c in CHAINS])
def test_prune_continuation_matches_oracle_list_walk(name, d_override, chain, clean):
    from ndnet.preprocessing.ndt_legacy import NDT_Sampler, core
    _, pts, labels, ncls, d = _case(name)
    d = d_override or d
    o = ndt_oracle.run(pts, d, labels, ncls)
    assert o.ret == 0
    walk = ListWalk(o)
    assert walk.prune(d) == o.prune_ret and walk.n_valid == o.num_valid and walk.num_kl == o.num_kl
    s = NDT_Sampler(np.asarray(pts, np.float64), labels, ncls)
    p, c, k = s.downsample(d)
    assert s.status == 0 and s.num_valid_nds.value == walk.n_valid and s.num_kl_divergences.value == walk.num_kl
    ref = RefHandles(pts, labels, ncls, d) if ref_ctypes.have_ref("det") else None
    compared_with_reference = 0
    for d2 in chain:
        clean_before = not walk.skipped and walk.start < walk.K
        want_ret = walk.prune(d2)
        # the legacy symbols directly: return code and counters (NDT_Sampler.prune ignores the return code, ndt_legacy.py:194)
        got_ret = core.prune_nds(s.nd_array_ptr, s.len_x.value, s.len_y.value, s.len_z.value, d2, C.byref(s.num_valid_nds),
                                 s.kl_divergences_ptr, C.byref(s.num_kl_divergences))
        assert got_ret == want_ret, (name, d2)
        assert s.num_valid_nds.value == walk.n_valid and s.num_kl_divergences.value == walk.num_kl, (name, d2)
        rows = walk.rows()
        n = len(rows)
        gp = np.zeros((n + 4, 3)); gc = np.zeros((n + 4, 9)); gk = np.zeros(n + 4, np.uint16)
        n_out = C.c_ulong(0)
        dp, usp = C.POINTER(C.c_double), C.POINTER(C.c_ushort)
        core.to_point_cloud(s.nd_array_ptr, s.len_x.value, s.len_y.value, s.len_z.value, s.offset_x.value, s.offset_y.value,
                            s.offset_z.value, s.voxel_size.value, gp.ctypes.data_as(dp), C.byref(n_out), gc.ctypes.data_as(dp),
                            gk.ctypes.data_as(usp))
        assert n_out.value == n, (name, d2, n_out.value, n)
        assert same_bits(gp[:n], o.mean[rows]), (name, d2)          # exactly these voxels, ascending index
        assert same_bits(gc[:n], o.cov[rows]), (name, d2)
        if labels is not None:
            assert np.array_equal(gk[:n], o.cls[rows]), (name, d2)
        assert np.all(gp[n:] == 0) and np.all(gc[n:] == 0)
        if ref is not None and ref.ret == 0 and want_ret == 0 and clean_before:
            rp, rc, rk = ref.prune(d2)                               # the reference itself, same call sequence
            assert int(ref.n_valid.value) == n and same_bits(rp, gp[:n]) and same_bits(rc, gc[:n]), (name, d2)
            if labels is not None:
                assert np.array_equal(rk, gk[:n]), (name, d2)
            compared_with_reference += 1
        elif ref is not None:
            ref = None      # the reference's list is no longer well formed (or it stopped): do not call it again
    s.cleanup()
    if clean and ref_ctypes.have_ref("det"):
        assert compared_with_reference == len(chain)    # the whole chain was also checked against the reference itself
    if name == "modelnet4099_d100_tail":
        assert want_ret == -2 and walk.n_valid > 1
    if name == "lattice112_isolated_prune_stop":
        assert o.prune_ret == -2 and want_ret == -2 and walk.n_valid == o.num_valid0


def test_sampler_prune_method_returns_the_same_rows():
    """NDT_Sampler.prune (the Python method, ndt_legacy.py:173-240) on top of the symbols checked above."""
    from ndnet.preprocessing.ndt_legacy import NDT_Sampler
    _, pts, labels, ncls, d = _case("lidar16k_d1000_labels")
    o = ndt_oracle.run(pts, d, labels, ncls)
    walk = ListWalk(o)
    walk.prune(d)
    s = NDT_Sampler(np.asarray(pts, np.float64), labels, ncls)
    s.downsample(d)
    for d2 in (812, 333):
        assert walk.prune(d2) == 0
        p2, c2, k2 = s.prune(d2)
        rows = walk.rows()
        assert p2.shape == (d2, 3) and same_bits(p2, o.mean[rows]) and same_bits(c2, o.cov[rows])
        assert np.array_equal(k2.astype(np.uint16), o.cls[rows])
    s.cleanup()


def test_prune_refuses_more_than_valid_without_touching_state():
    """ndt.c:36-39: desired > valid -> -1; also with a NULL num_valid pointer (the guard must not depend on it)."""
    from ndnet.preprocessing.ndt_legacy import NDT_Sampler, core
    _, pts, labels, ncls, d = _case("modelnet2048_d512")
    s = NDT_Sampler(np.asarray(pts, np.float64), labels, ncls)
    s.downsample(d)
    before = s.num_valid_nds.value
    r = core.prune_nds(s.nd_array_ptr, s.len_x.value, s.len_y.value, s.len_z.value, before + 7, C.byref(s.num_valid_nds),
                       s.kl_divergences_ptr, C.byref(s.num_kl_divergences))
    assert r == -1 and s.num_valid_nds.value == before
    r = core.prune_nds(s.nd_array_ptr, s.len_x.value, s.len_y.value, s.len_z.value, before + 7, None,
                       s.kl_divergences_ptr, None)
    assert r == -1
    p2, _, _ = s.prune(before)            # nothing to remove: all rows still there
    assert p2.shape == (before, 3) and s.num_valid_nds.value == before
    s.cleanup()
