"""GPU parity tests of the NDT hot path: the CUDA library, called through its C ABI (device-pointer batched
entry, host-pointer batched entry and the legacy ndt_downsample symbols), against the CPU oracle and
against the golden vectors the reference's own C sources produced.

Bar (BASELINE.json north_star): voxel ids, per-voxel membership, the retained-distribution set, labels,
means and the exported (LU-mangled) covariances are BIT-EXACT (any NaN == any NaN).  The pseudo-KL values
may differ from the CPU in the last bits because CUDA's log() and glibc's log() round differently:
tolerance |d_gpu - d_cpu| <= KL_TOL_ULPS ulps of max(|trace|-scale terms), checked below; the selection must
still be identical unless two list entries at the cut are closer than that tolerance.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ndt_oracle
from tests import cases
from tests.helpers import same_bits

pytestmark = pytest.mark.gpu

KL_REL_TOL = 1e-13     # relative, on entries whose CPU and GPU values are finite and non-zero
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ndt_ref_golden.npz")


@pytest.fixture(scope="module")
def engine():
    from ndnet_b200.engine import NdtEngine
    e = NdtEngine(0)
    e.keep_point_voxels(True)      # the parity tests compare every point's voxel id
    e.keep_kl_list(True)           # ... and the whole sorted divergence list
    return e


def _run_gpu(engine, pts, labels, ncls, d):
    t = torch.from_numpy(np.ascontiguousarray(pts)).cuda()[None]
    lab = None if labels is None else torch.from_numpy(labels.astype(np.int16)).cuda()[None]
    out = engine.downsample(t, d, lab, ncls, nan_to_num=False, want_f64=True, want_voxel=True)
    torch.cuda.synchronize()
    return out


def _check_against_oracle(engine, name, pts, labels, ncls, d):
    out = _run_gpu(engine, pts, labels, ncls, d)
    o = ndt_oracle.run(pts, d, labels, ncls)
    info = out.info[0]
    assert info["status"] == o.ret, name
    assert tuple(int(x) for x in info["len"]) == o.lens, name
    assert info["voxel_size"] == o.voxel_size, name
    assert info["evaluations"] == o.evaluations, name
    assert np.array_equal(info["limits"], o.limits), name
    n = pts.shape[0]
    pv = engine.last_point_voxels(1, n).cpu().numpy()[0]
    if o.ret != 0:
        assert np.all(out.feat64.cpu().numpy() == 0) and info["num_out"] == 0, name
        return
    # voxel ids and membership: bit-exact
    assert np.array_equal(pv, o.point_voxel), name
    assert info["num_voxels"] == o.num_valid0 and info["num_valid"] == o.num_valid, name
    assert info["prune_status"] == o.prune_ret, name
    # divergence list: same (p, q) multiset in the same order unless values within tolerance swapped
    div, p, q = engine.last_kl_list(0, int(info["num_kl"]) + 8)
    assert len(div) == o.num_kl0, name
    gpu_map = {(int(a), int(b)): v for a, b, v in zip(p, q, div)}
    assert len(gpu_map) == len(div), name
    for a, b, v in zip(o.kl_p0, o.kl_q0, o.kl_div0):
        w = gpu_map[(int(a), int(b))]
        if np.isnan(v) or np.isnan(w):
            assert np.isnan(v) and np.isnan(w), name
        elif v != w:
            assert np.isfinite(v) and np.isfinite(w) and abs(v - w) <= KL_REL_TOL * max(abs(v), 1.0), (name, v, w)
    # retained-distribution set, rows in ascending voxel index
    rows = min(o.num_out, d)
    assert info["num_out"] == rows and info["num_survivors"] == o.num_out, name
    vox = out.voxel[0].cpu().numpy()
    if not np.array_equal(vox[:rows], o.out_voxel[:rows]):
        # only acceptable when the divergences at the cut are tied within tolerance
        cpu_removed = set(np.flatnonzero((o.num_samples0 > 0) & (o.num_samples == 0)).tolist())
        gpu_removed = set(np.flatnonzero(o.num_samples0 > 0).tolist()) - set(vox[:rows].tolist())
        diff = cpu_removed ^ gpu_removed
        best = {}
        for a, v in zip(o.kl_p0, o.kl_div0):
            best.setdefault(int(a), v)
        vals = [best[x] for x in diff if x in best]
        assert vals and (max(vals) - min(vals)) <= KL_REL_TOL * max(abs(max(vals)), 1.0), (name, sorted(diff))
        return
    assert np.all(vox[rows:] == -1), name
    f = out.feat64[0].cpu().numpy()
    assert same_bits(f[:rows, :3], o.out_pts[:rows]), name
    assert same_bits(f[:rows, 3:], o.out_cov[:rows]), name
    assert np.all(f[rows:] == 0), name
    if labels is not None:
        assert np.array_equal(out.labels[0].cpu().numpy().astype(np.uint16)[:rows], o.out_cls[:rows]), name


@pytest.mark.parametrize("case", cases.small_cases(), ids=lambda c: c[0])
def test_small_cases_match_oracle(engine, case):
    _check_against_oracle(engine, *case)


def test_random_clouds_match_oracle(engine):
    """Differential run on 40 seeded random clouds (sizes 9..9000, D 2..400, with and without labels, exact duplicates,
    lattice points whose cells sit on exact boundaries, anisotropic scales) - the same generator the oracle is checked
    with against the reference build (tests/test_oracle.py::test_oracle_matches_reference_live)."""
    rng = np.random.default_rng(5)
    for k in range(40):
        n = int(rng.integers(500, 9000)) if k % 2 == 0 else int(rng.integers(9, 400))
        d = int(rng.integers(20, 400)) if k % 2 == 0 else int(rng.integers(2, max(3, n // 3)))
        pts = (rng.normal(size=(n, 3)) * rng.uniform(0.5, 20, 3)).astype(np.float32)
        if k % 5 == 4:
            pts[n // 2:] = pts[: n - n // 2]
        if k % 7 == 6:
            pts = np.round(pts * 4) / 4
        ncls = 6 if k % 3 == 1 else (70 if k % 3 == 2 else 0)          # 70 classes: the global-histogram vote path
        labels = rng.integers(0, ncls + 1, n).astype(np.uint16) if ncls else None
        _check_against_oracle(engine, f"random{k}", pts, labels, ncls, d)


@pytest.mark.parametrize("case", cases.medium_cases(), ids=lambda c: c[0])
def test_baseline_sized_clouds_match_oracle(engine, case):
    _check_against_oracle(engine, *case)


def test_wide_label_set_uses_the_global_vote_path(engine):
    """More than 64 classes: the label vote goes through global atomics instead of the shared-memory histogram."""
    from ndnet_b200.synth import lidar_cloud
    pts = lidar_cloud(30000, 61)
    labels = np.random.default_rng(61).integers(0, 201, 30000).astype(np.uint16)
    _check_against_oracle(engine, "wide_labels", pts, labels, 200, 700)


def test_large_d_uses_the_global_memory_sort(engine):
    """n_desired_nds = 8160 (tools/train_multiscale.py:33-36): ~30k list entries, beyond the shared-memory sort."""
    from ndnet_b200.synth import lidar_cloud
    pts = lidar_cloud(120000, 62)
    _check_against_oracle(engine, "d8160", pts, None, 0, 8160)


def test_fp64_batch_equals_fp32_batch_on_fp32_values(engine):
    """The f64 entry (legacy ABI dtype) and the f32 entry agree on clouds whose coordinates are fp32 values."""
    from ndnet_b200.synth import lidar_batch
    pts = lidar_batch(3, 20000, seed0=63)
    a = engine.downsample(torch.from_numpy(pts).cuda(), 400, nan_to_num=False, want_f64=True, want_voxel=True)
    b = engine.downsample(torch.from_numpy(pts.astype(np.float64)).cuda(), 400, nan_to_num=False, want_f64=True, want_voxel=True)
    assert torch.equal(a.voxel, b.voxel) and same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy())


def test_multiscale_resolutions_match_oracle(engine):
    """BASELINE config 5: n_desired_nds 4096 / 1024 / 256 for the same scan."""
    from ndnet_b200.synth import lidar_cloud
    pts = lidar_cloud(120000, 71)
    outs = engine.downsample_multiscale(torch.from_numpy(pts).cuda()[None], [4096, 1024, 256], nan_to_num=False,
                                        want_f64=True, want_voxel=True)
    for d, out in zip([4096, 1024, 256], outs):
        o = ndt_oracle.run(pts, d)
        assert out.info[0]["status"] == o.ret == 0 and out.info[0]["num_out"] == d
        assert np.array_equal(out.voxel[0].cpu().numpy(), o.out_voxel)
        assert same_bits(out.feat64[0, :, :3].cpu().numpy(), o.out_pts) and same_bits(out.feat64[0, :, 3:].cpu().numpy(), o.out_cov)


def test_golden_vectors_through_c_abi(engine):
    """The reference's own outputs (tests/golden, made from core_legacy/src by make_golden.py)."""
    g = np.load(GOLDEN)
    by_name = {c[0]: c for c in cases.small_cases() + cases.medium_cases()[:3]}
    for name in [str(n) for n in g["names"]]:
        _, pts, labels, ncls, d = by_name[name]
        out = _run_gpu(engine, pts, labels, ncls, d)
        info = out.info[0]
        hdr = g[name + "/hdr"]
        assert info["status"] == hdr[0], name
        assert tuple(int(x) for x in info["len"]) == tuple(int(x) for x in hdr[1:4]), name
        assert info["voxel_size"] == g[name + "/vs"][0], name
        if hdr[0] != 0:
            continue
        rows = min(int(hdr[4]), d)
        f = out.feat64[0].cpu().numpy()
        assert same_bits(f[:rows, :3], g[name + "/pts"][:rows]), name
        assert same_bits(f[:rows, 3:], g[name + "/cov"][:rows]), name
        if labels is not None:
            assert np.array_equal(out.labels[0].cpu().numpy().astype(np.uint16)[:rows], g[name + "/cls"][:rows]), name
        assert info["num_valid"] == hdr[5] and info["num_kl_after"] == hdr[6], name


def test_batch_of_mixed_clouds_equals_single_cloud_runs(engine):
    """Clouds in a batch are independent: a batch gives exactly what B single-cloud calls give, including
    a cloud that fails (-3) in the middle of the batch."""
    from ndnet_b200.synth import lidar_cloud
    n, d = 20000, 500
    pts = np.stack([lidar_cloud(n, 40 + b) for b in range(5)])
    pts[2, :, 2] = 1.0      # degenerate axis -> -3
    out = engine.downsample(torch.from_numpy(pts).cuda(), d, nan_to_num=False, want_f64=True, want_voxel=True)
    for b in range(5):
        single = engine.downsample(torch.from_numpy(pts[b:b + 1]).cuda(), d, nan_to_num=False, want_f64=True, want_voxel=True)
        assert out.info[b]["status"] == single.info[0]["status"] == (-3 if b == 2 else 0)
        assert same_bits(out.feat64[b].cpu().numpy(), single.feat64[0].cpu().numpy())
        assert torch.equal(out.voxel[b], single.voxel[0])


def test_f32_features_and_nan_to_num(engine):
    """ndtnet_preprocessing.py:60-69: float32 cast then NaN/inf -> 0."""
    name, pts, labels, ncls, d = cases.small_cases()[0]
    t = torch.from_numpy(pts).cuda()[None]
    raw = engine.downsample(t, d, nan_to_num=False, want_f64=True)
    clean = engine.downsample(t, d, nan_to_num=True)
    ref = torch.nan_to_num(raw.feat64.to(torch.float32), nan=0.0, posinf=0.0, neginf=0.0)
    assert torch.equal(clean.feat, ref)
    assert torch.isfinite(clean.feat).all()


def test_full_size_batch_properties(engine):
    """BASELINE.json config 4 shape (120k-point scans, D=1000), size-independent properties: every cloud
    converges, D rows each, rows sorted by voxel index, idempotent across repeated calls, one cloud of the
    batch equals the oracle."""
    from ndnet_b200.synth import lidar_batch
    B, n, d = 16, 120000, 1000
    pts = lidar_batch(B, n, seed0=200)
    t = torch.from_numpy(pts).cuda()
    a = engine.downsample(t, d, nan_to_num=False, want_f64=True, want_voxel=True)
    b = engine.downsample(t, d, nan_to_num=False, want_f64=True, want_voxel=True)
    assert same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy()) and torch.equal(a.voxel, b.voxel)
    assert np.all(a.info["status"] == 0) and np.all(a.info["num_out"] == d) and np.all(a.info["num_valid"] == d)
    vox = a.voxel.cpu().numpy()
    assert np.all(np.diff(vox, axis=1) > 0)
    o = ndt_oracle.run(pts[5], d)
    assert np.array_equal(vox[5], o.out_voxel) and same_bits(a.feat64[5, :, :3].cpu().numpy(), o.out_pts)
    assert same_bits(a.feat64[5, :, 3:].cpu().numpy(), o.out_cov)


def test_host_entry_matches_device_entry(engine):
    from ndnet_b200.synth import lidar_batch
    pts, lab = lidar_batch(3, 30000, seed0=300, with_labels=True)
    feat_h, lab_h, info_h = engine.downsample_host(pts, 700, lab, 28)
    dev = engine.downsample(torch.from_numpy(pts).cuda(), 700, torch.from_numpy(lab.astype(np.int16)).cuda(), 28)
    assert torch.equal(feat_h, dev.feat.cpu())
    assert np.array_equal(lab_h, dev.labels.cpu().numpy().astype(np.uint16))
    assert np.array_equal(info_h["num_valid"], dev.info["num_valid"])


def test_ndt_preprocessing_drop_in(engine):
    """Same call as ndnet/preprocessing/ndtnet_preprocessing.py:6 (one-hot classes in, one-hot classes out)."""
    from ndnet.preprocessing.ndtnet_preprocessing import ndt_preprocessing
    from ndnet_b200.synth import lidar_batch
    B, n, d, C = 2, 16000, 500, 28
    pts, lab = lidar_batch(B, n, seed0=400, with_labels=True)
    onehot = torch.zeros((B, n, C + 1)); onehot.scatter_(2, torch.from_numpy(lab.astype(np.int64))[..., None], 1.0)
    p, c, k = ndt_preprocessing(d, torch.from_numpy(pts).cuda(), onehot.cuda(), C)
    assert p.shape == (B, d, 3) and c.shape == (B, d, 9) and k.shape == (B, d, C + 1)
    assert p.dtype == c.dtype == k.dtype == torch.float32 and p.is_cuda
    for b in range(B):
        o = ndt_oracle.run(pts[b], d, lab[b], C)
        assert torch.equal(p[b].cpu(), torch.from_numpy(o.out_pts).float())
        ref_c = torch.nan_to_num(torch.from_numpy(o.out_cov).float(), nan=0.0, posinf=0.0, neginf=0.0)
        assert torch.equal(c[b].cpu(), ref_c)
        assert torch.equal(k[b].argmax(1).cpu(), torch.from_numpy(o.out_cls.astype(np.int64)))
        assert torch.all(k[b].sum(1) == 1)


def test_count_division_is_ieee_exact():
    """k_stats divides by the running count with a reciprocal + one FMA correction; it must equal the IEEE
    quotient bit for bit (2e8 pseudo-random operand pairs, counts up to 6.7e7, exponents up to +-1000)."""
    from ndnet_b200 import _lib
    assert _lib.lib().ndnet_b200_selftest_div(200_000_000, 12345) == 0


def test_legacy_sampler_drop_in():
    """NDT_Sampler (ndt_legacy.py:45-240) bound to the legacy symbols of libndnet_b200.so."""
    from ndnet.preprocessing.ndt_legacy import NDT_Sampler
    name, pts, labels, ncls, d = cases.small_cases()[1]
    s = NDT_Sampler(pts.astype(np.float64), labels, ncls)
    p, c, k = s.downsample(d)
    o = ndt_oracle.run(pts, d, labels, ncls)
    assert s.status == 0 and same_bits(p, o.out_pts) and same_bits(c, o.out_cov) and np.array_equal(k, o.out_cls)
    assert (s.len_x.value, s.len_y.value, s.len_z.value) == o.lens and s.voxel_size.value == o.voxel_size
    assert s.num_valid_nds.value == o.num_valid and s.num_kl_divergences.value == o.num_kl
    # continuation on the retained handles: same as asking for the smaller number on the same list
    p2, c2, k2 = s.prune(d - 100)
    assert p2.shape == (d - 100, 3) and s.num_valid_nds.value == d - 100
    kept = {tuple(r) for r in p.tolist()}
    assert all(tuple(r) in kept for r in p2.tolist())
    s.cleanup()
    # failure path never crashes (A16)
    bad = NDT_Sampler(cases.small_cases()[11][1].astype(np.float64))
    bad.downsample(64)
    assert bad.status == -3
    bad.cleanup()


def test_nan_coordinates_are_refused_not_voxelised(engine):
    """A NaN coordinate makes the reference convert floor(NaN) to unsigned (undefined behaviour, voxel.c:89-91).  The
    library refuses such a cloud (status -5, zero rows) and leaves the other clouds of the batch untouched; +-inf
    coordinates cannot be held by any grid and end in -1."""
    from ndnet_b200.synth import lidar_batch
    pts = lidar_batch(4, 20000, seed0=910)
    clean = engine.downsample(torch.from_numpy(pts).cuda(), 400, nan_to_num=False, want_f64=True, want_voxel=True)
    bad = pts.copy()
    bad[1, 12345, 2] = np.nan
    bad[3, 7, 0] = np.inf
    out = engine.downsample(torch.from_numpy(bad).cuda(), 400, nan_to_num=False, want_f64=True, want_voxel=True)
    assert out.info["status"].tolist() == [0, -5, 0, -1]
    for b in (1, 3):
        assert out.info[b]["num_out"] == 0 and torch.all(out.feat64[b] == 0) and torch.all(out.voxel[b] == -1)
    for b in (0, 2):
        assert same_bits(out.feat64[b].cpu().numpy(), clean.feat64[b].cpu().numpy()) and torch.equal(out.voxel[b], clean.voxel[b])
    # the legacy symbol reports it too
    from ndnet.preprocessing.ndt_legacy import NDT_Sampler
    s = NDT_Sampler(bad[1].astype(np.float64))
    p, c, k = s.downsample(400)
    assert s.status == -5 and np.all(p == 0) and np.all(c == 0)
    s.cleanup()


def test_failed_workspace_allocation_leaves_a_usable_context():
    """A workspace allocation that fails must not leave capacities behind (the next call would run every kernel on
    null buffers): the failing call returns the error, the next one allocates afresh and gives the usual result."""
    from ndnet_b200 import _lib
    from ndnet_b200.engine import NdtEngine
    from ndnet_b200.synth import lidar_batch
    e = NdtEngine(0)
    pts = torch.from_numpy(lidar_batch(2, 16000, seed0=920)).cuda()
    first = e.downsample(pts, 300, nan_to_num=False, want_f64=True)
    assert _lib.lib().ndnet_b200_test_fail_next_reserve(e.handle) == 0
    bigger = torch.from_numpy(lidar_batch(3, 16000, seed0=920)).cuda()          # needs a larger workspace -> reserve() runs
    with pytest.raises(RuntimeError, match="workspace allocation"):
        e.downsample(bigger, 300, nan_to_num=False, want_f64=True)
    again = e.downsample(pts, 300, nan_to_num=False, want_f64=True)              # same shape as before the failure
    torch.cuda.synchronize()
    assert np.all(again.info["status"] == 0) and same_bits(again.feat64.cpu().numpy(), first.feat64.cpu().numpy())
    third = e.downsample(bigger, 300, nan_to_num=False, want_f64=True)
    assert same_bits(third.feat64[:2].cpu().numpy(), first.feat64.cpu().numpy())
    e.close()


def test_partial_sort_selects_what_the_full_sort_selects(engine):
    """The batched calls sort only the head of the divergence list (the part the prune walk can reach, found by a radix
    selection); with ndnet_b200_keep_kl_list the whole list is sorted.  Same retained set, rows, labels and bookkeeping on
    LiDAR-like and object clouds, several n_desired_nds (few and many removals), a labelled batch."""
    from ndnet_b200.engine import NdtEngine
    from ndnet_b200.synth import lidar_batch, modelnet_cloud
    part = NdtEngine(0)                       # keep_kl_list off: the production configuration
    pts, lab = lidar_batch(6, 40000, seed0=300, with_labels=True)
    objs = np.stack([modelnet_cloud(2048, s) for s in range(6)])
    from ndnet_b200.synth import lidar_cloud
    big = lidar_cloud(120000, 62)[None]       # D = 4096: the launch grows its shared memory with D; D = 16000: 17 k reachable
    for cloud, labels, ncls, ds in ((pts, lab, 28, (1000, 400, 64)), (objs, None, 0, (512, 128, 16)), (big, None, 0, (4096, 16000))):   # entries, beyond the largest launch: global-memory sort
        t = torch.from_numpy(cloud).cuda()
        l = None if labels is None else torch.from_numpy(labels.astype(np.int16)).cuda()
        for d in ds:
            a = engine.downsample(t, d, l, ncls, nan_to_num=False, want_f64=True, want_voxel=True)
            b = part.downsample(t, d, l, ncls, nan_to_num=False, want_f64=True, want_voxel=True)
            assert torch.equal(a.voxel, b.voxel), d
            assert same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy()), d
            if l is not None:
                assert torch.equal(a.labels, b.labels), d
            for k in ("status", "prune_status", "num_voxels", "num_valid", "num_kl", "num_kl_after", "num_out", "num_survivors"):
                assert np.array_equal(a.info[k], b.info[k]), (d, k)
    with pytest.raises(RuntimeError):
        part.last_kl_list(0, 16)              # the whole list was not kept


def test_graph_replay_equals_direct_launches():
    """ndnet_b200_set_ndt_graph: the captured chain, replayed on the context's own buffers, returns bit for bit what the direct
    launches return - for new scans of the same shape (the graph re-reads its inputs), with labels, and the launch counter
    advances by the same amount; a new shape goes back to direct launches until it is seen again."""
    from ndnet_b200 import _lib
    from ndnet_b200.engine import NdtEngine
    from ndnet_b200.synth import lidar_batch
    L = _lib.lib()
    direct, graphed = NdtEngine(0), NdtEngine(0)
    direct.set_graph(0)
    graphed.set_graph(1)
    B, N, D, C = 3, 20000, 150, 12

    def run(e, seed, n=N):
        p, l = lidar_batch(B, n, seed0=seed, with_labels=True, num_classes=C)
        pt, lt = torch.from_numpy(p).cuda(), torch.from_numpy(l.astype(np.int16)).cuda()
        n0 = L.ndnet_b200_launch_count()
        o = e.downsample(pt, D, lt, C, nan_to_num=False, want_f64=True, want_voxel=True)
        torch.cuda.synchronize()
        return o, L.ndnet_b200_launch_count() - n0

    for call, seed in enumerate([11, 11, 12, 13, 11]):      # calls 0-1: direct + capture; 2-4: replays on other scans
        a, na = run(direct, seed)
        b, nb = run(graphed, seed)
        assert na == nb and na >= 20, (call, na, nb)
        assert same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy()), call
        assert same_bits(a.feat.cpu().numpy(), b.feat.cpu().numpy()), call          # nan_to_num is off: NaNs compare by bits
        assert torch.equal(a.voxel, b.voxel) and torch.equal(a.labels, b.labels), call
        assert a.info.tobytes() == b.info.tobytes(), call
        assert np.any(a.info["num_out"] > 0), call
    # another shape in between, then the first shape again: still the same answers
    a, _ = run(direct, 21, n=N + 8)
    b, _ = run(graphed, 21, n=N + 8)
    assert same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy())
    a, _ = run(direct, 14)
    b, _ = run(graphed, 14)
    assert same_bits(a.feat64.cpu().numpy(), b.feat64.cpu().numpy()) and torch.equal(a.voxel, b.voxel)
    direct.close(); graphed.close()
