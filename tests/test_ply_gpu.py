"""GPU parity of the ASCII-PLY ingest (ndnet_b200_ply_*) against the oracle restatement of the
reference reader (oracle/ply_oracle.py <- /root/reference/ndnet/datasets/CARLA_Seg.py:96-183) and against the fixture
the reference itself produced.  Bit-exact: points are float32(correctly rounded double), tags and one-hot exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ply_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HEAD = b"".join(b"header line %d\n" % i for i in range(10))


def body(n, seed, n_classes=28, eol="\n", trailing=True):
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        x, y, z = (float(v) for v in rng.normal(0, 40, 3))
        tag = int(rng.integers(0, n_classes + 1))
        k = i % 6
        if k == 0:
            xyz = f"{x:.4f} {y:.4f} {z:.4f}"
        elif k == 1:
            xyz = f"{x!r} {y!r} {z!r}"
        elif k == 2:
            xyz = f"{x:.7e} {y:.2E} {z:+.6f}"
        elif k == 3:
            xyz = f"{np.float32(x).item()!r}\t{np.float32(y).item()!r}  {np.float32(z).item()!r}"
        elif k == 4:
            xyz = f" {int(x)} {y:.1f} .{abs(int(z * 1000)) % 1000:03d}"
        else:
            xyz = f"{x * 1e-9:.12e} {y * 1e7!r} {z:.15f}"
        rows.append(f"{xyz} {rng.uniform(-1, 1):.4f} {int(rng.integers(0, 999))} {tag}")
    text = eol.join(rows) + (eol if trailing else "")
    return text.encode()


def same(cloud, raw, n_classes, indexes, header=10):
    from ndnet_b200.ply import PlyCloud
    n_rows = len(ply_oracle.read_lines(raw)) - header
    want_p, want_gt, want_t = ply_oracle.get_data_pcl(raw, n_classes, indexes if indexes is not None else np.arange(n_rows), header)
    c = PlyCloud(raw, n_classes, header) if cloud is None else cloud
    assert c.num_points == len(ply_oracle.read_lines(raw)) - header
    p, gt, t = c.sample(indexes)
    assert np.array_equal(p.cpu().numpy().view(np.uint32), want_p.view(np.uint32))
    assert np.array_equal(gt.cpu().numpy(), want_gt)
    assert np.array_equal(t.cpu().numpy().astype(np.int64), want_t.astype(np.int64))
    return c


def test_reference_fixture_through_c_abi():
    from ndnet_b200.ply import PlyCloud
    g = np.load(os.path.join(GOLD, "ply_ref_golden.npz"))
    with open(os.path.join(GOLD, "carla_like.ply"), "rb") as f:
        raw = f.read()
    c = PlyCloud(raw, int(g["n_classes"]))
    assert c.num_points == 700
    p, gt, _ = c.sample(g["indexes"])
    assert np.array_equal(p.cpu().numpy().view(np.uint32), g["points"].view(np.uint32))
    assert np.array_equal(gt.cpu().numpy(), g["gt"])


@pytest.mark.parametrize("n,eol,trailing", [(1, "\n", True), (37, "\n", False), (5000, "\r\n", True), (4097, "\r\n", False),
                                            (20000, "\n", True)])
def test_generated_files_match_oracle(n, eol, trailing):
    raw = HEAD + body(n, n, eol=eol, trailing=trailing)
    rng = np.random.default_rng(n)
    c = same(None, raw, 28, None)
    same(c, raw, 28, rng.choice(n, max(1, n // 3), replace=False))


def test_full_size_scan_file():
    """120k data lines (BASELINE config 4's scan size): every point against the oracle."""
    raw = HEAD + body(120_000, 7)
    c = same(None, raw, 28, None)
    idx = np.random.default_rng(0).choice(120_000, 16_000, replace=False)
    same(c, raw, 28, idx)
    # size-independent property: sampling everything in file order equals the identity gather
    p_all, _, t_all = c.sample(None)
    p_idx, _, t_idx = c.sample(torch.from_numpy(idx).cuda())
    assert torch.equal(p_all[idx], p_idx) and torch.equal(t_all.to(torch.int32)[idx], t_idx.to(torch.int32))


def test_header_line_count_and_device_resident_text():
    from ndnet_b200.ply import PlyCloud
    raw = b"a\nb\nc\n" + body(300, 3)
    dev_text = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    c = PlyCloud(dev_text, 28, num_header_lines=3)
    same(c, raw, 28, None, header=3)
    empty = PlyCloud(HEAD, 28)
    assert empty.num_points == 0
    assert PlyCloud(b"", 28).num_points == 0


@pytest.mark.parametrize("line,exc,match,oracle_raises", [
    (b"1.0 2.0\n", IndexError, None, True),
    (b"\n", IndexError, None, True),
    (b"1.0 abc 3.0 4\n", ValueError, None, True),
    (b"1.0 2.0 3.0 4.5\n", ValueError, None, True),
    (b"1.0 2.0 3.0 29\n", ValueError, "Class tag 29 out of bounds", True),
    (b"1.0 2.0 3.0 -2\n", OverflowError, None, True),
    (b"1.0 nan 3.0 1\n", ValueError, "converts exactly", False),          # CPython accepts it; we refuse loudly
    (b"1.0 2.0 3.0 1\rnext", ValueError, "carriage-return", False),       # old-Mac line ends: refused loudly
])
def test_errors_follow_the_reference(line, exc, match, oracle_raises):
    from ndnet_b200.ply import PlyCloud
    raw = HEAD + body(50, 1) + line + body(50, 2)
    with pytest.raises(exc, match=match):
        PlyCloud(raw, 28)
    if oracle_raises:
        with pytest.raises(exc, match=match):
            ply_oracle.parse(raw, 28)


def test_first_offending_line_wins():
    from ndnet_b200.ply import PlyCloud
    raw = HEAD + body(10, 1) + b"1 2 3 -1\n" + body(500, 2) + b"1 2 3 40\n" + body(5, 3) + b"1 2\n"
    with pytest.raises(ValueError, match="Class tag 40 out of bounds"):       # the loop's error precedes the uint16 cast
        PlyCloud(raw, 28)
    with pytest.raises(ValueError, match="Class tag 40 out of bounds"):
        ply_oracle.parse(raw, 28)


def test_dataset_drop_in(tmp_path):
    """ndnet.datasets.CARLA_Seg: same constructor and outputs as the reference class on a seeded draw."""
    from ndnet.datasets.CARLA_Seg import CARLA_Seg
    raws = {}
    for k in range(3):
        raws[f"scan_{k}.ply"] = HEAD + body(3000 + k, 10 + k)
        (tmp_path / f"scan_{k}.ply").write_bytes(raws[f"scan_{k}.ply"])
    ds = CARLA_Seg(28, 1000, str(tmp_path))
    assert len(ds) == 3
    for k in range(3):
        np.random.seed(k)
        pts, gt = ds[k]
        np.random.seed(k)
        idx = np.random.choice(3000 + k, 1000, replace=False)
        want_p, want_gt, _ = ply_oracle.get_data_pcl(raws[f"scan_{k}.ply"], 28, idx)
        assert pts.dtype == torch.float32 and pts.device.type == "cpu" and tuple(gt.shape) == (1000, 29)
        assert np.array_equal(pts.numpy().view(np.uint32), want_p.view(np.uint32)) and np.array_equal(gt.numpy(), want_gt)
    with pytest.raises(IndexError):
        ds[3]
    with pytest.raises(ValueError):                     # np.random.choice: more samples than points (:141)
        CARLA_Seg(28, 5000, str(tmp_path))[0]
    with pytest.raises(FileNotFoundError):
        CARLA_Seg(28, 10, str(tmp_path / "missing"))
