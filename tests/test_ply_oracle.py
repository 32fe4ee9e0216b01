"""CPU: the PLY oracle (oracle/ply_oracle.py) against the fixture produced by the reference's own reader
(tests/golden/make_ply_golden.py ran /root/reference/ndnet/datasets/CARLA_Seg.py:96-183 here)."""
import os

import numpy as np
import pytest

from oracle import ply_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_the_reference_reader():
    g = np.load(os.path.join(GOLD, "ply_ref_golden.npz"))
    with open(os.path.join(GOLD, "carla_like.ply"), "rb") as f:
        raw = f.read()
    pts, gt, tags = ply_oracle.get_data_pcl(raw, int(g["n_classes"]), g["indexes"])
    assert np.array_equal(pts.view(np.uint32), g["points"].view(np.uint32))
    assert np.array_equal(gt, g["gt"])
    assert np.array_equal(np.argmax(gt, 1), tags)


def test_oracle_error_behaviour():
    head = b"h\n" * 10
    with pytest.raises(IndexError):
        ply_oracle.parse(head + b"1 2\n", 5)
    with pytest.raises(ValueError):
        ply_oracle.parse(head + b"1 x 3 0\n", 5)
    with pytest.raises(ValueError, match="Class tag 6 out of bounds"):
        ply_oracle.parse(head + b"1 2 3 6\n", 5)
    with pytest.raises(OverflowError):
        ply_oracle.parse(head + b"1 2 3 -1\n", 5)
    p, c = ply_oracle.parse(head + b"1 2 3 4\r\n5 6 7 0", 5)
    assert p.tolist() == [[1, 2, 3], [5, 6, 7]] and c.tolist() == [4, 0]
