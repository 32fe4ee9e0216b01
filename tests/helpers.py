"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import numpy as np


def same_bits(a, b):
    """Bit-exact equality of float64 arrays where any NaN equals any NaN (payload/sign of NaN is not
    part of the contract: x86 and sm_100a produce different NaN bit patterns)."""
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    if a.shape != b.shape:
        return False
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def first_diff(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    bad = ~((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b)))
    idx = np.argwhere(bad)
    return None if len(idx) == 0 else (tuple(idx[0]), a[tuple(idx[0])], b[tuple(idx[0])], int(bad.sum()))
