"""oracle/ndt_oracle.py — TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/libndt_oracle.so (built from oracle/ndt_oracle.c by oracle/Makefile,
which __graft_entry__.build() runs).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libndt_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ndt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "oracle"])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ndt_oracle_create.restype = C.c_void_p
        L.ndt_oracle_destroy.argtypes = [C.c_void_p]
        L.ndt_oracle_run.restype = C.c_int
        L.ndt_oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_ushort, C.c_ulong]
        for name, rt in [("ret", C.c_int), ("evaluations", C.c_int), ("voxel_size", C.c_double), ("G", C.c_ulong),
                         ("num_nds", C.c_ulong), ("num_kl0", C.c_ulong), ("num_kl", C.c_ulong),
                         ("num_valid0", C.c_ulong), ("num_valid", C.c_ulong), ("num_out", C.c_ulong),
                         ("prune_ret", C.c_int), ("prune_walk", C.c_ulong)]:
            f = getattr(L, "ndt_oracle_get_" + name)
            f.restype, f.argtypes = rt, [C.c_void_p]
        for name in ["len", "offset", "limits", "point_voxel", "num_samples", "num_samples0", "mean", "cov", "cov0",
                     "cls", "kl_div0", "kl_p0", "kl_q0", "kl_div", "kl_p", "kl_q", "out_pts", "out_cov", "out_cls",
                     "out_voxel"]:
            f = getattr(L, "ndt_oracle_get_" + name)
            f.restype, f.argtypes = C.c_void_p, [C.c_void_p]
        L.ndt_oracle_limits.argtypes = [C.c_void_p, C.c_ulong, C.c_void_p]
        L.ndt_oracle_neighbor.restype = C.c_int
        L.ndt_oracle_neighbor.argtypes = [C.c_ulong, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.POINTER(C.c_ulong)]
        L.ndt_oracle_lu3.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.ndt_oracle_lu3_invert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ndt_oracle_lu3_det.restype = C.c_double
        L.ndt_oracle_lu3_det.argtypes = [C.c_void_p, C.c_int]
        L.ndt_oracle_kl_pair.restype = C.c_int
        L.ndt_oracle_kl_pair.argtypes = [C.c_ulong, C.c_void_p, C.c_ulong, C.c_void_p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _arr(ptr, dtype, shape):
    n = int(np.prod(shape))
    if not ptr or n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape).copy()


class OracleResult:
    pass


def limits(cloud: np.ndarray) -> np.ndarray:
    cloud = np.ascontiguousarray(cloud, np.float64)
    out = np.zeros(6, np.float64)
    lib().ndt_oracle_limits(cloud.ctypes.data, cloud.shape[0], out.ctypes.data)
    return out


def neighbor(index: int, lx: int, ly: int, lz: int, direction: int):
    out = C.c_ulong(0)
    r = lib().ndt_oracle_neighbor(index, lx, ly, lz, direction, C.byref(out))
    return r, out.value


def run(cloud: np.ndarray, num_desired: int, classes: np.ndarray | None = None, num_classes: int = 0) -> OracleResult:
    """Run the whole restated path on one cloud ([N,3], any float dtype; widened exactly to f64)."""
    L = lib()
    cloud = np.ascontiguousarray(cloud, dtype=np.float64)
    n = cloud.shape[0]
    cls_ptr = None
    if classes is not None:
        classes = np.ascontiguousarray(classes, dtype=np.uint16)
        cls_ptr = classes.ctypes.data
    h = L.ndt_oracle_create()
    try:
        r = OracleResult()
        r.ret = L.ndt_oracle_run(h, cloud.ctypes.data, n, cls_ptr, num_classes, num_desired)
        g = lambda name: getattr(L, "ndt_oracle_get_" + name)(h)
        r.evaluations = g("evaluations")
        r.voxel_size = g("voxel_size")
        r.lens = tuple(int(v) for v in _arr(g("len"), np.uint32, (3,)))
        r.offsets = _arr(g("offset"), np.float64, (3,))
        r.limits = _arr(g("limits"), np.float64, (6,))
        r.G = g("G")
        r.num_nds = g("num_nds")
        r.point_voxel = _arr(g("point_voxel"), np.int64, (n,))
        if r.ret != 0:
            return r
        G = r.G
        r.num_samples = _arr(g("num_samples"), np.uint64, (G,))
        r.num_samples0 = _arr(g("num_samples0"), np.uint64, (G,))
        r.mean = _arr(g("mean"), np.float64, (G, 3))
        r.cov = _arr(g("cov"), np.float64, (G, 9))
        r.cov0 = _arr(g("cov0"), np.float64, (G, 9))
        r.cls = _arr(g("cls"), np.uint16, (G,))
        r.num_kl0, r.num_kl = g("num_kl0"), g("num_kl")
        r.num_valid0, r.num_valid = g("num_valid0"), g("num_valid")
        r.prune_ret, r.prune_walk = g("prune_ret"), g("prune_walk")
        r.kl_div0 = _arr(g("kl_div0"), np.float64, (r.num_kl0,))
        r.kl_p0 = _arr(g("kl_p0"), np.int64, (r.num_kl0,))
        r.kl_q0 = _arr(g("kl_q0"), np.int64, (r.num_kl0,))
        r.num_out = g("num_out")
        r.out_pts = _arr(g("out_pts"), np.float64, (r.num_out, 3))
        r.out_cov = _arr(g("out_cov"), np.float64, (r.num_out, 9))
        r.out_cls = _arr(g("out_cls"), np.uint16, (r.num_out,))
        r.out_voxel = _arr(g("out_voxel"), np.int64, (r.num_out,))
        return r
    finally:
        L.ndt_oracle_destroy(h)
