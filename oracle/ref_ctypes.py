"""oracle/ref_ctypes.py — TEST INFRASTRUCTURE ONLY (never imported by the product path).

ctypes view of the *reference's own* legacy C ABI (core_legacy/include/ndnet_core/ndt.h:59-116,
normal_distributions.h:41-51, kullback_leibler.h:40-44).  It can load either a reference build
from oracle/_ref/ (made by oracle/Makefile from /root/reference/core_legacy/src, GSL replaced by
oracle/gsl_shim) or any other library exporting the same symbols.  Unlike the reference's
ndt_legacy.py it declares the real struct layouts, so tests can look inside the returned
nd_array / kl_divergences handles of a *reference* library (ours keeps them opaque).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class NormalDistribution(C.Structure):
    # core_legacy/include/ndnet_core/normal_distributions.h:41-51 (sizeof == 184)
    _fields_ = [
        ("index", C.c_ulong),
        ("mean", C.c_double * 3),
        ("old_mean", C.c_double * 3),
        ("covariance", C.c_double * 9),
        ("m2", C.c_double * 3),
        ("num_samples", C.c_ulong),
        ("cls", C.c_ushort),
        ("num_class_samples", C.POINTER(C.c_uint)),
        ("being_updated", C.c_bool),
    ]


class KLDivergence(C.Structure):
    # core_legacy/include/ndnet_core/kullback_leibler.h:40-44 (sizeof == 24)
    _fields_ = [("divergence", C.c_double), ("p", C.POINTER(NormalDistribution)), ("q", C.POINTER(NormalDistribution))]


assert C.sizeof(NormalDistribution) == 184 and C.sizeof(KLDivergence) == 24

ND_DTYPE = np.dtype(
    {
        "names": ["index", "mean", "old_mean", "covariance", "m2", "num_samples", "cls", "ncs_ptr", "being_updated"],
        "formats": ["<u8", ("<f8", 3), ("<f8", 3), ("<f8", 9), ("<f8", 3), "<u8", "<u2", "<u8", "u1"],
        "offsets": [0, 8, 32, 56, 128, 152, 160, 168, 176],
        "itemsize": 184,
    }
)


def ref_lib_path(variant: str = "det") -> str:
    name = {"det": "libndnet_ref_det.so", "threaded": "libndnet_ref.so", "O2": "libndnet_ref_O2.so"}[variant]
    return os.path.join(HERE, "_ref", name)


def have_ref(variant: str = "det") -> bool:
    return os.path.exists(ref_lib_path(variant))


_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
_ulp = C.POINTER(C.c_ulong)
_usp = C.POINTER(C.c_ushort)


def load(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    lib.ndt_downsample.restype = C.c_int
    lib.ndt_downsample.argtypes = [
        _dp, C.c_ushort, C.c_ulong, _up, _up, _up, _dp, _dp, _dp, _dp, _usp, C.c_ushort, C.c_ulong,
        _dp, _ulp, _dp, _usp, C.POINTER(C.c_void_p), _ulp, C.POINTER(C.c_void_p), _ulp,
    ]
    lib.prune_nds.restype = C.c_int
    lib.prune_nds.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_ulong, _ulp, C.c_void_p, _ulp]
    lib.to_point_cloud.restype = C.c_int
    lib.to_point_cloud.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double,
                                   C.c_double, _dp, _ulp, _dp, _usp]
    lib.free_nds.restype = None
    lib.free_nds.argtypes = [C.c_void_p, C.c_ulong]
    lib.free_kl_divergences.restype = None
    lib.free_kl_divergences.argtypes = [C.c_void_p]
    return lib


class Result:
    """Everything one ndt_downsample call returns, plus (for a reference build) the internals."""

    def __init__(self):
        self.ret = None
        self.lens = (0, 0, 0)
        self.offsets = (0.0, 0.0, 0.0)
        self.voxel_size = 0.0
        self.points = None
        self.covs = None
        self.classes = None
        self.num_out = 0
        self.num_valid = 0
        self.num_kl = 0
        self.nd = None           # structured copy of the nd_array (reference builds only)
        self.kl_div = None       # divergences in list order
        self.kl_p = None         # voxel index of p per list entry
        self.kl_q = None


def downsample(lib: C.CDLL, cloud: np.ndarray, num_desired: int, classes: np.ndarray | None = None,
               num_classes: int = 0, introspect: bool = False, out_rows: int | None = None) -> Result:
    """Call lib.ndt_downsample the way ndnet/preprocessing/ndt_legacy.py:111-171 does.

    `out_rows` over-allocates the output buffers (the reference can overrun `num_desired` rows,
    SURVEY.md A15); by default 1.25*num_desired+8 rows are provided and `num_out` tells how many
    the library claims to have written.
    """
    cloud = np.ascontiguousarray(cloud, dtype=np.float64)
    n = cloud.shape[0]
    rows = out_rows if out_rows is not None else int(num_desired * 1.25) + 8
    r = Result()
    pts = np.zeros((rows, 3), np.float64)
    covs = np.zeros((rows, 9), np.float64)
    cls_out = np.zeros(rows, np.uint16)
    lx, ly, lz = C.c_uint(0), C.c_uint(0), C.c_uint(0)
    ox, oy, oz, vs = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
    n_out, n_valid, n_kl = C.c_ulong(0), C.c_ulong(0), C.c_ulong(0)
    nd_ptr, kl_ptr = C.c_void_p(None), C.c_void_p(None)
    cls_ptr = None
    if classes is not None:
        classes = np.ascontiguousarray(classes, dtype=np.uint16)
        cls_ptr = classes.ctypes.data_as(_usp)
    r.ret = lib.ndt_downsample(
        cloud.ctypes.data_as(_dp), 3, n, C.byref(lx), C.byref(ly), C.byref(lz), C.byref(ox), C.byref(oy), C.byref(oz),
        C.byref(vs), cls_ptr, num_classes, num_desired, pts.ctypes.data_as(_dp), C.byref(n_out),
        covs.ctypes.data_as(_dp), cls_out.ctypes.data_as(_usp) if classes is not None else None,
        C.byref(nd_ptr), C.byref(n_valid), C.byref(kl_ptr), C.byref(n_kl))
    r.lens = (lx.value, ly.value, lz.value)
    r.offsets = (ox.value, oy.value, oz.value)
    r.voxel_size = vs.value
    r.num_out, r.num_valid, r.num_kl = n_out.value, n_valid.value, n_kl.value
    r.points, r.covs, r.classes = pts, covs, cls_out
    g = r.lens[0] * r.lens[1] * r.lens[2]
    if r.ret == 0 and introspect and nd_ptr.value:
        buf = (C.c_char * (g * 184)).from_address(nd_ptr.value)
        r.nd = np.frombuffer(buf, dtype=ND_DTYPE, count=g).copy()
        if kl_ptr.value and r.num_kl:
            kbuf = (C.c_char * (r.num_kl * 24)).from_address(kl_ptr.value)
            kl = np.frombuffer(kbuf, dtype=np.dtype([("d", "<f8"), ("p", "<u8"), ("q", "<u8")]), count=r.num_kl).copy()
            r.kl_div = kl["d"]
            r.kl_p = ((kl["p"] - nd_ptr.value) // 184).astype(np.int64)
            r.kl_q = ((kl["q"] - nd_ptr.value) // 184).astype(np.int64)
    if r.ret == 0:
        if nd_ptr.value:
            lib.free_nds(nd_ptr, g)
        if kl_ptr.value:
            lib.free_kl_divergences(kl_ptr)
    return r
