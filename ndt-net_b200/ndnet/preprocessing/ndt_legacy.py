"""Drop-in for the reference's ndnet/preprocessing/ndt_legacy.py (same module path, class and method
names, argument meaning and return values; reference lines cited below), bound to libndnet_b200.so —
the B200 library that exports the legacy `libndnet.so` symbols — instead of /usr/local/lib/libndnet.so
(ndt_legacy.py:28).  All arithmetic runs on the GPU; the nd_array / kl_divergences handles are opaque
tokens owned by the library.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ndnet_b200 import _lib

core = _lib.lib()   # raises if the CUDA library is missing: there is no CPU fallback

_dbl_p = ctypes.POINTER(ctypes.c_double)
_uint_p = ctypes.POINTER(ctypes.c_uint)
_ulong_p = ctypes.POINTER(ctypes.c_ulong)
_ushort_p = ctypes.POINTER(ctypes.c_ushort)

# signatures: core_legacy/include/ndnet_core/ndt.h:59-116 (mirrors ndt_legacy.py:31-43,186-191,215-223)
core.ndt_downsample.restype = ctypes.c_int
core.ndt_downsample.argtypes = [
    _dbl_p, ctypes.c_ushort, ctypes.c_ulong, _uint_p, _uint_p, _uint_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p,
    _ushort_p, ctypes.c_ushort, ctypes.c_ulong, _dbl_p, _ulong_p, _dbl_p, _ushort_p,
    ctypes.POINTER(ctypes.c_void_p), _ulong_p, ctypes.POINTER(ctypes.c_void_p), _ulong_p]
core.prune_nds.restype = ctypes.c_int
core.prune_nds.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_ulong, _ulong_p,
                           ctypes.c_void_p, _ulong_p]
core.to_point_cloud.restype = ctypes.c_int
core.to_point_cloud.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_double,
                                ctypes.c_double, ctypes.c_double, ctypes.c_double, _dbl_p, _ulong_p, _dbl_p, _ushort_p]
core.free_nds.restype = None
core.free_nds.argtypes = [ctypes.c_void_p, ctypes.c_ulong]
core.free_kl_divergences.restype = None
core.free_kl_divergences.argtypes = [ctypes.c_void_p]


class NDT_Sampler:
    """NDT down-sampler over one point cloud (reference: ndt_legacy.py:45-240).

    pointcloud: float64 [N, 3]; classes: uint16 [N] or None; num_classes: labels are 0..num_classes.
    """

    def __init__(self, pointcloud: np.ndarray, classes: np.ndarray = None, num_classes: int = None) -> None:
        self.pointcloud = np.ascontiguousarray(pointcloud, dtype=np.float64)
        self.covariances = None
        self.classes = None if classes is None else np.ascontiguousarray(classes, dtype=np.uint16)
        self.num_classes = 0 if num_classes is None else int(num_classes)
        self.num_points = len(pointcloud)
        self.num_valid_nds = ctypes.c_ulong(0)
        self.len_x, self.len_y, self.len_z = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_uint(0)
        self.offset_x, self.offset_y, self.offset_z = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        self.voxel_size = ctypes.c_double(0)
        self.nd_array_ptr = ctypes.c_void_p(None)
        self.kl_divergences_ptr = ctypes.c_void_p(None)
        self.num_kl_divergences = ctypes.c_ulong(0)
        self.status = None
        self.destroyed = False

    def cleanup(self) -> None:
        """Release the library-side state (ndt_legacy.py:84-92).  Safe after a failed downsample,
        where the reference would dereference NULL."""
        if self.destroyed:
            return
        cells = int(self.len_x.value) * int(self.len_y.value) * int(self.len_z.value)
        core.free_nds(self.nd_array_ptr, cells)
        core.free_kl_divergences(self.kl_divergences_ptr)
        self.nd_array_ptr = ctypes.c_void_p(None)
        self.kl_divergences_ptr = ctypes.c_void_p(None)
        self.destroyed = True

    def __del__(self) -> None:
        try:
            self.cleanup()
        except Exception:
            pass

    def downsample(self, num_desired_points: int):
        """-> (points f64 [D,3], covariances f64 [D,9], classes u16 [D])   (ndt_legacy.py:111-171)"""
        d = int(num_desired_points)
        new_pcl = np.zeros((d, 3), dtype=np.float64)
        covariances = np.zeros((d, 9), dtype=np.float64)
        new_classes = np.zeros(d, dtype=np.uint16)
        n_out = ctypes.c_ulong(0)
        cls_ptr = None if self.classes is None else self.classes.ctypes.data_as(_ushort_p)
        self.destroyed = False
        self.status = core.ndt_downsample(
            self.pointcloud.ctypes.data_as(_dbl_p), 3, self.num_points,
            ctypes.byref(self.len_x), ctypes.byref(self.len_y), ctypes.byref(self.len_z),
            ctypes.byref(self.offset_x), ctypes.byref(self.offset_y), ctypes.byref(self.offset_z),
            ctypes.byref(self.voxel_size), cls_ptr, self.num_classes, d,
            new_pcl.ctypes.data_as(_dbl_p), ctypes.byref(n_out), covariances.ctypes.data_as(_dbl_p),
            new_classes.ctypes.data_as(_ushort_p),
            ctypes.byref(self.nd_array_ptr), ctypes.byref(self.num_valid_nds),
            ctypes.byref(self.kl_divergences_ptr), ctypes.byref(self.num_kl_divergences))
        self.num_points = d
        return new_pcl, covariances, new_classes

    def prune(self, new_desired_points: int):
        """Prune the retained distributions further (ndt_legacy.py:173-240)."""
        d = int(new_desired_points)
        core.prune_nds(self.nd_array_ptr, self.len_x.value, self.len_y.value, self.len_z.value, d,
                       ctypes.byref(self.num_valid_nds), self.kl_divergences_ptr, ctypes.byref(self.num_kl_divergences))
        rows = max(d, int(self.num_valid_nds.value))
        new_pcl = np.zeros((rows, 3))
        covariances = np.zeros((rows, 9), dtype=np.float64)
        new_classes = np.zeros(rows, dtype=np.uint16)
        n_out = ctypes.c_ulong(0)
        core.to_point_cloud(self.nd_array_ptr, self.len_x.value, self.len_y.value, self.len_z.value,
                            self.offset_x.value, self.offset_y.value, self.offset_z.value, self.voxel_size.value,
                            new_pcl.ctypes.data_as(_dbl_p), ctypes.byref(n_out), covariances.ctypes.data_as(_dbl_p),
                            new_classes.ctypes.data_as(_ushort_p))
        new_pcl, covariances, new_classes = new_pcl[:d], covariances[:d], new_classes[:d]
        self.num_points = d
        self.pointcloud, self.covariances, self.classes = new_pcl, covariances, new_classes
        return new_pcl, covariances, new_classes.astype(np.int16)
