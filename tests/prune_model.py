"""Test infrastructure for the retained-handle continuation: the reference's prune walk
(core_legacy/src/ndt.c:45-72) restated on the oracle's sorted list, and the reference build driven call by call."""
import ctypes as C

import numpy as np

from oracle import ref_ctypes


class ListWalk:
    """ndt.c:45-72 restated on the oracle's sorted pre-prune list, continuation on the well-formed list."""

    def __init__(self, o):
        self.p = [int(v) for v in o.kl_p0]
        self.K = len(self.p)
        self.start = 0                       # first entry still in the list
        self.alive = set(int(v) for v in np.flatnonzero(o.num_samples0 > 0))
        self.n_valid = len(self.alive)
        self.num_kl = self.K                 # the reference's *num_kl_divergences
        self.skipped = False                 # some walk so far passed over an already-removed p

    def prune(self, desired):
        if desired > self.n_valid:
            return -1                        # ndt.c:36-39
        to_remove = self.n_valid - desired
        L = self.K - self.start
        idx = removed = 0
        while removed < to_remove:
            if idx >= L - removed:           # ndt.c:53 with *num_kl_divergences shrinking at :65
                self.n_valid -= removed
                self.num_kl -= removed
                self.start = self.K          # nothing behind this point can ever be removed (see test docstring)
                return -2
            v = self.p[self.start + idx]
            idx += 1
            if v not in self.alive:
                self.skipped = True
                continue
            self.alive.remove(v)
            removed += 1
        self.n_valid -= removed
        self.num_kl -= removed
        self.start += idx
        return 0

    def rows(self):
        return sorted(self.alive)


class RefHandles:
    """The reference build driven call by call with its handles kept (ref_ctypes.downsample frees them)."""

    def __init__(self, pts, labels, ncls, d):
        self.lib = ref_ctypes.load(ref_ctypes.ref_lib_path("det"))
        cloud = np.ascontiguousarray(pts, np.float64)
        self.has_labels = labels is not None
        cls = None if labels is None else np.ascontiguousarray(labels, np.uint16)
        rows = int(d * 1.25) + 8
        self.pts = np.zeros((rows, 3)); self.cov = np.zeros((rows, 9)); self.cls = np.zeros(rows, np.uint16)
        self.l = [C.c_uint(0) for _ in range(3)]
        self.o = [C.c_double(0) for _ in range(4)]
        self.n_out, self.n_valid, self.n_kl = C.c_ulong(0), C.c_ulong(0), C.c_ulong(0)
        self.nd, self.kl = C.c_void_p(None), C.c_void_p(None)
        dp, usp = ref_ctypes._dp, ref_ctypes._usp
        self.ret = self.lib.ndt_downsample(
            cloud.ctypes.data_as(dp), 3, cloud.shape[0], C.byref(self.l[0]), C.byref(self.l[1]), C.byref(self.l[2]),
            C.byref(self.o[0]), C.byref(self.o[1]), C.byref(self.o[2]), C.byref(self.o[3]),
            cls.ctypes.data_as(usp) if cls is not None else None, ncls, d, self.pts.ctypes.data_as(dp), C.byref(self.n_out),
            self.cov.ctypes.data_as(dp), self.cls.ctypes.data_as(usp) if cls is not None else None,
            C.byref(self.nd), C.byref(self.n_valid), C.byref(self.kl), C.byref(self.n_kl))

    def prune(self, d):
        dp, usp = ref_ctypes._dp, ref_ctypes._usp
        self.lib.prune_nds(self.nd, self.l[0], self.l[1], self.l[2], d, C.byref(self.n_valid), self.kl, C.byref(self.n_kl))
        rows = int(self.n_valid.value) + 8
        pts = np.zeros((rows, 3)); cov = np.zeros((rows, 9)); cls = np.zeros(rows, np.uint16)
        n = C.c_ulong(0)
        self.lib.to_point_cloud(self.nd, self.l[0], self.l[1], self.l[2], self.o[0], self.o[1], self.o[2], self.o[3],
                                pts.ctypes.data_as(dp), C.byref(n), cov.ctypes.data_as(dp),
                                cls.ctypes.data_as(usp) if self.has_labels else None)
        k = int(self.n_valid.value)
        return pts[:k], cov[:k], cls[:k]

    def close(self):
        g = self.l[0].value * self.l[1].value * self.l[2].value
        if self.nd.value:
            self.lib.free_nds(self.nd, g)
        if self.kl.value:
            self.lib.free_kl_divergences(self.kl)
