"""ASCII-PLY ingest on the GPU: host mirror of ndnet_b200_ply_* (include/ndnet_b200.h (3)).

Replaces the per-line Python loop of the reference reader /root/reference/ndnet/datasets/CARLA_Seg.py:96-183 and
raises what that loop raises (IndexError for a short line, ValueError for a bad literal or an out-of-bounds class
tag, OverflowError for a negative tag).  No CPU parsing path: without the CUDA library the import fails.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_MESSAGES = {
    -304: "literal outside what ndnet_b200 converts exactly (inf/nan/underscores, more than 19 significant digits, "
          "or a decimal exponent beyond 1e+-55)",
    -306: "lone carriage-return line ends or non-ASCII bytes are not supported by the GPU PLY reader",
}


class PlyCloud:
    """One parsed PLY body, resident on the device: `num_points` rows of (x, y, z) float32 and a uint16 class tag."""

    def __init__(self, raw, n_classes: int, num_header_lines: int = 10, device: int | torch.device = 0):
        dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ndnet_b200 parses PLY text on a CUDA device only")
        self.device = dev
        self.n_classes = int(n_classes)
        self._L = _lib.lib()
        self._handle = C.c_void_p()
        on_device = isinstance(raw, torch.Tensor)
        if on_device:
            assert raw.is_cuda and raw.dtype == torch.uint8 and raw.is_contiguous()
            ptr, nbytes = raw.data_ptr(), raw.numel()
        else:
            raw = bytes(raw)
            self._keep = raw
            ptr, nbytes = C.cast(C.c_char_p(raw), C.c_void_p).value, len(raw)
        n, bad_line, bad_value = C.c_ulong(0), C.c_long(-1), C.c_long(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        r = self._L.ndnet_b200_ply_load(dev.index or 0, ptr, nbytes, int(on_device), int(num_header_lines), self.n_classes,
                                        stream, C.byref(self._handle), C.byref(n), C.byref(bad_line), C.byref(bad_value))
        self._keep = None
        if r != 0:
            self._handle = C.c_void_p()
            where = f" (line {bad_line.value + 1} of the file)" if bad_line.value >= 0 else ""
            if r == -301:
                raise IndexError("list index out of range" + where)
            if r == -302:
                raise ValueError("could not convert a token of the PLY body" + where)
            if r == -303:
                raise ValueError(f"Class tag {bad_value.value} out of bounds")        # CARLA_Seg.py:128
            if r == -305:
                raise OverflowError(f"Python integer {bad_value.value} out of bounds for uint16" + where)
            if r in _MESSAGES:
                raise ValueError(_MESSAGES[r] + where)
            raise RuntimeError(f"ndnet_b200_ply_load failed with {r}")
        self.num_points = int(n.value)

    def sample(self, indexes=None, one_hot: bool = True):
        """Rows `indexes` (None = all, file order) -> (points f32 [n,3], gt f32 [n, n_classes+1] or None, tags u16 [n]),
        all on the device (CARLA_Seg.py:141-147,173-179)."""
        if self._handle.value is None:
            raise RuntimeError("PlyCloud is closed")
        idx_ptr, on_device, keep = None, 0, None
        if indexes is None:
            n = self.num_points
        elif isinstance(indexes, torch.Tensor) and indexes.is_cuda:
            keep = indexes.to(torch.int64).contiguous()
            idx_ptr, on_device, n = keep.data_ptr(), 1, keep.numel()
        else:
            keep = np.ascontiguousarray(np.asarray(indexes), dtype=np.int64)
            idx_ptr, n = keep.ctypes.data, keep.size
        pts = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        tags = torch.empty((n,), dtype=torch.uint16, device=self.device)
        gt = torch.empty((n, self.n_classes + 1), dtype=torch.float32, device=self.device) if one_hot else None
        r = self._L.ndnet_b200_ply_sample(self._handle, idx_ptr, n, on_device, pts.data_ptr(), tags.data_ptr(),
                                          gt.data_ptr() if one_hot else None, torch.cuda.current_stream(self.device).cuda_stream)
        if r == -307:
            raise IndexError("sample index out of range")
        if r != 0:
            raise RuntimeError(f"ndnet_b200_ply_sample failed with {r}")
        return pts, gt, tags

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value is not None:
            self._L.ndnet_b200_ply_free(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_ply(path: str, n_classes: int, num_header_lines: int = 10, device: int | torch.device = 0) -> PlyCloud:
    with open(path, "rb") as f:
        raw = f.read()
    return PlyCloud(raw, n_classes, num_header_lines, device)
