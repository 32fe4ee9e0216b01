"""Host-side driver of the batched NDT path: torch tensors in, torch tensors out, all work on the
current CUDA stream through the C ABI (include/ndnet_b200.h).  PyTorch is plumbing only (device memory
and streams)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class NdtBatch:
    feat: torch.Tensor            # [B, D, 12] f32 (mean, row-major 3x3 "covariance")
    labels: torch.Tensor | None   # [B, D] int16 view of the uint16 class per distribution
    voxel: torch.Tensor | None    # [B, D] int32 linear voxel index per row, -1 = padding
    feat64: torch.Tensor | None   # [B, D, 12] f64, exact values
    info: np.ndarray | None       # structured array (INFO_DTYPE), one record per cloud (host)


class NdtEngine:
    """One context (device workspace) per engine; reuse it across batches of the same shape."""

    def __init__(self, device: int | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("ndnet_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.ndnet_b200_create(C.byref(h), self.device.index or 0)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_create failed ({rc})")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ndnet_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._L.ndnet_b200_last_error(self._h).decode()}")

    def downsample(self, points: torch.Tensor, num_desired: int, labels: torch.Tensor | None = None,
                   num_classes: int = 0, nan_to_num: bool = True, want_f64: bool = False, want_voxel: bool = False,
                   want_info: bool = True, textbook_kl: bool = False) -> NdtBatch:
        """points: CUDA tensor [B, N, 3] f32 or f64; labels: CUDA int16/uint16 [B, N] or None.
        textbook_kl: the README's algorithm (true covariances and KL, least divergent removed first) instead of the
        compiled reference's behaviour (include/ndnet_b200.h NDNET_B200_TEXTBOOK_KL)."""
        assert points.is_cuda and points.dim() == 3 and points.shape[2] == 3
        assert points.dtype in (torch.float32, torch.float64)
        points = points.contiguous()
        B, N, _ = points.shape
        D = int(num_desired)
        dev = points.device
        feat = torch.empty((B, D, 12), dtype=torch.float32, device=dev)
        feat64 = torch.empty((B, D, 12), dtype=torch.float64, device=dev) if want_f64 else None
        out_lab = None
        lab_ptr = None
        if labels is not None:
            assert labels.is_cuda and labels.shape == (B, N) and labels.dtype in (torch.int16, torch.uint16)
            labels = labels.contiguous()
            lab_ptr = labels.data_ptr()
            out_lab = torch.empty((B, D), dtype=torch.int16, device=dev)
        voxel = torch.empty((B, D), dtype=torch.int32, device=dev) if want_voxel else None
        info_dev = torch.empty((B, _lib.INFO_DTYPE.itemsize), dtype=torch.uint8, device=dev) if want_info else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = self._L.ndnet_b200_downsample_batch(
                self._h, points.data_ptr(), _lib.F32 if points.dtype == torch.float32 else _lib.F64, lab_ptr, B, N,
                int(num_classes), D, (_lib.NAN_TO_NUM if nan_to_num else 0) | (_lib.TEXTBOOK_KL if textbook_kl else 0), feat.data_ptr(),
                feat64.data_ptr() if want_f64 else None, out_lab.data_ptr() if out_lab is not None else None,
                voxel.data_ptr() if want_voxel else None, info_dev.data_ptr() if want_info else None, stream)
        self._check(rc, "ndnet_b200_downsample_batch")
        info = None
        if want_info:
            info = np.frombuffer(info_dev.cpu().numpy().tobytes(), dtype=_lib.INFO_DTYPE).copy()
        return NdtBatch(feat, out_lab, voxel, feat64, info)

    def downsample_multiscale(self, points: torch.Tensor, num_desired_list, labels: torch.Tensor | None = None,
                              num_classes: int = 0, **kw) -> list:
        """Several n_desired_nds for the same clouds (the multiscale use of tools/train_multiscale.py:33-43,
        BASELINE config 5: 4096 / 1024 / 256 per cloud).  Each resolution is an independent ndt_downsample, exactly as
        the reference's dataset would run it; returns one NdtBatch per entry of `num_desired_list`."""
        return [self.downsample(points, int(d), labels, num_classes, **kw) for d in num_desired_list]

    def downsample_host(self, points: np.ndarray | torch.Tensor, num_desired: int, labels=None, num_classes: int = 0,
                        nan_to_num: bool = True, out_feat: torch.Tensor | None = None):
        """HOST buffers in, HOST buffers out (H2D + kernels + D2H + sync inside the C call)."""
        if isinstance(points, torch.Tensor):
            assert not points.is_cuda
            pts_ptr, dtype, (B, N, _) = points.data_ptr(), points.dtype, points.shape
            dt = _lib.F32 if dtype == torch.float32 else _lib.F64
        else:
            points = np.ascontiguousarray(points)
            pts_ptr, (B, N, _) = points.ctypes.data, points.shape
            dt = _lib.F32 if points.dtype == np.float32 else _lib.F64
        D = int(num_desired)
        if out_feat is None:
            out_feat = torch.empty((B, D, 12), dtype=torch.float32).pin_memory()
        out_lab = None
        lab_ptr = None
        if labels is not None:
            if isinstance(labels, torch.Tensor):
                lab_ptr = labels.data_ptr()
            else:
                labels = np.ascontiguousarray(labels, dtype=np.uint16)
                lab_ptr = labels.ctypes.data
            out_lab = np.zeros((B, D), np.uint16)
        info = np.zeros(B, _lib.INFO_DTYPE)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self._L.ndnet_b200_downsample_batch_host(
                self._h, pts_ptr, dt, lab_ptr, B, N, int(num_classes), D, _lib.NAN_TO_NUM if nan_to_num else 0,
                out_feat.data_ptr(), None, out_lab.ctypes.data if out_lab is not None else None, None,
                info.ctypes.data, stream)
        self._check(rc, "ndnet_b200_downsample_batch_host")
        return out_feat, out_lab, info

    # ---- inspection helpers used by the parity tests -------------------------------------------
    def keep_point_voxels(self, enable: bool = True) -> None:
        self._check(self._L.ndnet_b200_keep_point_voxels(self._h, 1 if enable else 0), "ndnet_b200_keep_point_voxels")

    def set_graph(self, mode: int) -> None:
        """CUDA-graph replay of the NDT chain: -1 automatic (small batches, the default), 0 never, 1 every batch
        (include/ndnet_b200.h ndnet_b200_set_ndt_graph)."""
        self._check(self._L.ndnet_b200_set_ndt_graph(self._h, int(mode)), "ndnet_b200_set_ndt_graph")

    def keep_kl_list(self, enable: bool = True) -> None:
        """The batched calls sort only the head of the divergence list; enable this before a batch whose whole sorted list
        `last_kl_list` is to return."""
        self._check(self._L.ndnet_b200_keep_kl_list(self._h, 1 if enable else 0), "ndnet_b200_keep_kl_list")

    def last_point_voxels(self, B: int, N: int) -> torch.Tensor:
        out = torch.empty((B, N), dtype=torch.int32, device=self.device)
        self._check(self._L.ndnet_b200_last_point_voxels(self._h, out.data_ptr(),
                                                         torch.cuda.current_stream(self.device).cuda_stream),
                    "ndnet_b200_last_point_voxels")
        return out

    def last_kl_list(self, b: int, cap: int):
        div = np.zeros(cap, np.float64); p = np.zeros(cap, np.int32); q = np.zeros(cap, np.int32)
        k = self._L.ndnet_b200_last_kl_list(self._h, b, div.ctypes.data, p.ctypes.data, q.ctypes.data, cap)
        if k < 0:
            raise RuntimeError(f"ndnet_b200_last_kl_list failed ({k})")
        k = min(k, cap)
        return div[:k], p[:k], q[:k]


_default: dict[int, NdtEngine] = {}


def default_engine(device: torch.device | int = 0) -> NdtEngine:
    idx = device if isinstance(device, int) else (device.index or 0)
    if idx not in _default:
        _default[idx] = NdtEngine(idx)
    return _default[idx]
