"""Host-side wrapper of the CUDA NDT-Net forward (ndnet_b200_model_* in include/ndnet_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import default_engine

KIND_CLS, KIND_SEG = 0, 1                       # NDT-Net heads (ndtnet.py)
KIND_POINTNET_CLS, KIND_POINTNET_SEG = 2, 3     # PointNet heads (pointnet.py)


def deterministic_state_dict(module: torch.nn.Module, seed: int = 0) -> dict:
    """Random-but-reproducible weights keyed by parameter NAME (independent of torch's init order), with
    non-trivial BatchNorm running statistics so that BN folding is actually exercised."""
    out = {}
    for name, ref in sorted(module.state_dict().items()):
        rng = np.random.default_rng([seed, abs(hash_name(name)) % (2 ** 31)])
        shape = tuple(ref.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.tensor(1, dtype=ref.dtype)
        elif name.endswith("running_var"):
            out[name] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif name.endswith("running_mean"):
            out[name] = torch.from_numpy(rng.normal(0, 0.1, shape).astype(np.float32))
        elif ".bn" in name and name.endswith("weight"):
            out[name] = torch.from_numpy(rng.uniform(0.8, 1.2, shape).astype(np.float32))
        elif name.endswith("bias"):
            out[name] = torch.from_numpy(rng.normal(0, 0.05, shape).astype(np.float32))
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            out[name] = torch.from_numpy(rng.normal(0, 1.0 / np.sqrt(fan_in), shape).astype(np.float32))
    return out


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


class B200Model:
    def __init__(self, module: torch.nn.Module, kind: int, device: torch.device | int = 0):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.engine = default_engine(self.device)
        self._L = _lib.lib()
        sd = {k: v.detach().float().cpu().contiguous() for k, v in module.state_dict().items()
              if not k.endswith("num_batches_tracked")}
        names = list(sd)
        n = len(names)
        self._keep = [sd[k].numpy() for k in names]
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_data = (C.c_void_p * n)(*[a.ctypes.data for a in self._keep])
        shapes = [np.array(a.shape if a.ndim else (1,), np.int64) for a in self._keep]
        c_shapes = (C.c_void_p * n)(*[s.ctypes.data for s in shapes])
        c_nd = (C.c_int * n)(*[len(s) for s in shapes])
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._L.ndnet_b200_model_create(self.engine.handle, C.byref(h), kind, n, c_names, c_data, c_shapes, c_nd)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_model_create failed ({rc}): "
                               f"{self._L.ndnet_b200_last_error(self.engine.handle).decode()}")
        self._h = h
        self.kind = kind & 1                       # 0 classification, 1 segmentation
        self.in_dim = int(self._L.ndnet_b200_model_input_dim(h))
        self.n_out = int(module.num_classes) + (1 if self.kind == KIND_SEG else 0)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.ndnet_b200_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def __call__(self, feat: torch.Tensor) -> torch.Tensor:
        """feat: CUDA f32 [B, D, in_dim] -> cls: [B, n_out, 1] probabilities; seg: [B, D, n_out] log-probabilities."""
        assert feat.is_cuda and feat.dtype == torch.float32 and feat.dim() == 3 and feat.shape[2] == self.in_dim
        feat = feat.contiguous()
        B, D, _ = feat.shape
        if self.kind == KIND_SEG:
            out = torch.empty((B, D, self.n_out), dtype=torch.float32, device=feat.device)
        else:
            out = torch.empty((B, self.n_out), dtype=torch.float32, device=feat.device)
        stream = torch.cuda.current_stream(feat.device).cuda_stream
        with torch.cuda.device(feat.device):
            rc = self._L.ndnet_b200_model_forward(self.engine.handle, self._h, feat.data_ptr(), B, D, out.data_ptr(), stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_model_forward failed ({rc}): "
                               f"{self._L.ndnet_b200_last_error(self.engine.handle).decode()}")
        return out.unsqueeze(-1) if self.kind == KIND_CLS else out


    def infer_host(self, points: torch.Tensor, num_desired: int, labels: torch.Tensor | None, num_classes: int,
                   out: torch.Tensor, wait: bool = True) -> torch.Tensor:
        """HOST tensors in (pinned for speed), HOST tensor out: H2D + NDT + forward + D2H + sync in one C call.
        labels: int16/uint16 [B, N] (the reference's dtype) or uint8 [B, N] (one byte per point, num_classes <= 255).
        wait=False: returns once everything is enqueued (several batches in flight: the next batch's copies overlap this
        one's kernels); `out` is valid after `infer_wait()`, and every batch in flight needs its own `out`."""
        assert not points.is_cuda and points.is_contiguous() and not out.is_cuda
        B, N, _ = points.shape
        per_cloud = out.numel() // B
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = self._L.ndnet_b200_infer_host
        if labels is not None:
            assert labels.is_contiguous() and labels.dtype in (torch.int16, torch.uint16, torch.uint8)
            if labels.dtype == torch.uint8:
                fn = self._L.ndnet_b200_infer_host_u8
        if not wait:
            with torch.cuda.device(self.device):
                rc = self._L.ndnet_b200_infer_host_async(
                    self.engine.handle, self._h, points.data_ptr(), _lib.F32 if points.dtype == torch.float32 else _lib.F64,
                    labels.data_ptr() if labels is not None else None, int(labels is not None and labels.dtype == torch.uint8),
                    B, N, int(num_classes), int(num_desired), out.data_ptr(), per_cloud, stream)
            if rc != 0:
                raise RuntimeError(f"ndnet_b200_infer_host_async failed ({rc}): "
                                   f"{self._L.ndnet_b200_last_error(self.engine.handle).decode()}")
            return out
        with torch.cuda.device(self.device):
            rc = fn(
                self.engine.handle, self._h, points.data_ptr(), _lib.F32 if points.dtype == torch.float32 else _lib.F64,
                labels.data_ptr() if labels is not None else None, B, N, int(num_classes), int(num_desired),
                out.data_ptr(), per_cloud, stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_infer_host failed ({rc}): "
                               f"{self._L.ndnet_b200_last_error(self.engine.handle).decode()}")
        return out


    def infer_wait(self) -> None:
        """Blocks until every batch enqueued with infer_host(..., wait=False) on the current stream has its result in host memory."""
        rc = self._L.ndnet_b200_infer_wait(self.engine.handle, torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_infer_wait failed ({rc})")

    def infer_device(self, points: torch.Tensor, num_desired: int, labels: torch.Tensor | None = None,
                     num_classes: int = 0) -> torch.Tensor:
        """CUDA tensors in, CUDA tensor out, asynchronous on the current stream (chunks pipelined on internal lanes)."""
        assert points.is_cuda and points.is_contiguous()
        B, N, _ = points.shape
        D = int(num_desired)
        if self.kind == KIND_SEG:
            out = torch.empty((B, D, self.n_out), dtype=torch.float32, device=points.device)
        else:
            out = torch.empty((B, self.n_out), dtype=torch.float32, device=points.device)
        stream = torch.cuda.current_stream(points.device).cuda_stream
        with torch.cuda.device(points.device):
            rc = self._L.ndnet_b200_infer_device(
                self.engine.handle, self._h, points.data_ptr(), _lib.F32 if points.dtype == torch.float32 else _lib.F64,
                labels.data_ptr() if labels is not None else None, B, N, int(num_classes), D, out.data_ptr(),
                out.numel() // B, stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_infer_device failed ({rc}): "
                               f"{self._L.ndnet_b200_last_error(self.engine.handle).decode()}")
        return out.unsqueeze(-1) if self.kind == KIND_CLS else out

    def set_pipeline(self, lanes: int, chunk: int, device_chunk: int | None = None, stagger: bool | None = None) -> None:
        """lanes x chunk scans for infer_host; `device_chunk` scans per chunk for infer_device (library default 128);
        `stagger`: a chunk's front (limits, search, voxel assignment) starts behind the front of the chunk before it."""
        if stagger is not None:
            self._L.ndnet_b200_set_stagger(self.engine.handle, 1 if stagger else 0)
        rc = self._L.ndnet_b200_set_pipeline(self.engine.handle, int(lanes), int(chunk))
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_set_pipeline failed ({rc})")
        if device_chunk is not None:
            rc = self._L.ndnet_b200_set_device_chunk(self.engine.handle, int(device_chunk))
            if rc != 0:
                raise RuntimeError(f"ndnet_b200_set_device_chunk failed ({rc})")


def smoke_forward(engine, feat: torch.Tensor) -> None:
    """Tiny forward of the segmentation network on the features smoke() just produced, vs torch fp32."""
    from ndnet.models.ndtnet import NDTNetSegmentation
    torch.manual_seed(0)
    net = NDTNetSegmentation(num_classes=28, feature_dim=1024)
    net.load_state_dict(deterministic_state_dict(net, 0))
    net = net.cuda().eval()
    f = torch.nan_to_num(feat.float(), nan=0.0, posinf=0.0, neginf=0.0)
    with torch.no_grad():
        ref = net.forward_torch(f[:, :, :3], f[:, :, 3:])
        got = net(f[:, :, :3], f[:, :, 3:])                 # model(points, covs): dispatches to the CUDA path
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    agree = (got.argmax(-1) == ref.argmax(-1)).float().mean().item()
    assert err <= 2e-2 * scale and agree > 0.97, (err, scale, agree)   # bf16 operands: 2 % of the log-probability range
