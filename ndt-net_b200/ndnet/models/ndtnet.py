"""Drop-in for the reference's ndnet/models/ndtnet.py: the same classes, constructor arguments, forward
signatures, outputs and - so that reference checkpoints load unchanged - the same parameter / buffer names
(reference: ndtnet.py:17-30 TNet, :100-109 NDTNet, :177-179 classification head, :209-216 segmentation head).

`model(points, covariances)` - the call the reference's scripts make (tools/seg_viz.py:133, tools/train.py:69) - runs
the network through libndnet_b200.so whenever the inputs are CUDA tensors: `forward` dispatches to `forward_b200`
(tcgen05/TMA bf16 GEMMs with fused bias/BN/ReLU/max-pool epilogues in eval mode; the training kernels of train.cu, with
gradients, when the module is in training mode - segmentation and classification heads alike).  `forward_torch` is the
plain PyTorch fp32 definition of the same network: it is what the parity tests compare the CUDA path against and what CPU
tensors get.
Set `module.b200 = False` (or NDNET_B200_FORWARD=torch in the environment) to force the PyTorch definition.
"""
from __future__ import annotations

import os
from enum import Enum

import torch
from torch import nn


def _use_library(module: nn.Module, x: torch.Tensor, has_training_kernels: bool) -> bool:
    """Whether `module(x, ...)` runs through libndnet_b200.so: CUDA input, not opted out, and - in training mode - only
    when the library has the backward of this head (otherwise autograd needs the PyTorch graph)."""
    if not x.is_cuda or not getattr(module, "b200", True) or os.environ.get("NDNET_B200_FORWARD", "") == "torch":
        return False
    if not getattr(module, "_b200_supported", lambda: True)():
        return False
    return has_training_kernels or not module.training


def _pointwise(n_in: int, n_out: int) -> nn.Conv1d:
    return nn.Conv1d(n_in, n_out, kernel_size=1)


class TNet(nn.Module):
    """Transformation network: shared MLP in->64->128->1024, max-pool, FC 1024->512->256->in*in, + identity."""

    def __init__(self, in_dim: int = 64) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.conv1, self.conv2, self.conv3 = _pointwise(in_dim, 64), _pointwise(64, 128), _pointwise(128, 1024)
        self.fc1, self.fc2, self.fc3 = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, in_dim * in_dim)
        self.relu = nn.ReLU()
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.bn4, self.bn5 = nn.BatchNorm1d(512), nn.BatchNorm1d(256)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (batch, in_dim, num_points) -> (batch, in_dim, in_dim)"""
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            x = self.relu(bn(conv(x)))
        x = x.amax(dim=2)
        x = self.relu(self.bn4(self.fc1(x)))
        x = self.relu(self.bn5(self.fc2(x)))
        x = self.fc3(x) + torch.eye(self.in_dim, device=x.device, dtype=x.dtype).reshape(1, -1)
        return x.reshape(-1, self.in_dim, self.in_dim)


class NDTNet(nn.Module):
    """Feature extractor over 12-D (mean + flattened 3x3 covariance) points."""

    class AdditionalFeatures(Enum):
        NONE = "none"
        COVARIANCES = "covariances"
        FEATURE_VECTOR = "feature_vector"

    def __init__(self, point_dim: int = 3, feature_dim: int = 768,
                 extra_type: "NDTNet.AdditionalFeatures" = None) -> None:
        super().__init__()
        extra_type = NDTNet.AdditionalFeatures.COVARIANCES if extra_type is None else extra_type
        self.point_dim, self.feature_dim = point_dim, feature_dim
        self.extra_dim = {NDTNet.AdditionalFeatures.COVARIANCES: point_dim ** 2,
                          NDTNet.AdditionalFeatures.FEATURE_VECTOR: feature_dim + point_dim ** 2,
                          NDTNet.AdditionalFeatures.NONE: 0}[extra_type]
        self.conv1 = _pointwise(point_dim + self.extra_dim, 64)
        self.conv2, self.conv3 = _pointwise(64, 128), _pointwise(128, feature_dim)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(feature_dim)
        self.t1, self.t2 = TNet(in_dim=point_dim), TNet(in_dim=64)

    def forward(self, points: torch.Tensor, extra: torch.Tensor):
        """points (B, N, point_dim), extra (B, N, extra_dim) -> (features (B, feature_dim, N), x_t2 (B, 64, N))"""
        B, N, d = points.shape
        t = self.t1(points.transpose(1, 2))                           # (B, d, d)
        p = torch.bmm(points, t.transpose(1, 2))                      # rows: T p
        cov = torch.matmul(t.unsqueeze(1), extra.reshape(B, N, d, d)) # T Sigma (not T Sigma T^T)
        x = torch.cat((p, cov.reshape(B, N, d * d)), dim=2).transpose(1, 2)
        x = self.bn1(self.conv1(x))                                   # no ReLU in the trunk
        t = self.t2(x)
        x_t2 = torch.bmm(x.transpose(1, 2), t).transpose(1, 2)
        x = self.bn2(self.conv2(x_t2))
        x = self.bn3(self.conv3(x))
        return x, x_t2


def _state_stamp(module: nn.Module):
    """Changes whenever a parameter or buffer was modified in place (optimizer step, load_state_dict, BatchNorm update) or
    replaced: the eval-mode CUDA model holds folded copies and must be rebuilt then."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


class _B200Mixin:
    """Builds (once) and runs the CUDA model for an eval-mode module; `forward` = the reference's call signature."""
    _kind = 0
    b200 = True

    def _b200_supported(self) -> bool:
        """Shapes the CUDA kernels are built for: xyz points + 3x3 covariances; at most 32 segmentation outputs."""
        return self.point_dim == 3 and self.feature_extractor.extra_dim == 9 and (self._kind == 0 or self.num_classes + 1 <= 32)

    def forward(self, points: torch.Tensor, covariances: torch.Tensor) -> torch.Tensor:
        if _use_library(self, points, has_training_kernels=True):
            return self.forward_b200(points, covariances)
        return self.forward_torch(points, covariances)

    def forward_b200(self, points: torch.Tensor, covariances: torch.Tensor) -> torch.Tensor:
        from ndnet_b200.model import B200Model
        if self.training:
            from ndnet_b200.train import NetTrainer          # train-mode BatchNorm + gradients from train.cu
            tr = getattr(self, "_b200_trainer", None)
            if tr is None or tr.device != points.device:
                tr = NetTrainer(self, points.device, tf32=bool(getattr(self, "b200_tf32", False)))
                object.__setattr__(self, "_b200_trainer", tr)
            tr.overlap_allreduce = bool(getattr(self, "b200_overlap_allreduce", False))
            return tr(points, covariances)
        m = getattr(self, "_b200_model", None)
        stamp = _state_stamp(self)
        if m is None or m.device != points.device or getattr(self, "_b200_stamp", None) != stamp:
            m = B200Model(self, self._kind, points.device)       # folds BatchNorm into bf16 weights: rebuilt when the state changed
            object.__setattr__(self, "_b200_model", m)
            object.__setattr__(self, "_b200_stamp", stamp)
        feat = torch.cat((points, covariances), dim=2).float().contiguous()
        return m(feat)


class NDTNetClassification(_B200Mixin, nn.Module):
    _kind = 0

    def __init__(self, point_dim: int = 3, num_classes: int = 512, feature_dim: int = 768) -> None:
        super().__init__()
        self.point_dim, self.num_classes, self.feature_dim = point_dim, num_classes, feature_dim
        self.feature_extractor = NDTNet(point_dim, feature_dim=feature_dim)
        self.conv1, self.conv2, self.conv3 = _pointwise(feature_dim, 512), _pointwise(512, 256), _pointwise(256, num_classes)

    def forward_torch(self, points: torch.Tensor, covariances: torch.Tensor) -> torch.Tensor:
        """-> (B, num_classes, 1) class probabilities"""
        x, _ = self.feature_extractor(points, covariances)
        x = x.amax(dim=2, keepdim=True)
        x = torch.relu(self.conv1(x))
        x = torch.relu(self.conv2(x))
        return torch.softmax(self.conv3(x), dim=1)


class NDTNetSegmentation(_B200Mixin, nn.Module):
    _kind = 1

    def __init__(self, point_dim: int = 3, num_classes: int = 16, feature_dim: int = 1024) -> None:
        super().__init__()
        self.point_dim, self.num_classes, self.feature_dim = point_dim, num_classes, feature_dim
        self.feature_extractor = NDTNet(point_dim, feature_dim=feature_dim)
        self.conv1, self.conv2 = _pointwise(feature_dim + 64, 512), _pointwise(512, 256)
        self.conv3, self.conv4 = _pointwise(256, 128), _pointwise(128, num_classes + 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(512), nn.BatchNorm1d(256), nn.BatchNorm1d(128)

    def forward_torch(self, points: torch.Tensor, covariances: torch.Tensor) -> torch.Tensor:
        """-> (B, N, num_classes + 1) per-distribution log-probabilities"""
        x, x_t2 = self.feature_extractor(points, covariances)
        g = x.amax(dim=2, keepdim=True).expand(-1, -1, x_t2.shape[2])
        x = torch.cat((x_t2, g), dim=1)
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.relu(self.bn2(self.conv2(x)))
        x = torch.relu(self.bn3(self.conv3(x)))
        return torch.nn.functional.log_softmax(self.conv4(x), dim=1).transpose(1, 2)
