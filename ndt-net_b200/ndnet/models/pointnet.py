"""Drop-in for the reference's ndnet/models/pointnet.py: PointNet without the covariance branch, for points of
any width `point_dim` (e.g. the 12-D mean+covariance points of tools/train_pointnet.py).  Same classes,
arguments, outputs and parameter names (reference: pointnet.py:65-98 PointNet, :137-149 classification head,
:169-186 segmentation head).  `model(points)` runs the CUDA path (`forward_b200`) for CUDA tensors - the folded inference
kernels in eval mode, the training kernels of train.cu (with gradients) in training mode - and the plain PyTorch fp32
definition (`forward_torch`) otherwise - see ndtnet.py."""
from __future__ import annotations

import torch
from torch import nn

from .ndtnet import TNet, _pointwise, _use_library


class PointNet(nn.Module):
    def __init__(self, point_dim: int = 3, feature_dim: int = 768) -> None:
        super().__init__()
        self.point_dim, self.feature_dim = point_dim, feature_dim
        self.conv1, self.conv2, self.conv3 = _pointwise(point_dim, 64), _pointwise(64, 128), _pointwise(128, feature_dim)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(feature_dim)
        self.t1, self.t2 = TNet(in_dim=point_dim), TNet(in_dim=64)

    def forward(self, x: torch.Tensor):
        """x (B, N, point_dim) -> (features (B, feature_dim, N), x_t2 (B, 64, N))"""
        x = x.transpose(1, 2)
        x = torch.nan_to_num(torch.bmm(self.t1(x), x), nan=0.0)
        x = self.bn1(self.conv1(x))
        x_t2 = torch.bmm(x.transpose(1, 2), self.t2(x)).transpose(1, 2)
        x = self.bn3(self.conv3(self.bn2(self.conv2(x_t2))))
        return x, x_t2


class _B200PointMixin:
    _kind = 2
    b200 = True

    def _b200_supported(self) -> bool:
        return 1 <= self.point_dim <= 16 and (self._kind == 2 or self.num_classes + 1 <= 32)

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        if _use_library(self, points, has_training_kernels=True):
            return self.forward_b200(points)
        return self.forward_torch(points)

    def forward_b200(self, points: torch.Tensor) -> torch.Tensor:
        from ndnet_b200.model import B200Model
        if self.training:
            from ndnet_b200.train import NetTrainer          # train-mode BatchNorm + gradients from train.cu
            tr = getattr(self, "_b200_trainer", None)
            if tr is None or tr.device != points.device:
                tr = NetTrainer(self, points.device, tf32=bool(getattr(self, "b200_tf32", False)))
                object.__setattr__(self, "_b200_trainer", tr)
            tr.overlap_allreduce = bool(getattr(self, "b200_overlap_allreduce", False))
            return tr(points)
        from .ndtnet import _state_stamp
        m = getattr(self, "_b200_model", None)
        stamp = _state_stamp(self)
        if m is None or m.device != points.device or getattr(self, "_b200_stamp", None) != stamp:
            m = B200Model(self, self._kind, points.device)       # rebuilt when a parameter or buffer changed
            object.__setattr__(self, "_b200_model", m)
            object.__setattr__(self, "_b200_stamp", stamp)
        return m(points.float().contiguous())


class PointNetClassification(_B200PointMixin, nn.Module):
    _kind = 2

    def __init__(self, point_dim: int = 3, num_classes: int = 512, feature_dim: int = 768) -> None:
        super().__init__()
        self.point_dim, self.num_classes, self.feature_dim = point_dim, num_classes, feature_dim
        self.feature_extractor = PointNet(point_dim=point_dim, feature_dim=feature_dim)
        self.conv1, self.conv2, self.conv3 = _pointwise(feature_dim, 512), _pointwise(512, 256), _pointwise(256, num_classes)

    def forward_torch(self, points: torch.Tensor) -> torch.Tensor:
        x, _ = self.feature_extractor(points)
        x = x.amax(dim=2, keepdim=True)
        x = torch.relu(self.conv2(torch.relu(self.conv1(x))))
        return torch.softmax(self.conv3(x), dim=1)


class PointNetSegmentation(_B200PointMixin, nn.Module):
    _kind = 3

    def __init__(self, point_dim: int = 3, num_classes: int = 16, feature_dim: int = 768) -> None:
        super().__init__()
        self.point_dim, self.num_classes, self.feature_dim = point_dim, num_classes, feature_dim
        self.feature_extractor = PointNet(point_dim=point_dim, feature_dim=feature_dim)
        self.conv1, self.conv2 = _pointwise(feature_dim + 64, 512), _pointwise(512, 256)
        self.conv3, self.conv4 = _pointwise(256, 128), _pointwise(128, num_classes + 1)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(512), nn.BatchNorm1d(256), nn.BatchNorm1d(128)

    def forward_torch(self, points: torch.Tensor) -> torch.Tensor:
        x, x_t2 = self.feature_extractor(points)
        g = x.amax(dim=2, keepdim=True).expand(-1, -1, x_t2.shape[2])
        x = torch.cat((x_t2, g), dim=1)
        x = torch.relu(self.bn1(self.conv1(x)))
        x = torch.relu(self.bn2(self.conv2(x)))
        x = torch.relu(self.bn3(self.conv3(x)))
        return torch.nn.functional.log_softmax(self.conv4(x), dim=1).transpose(1, 2)
