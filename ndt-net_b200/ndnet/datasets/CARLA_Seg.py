"""Drop-in for the reference's `ndnet.datasets.CARLA_Seg` (/root/reference/ndnet/datasets/CARLA_Seg.py): same class,
constructor, `__len__`/`__getitem__`, colour helpers and `get_data_pcl(pcl_filename, num_header_lines=10)` contract,
with the per-line Python parse (:115-136), the gather (:141-147) and the one-hot loop (:176-179) replaced by the GPU
reader of ndnet_b200.ply.  The random subsample is still drawn on the host by `np.random.choice(N, n_samples,
replace=False)` (:141) so that a seeded run picks the same points as the reference.

`out_device=None` (default) returns CPU tensors like the reference (usable from existing DataLoader loops);
`out_device="cuda"` keeps the sample on the GPU for `ndt_preprocessing`.
"""
import os
from typing import List, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

from ndnet_b200.ply import read_ply


class CARLA_Seg(Dataset):
    def __init__(self, n_classes: int, n_samples: int, path: str, out_device=None, device: int = 0) -> None:
        super().__init__()
        self.n_classes: int = n_classes
        self.n_samples = n_samples
        self.path: str = path
        self.out_device = out_device
        self.device = device
        if not os.path.exists(self.path):                                   # CARLA_Seg.py:25-26
            raise FileNotFoundError(f"Dataset not found at {self.path}")
        self.filenames: List[str] = os.listdir(self.path)                   # :29-31
        self.filenames.sort()

    def __len__(self) -> int:
        return len(self.filenames)

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if idx < 0 or idx >= len(self.filenames):                           # :47-48
            raise IndexError(f"Index {idx} out of bounds")
        return self.get_data_pcl(os.path.join(self.path, self.filenames[idx]))

    def color_to_class(self, color: np.ndarray) -> int:                     # :59-76
        color = (color * 255).astype(np.uint8)
        return color[0] << 16 | color[1] << 8 | color[2]

    def class_to_color(self, class_tag: int) -> np.ndarray:                 # :78-94
        r, g, b = (class_tag >> 16) & 0xff, (class_tag >> 8) & 0xff, class_tag & 0xff
        return np.array([r, g, b], dtype=np.float32) / 255.0

    def get_data_pcl(self, pcl_filename: str, num_header_lines: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
        cloud = read_ply(pcl_filename, self.n_classes, num_header_lines, self.device)
        try:
            point_indexes = np.random.choice(cloud.num_points, self.n_samples, replace=False)   # :141
            points, gt, _ = cloud.sample(point_indexes)
        finally:
            cloud.close()
        if self.out_device is None:
            return points.cpu(), gt.cpu()
        return points.to(self.out_device), gt.to(self.out_device)
