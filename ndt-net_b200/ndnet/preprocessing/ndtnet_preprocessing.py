"""Drop-in for the reference's ndnet/preprocessing/ndtnet_preprocessing.py:6-73: same function name,
arguments and return values, but the per-cloud Python loop (GPU->CPU->C->CPU->GPU per cloud) is one
batched call into libndnet_b200.so on the current CUDA stream."""
from __future__ import annotations

import os
from typing import Tuple

import torch

from ndnet_b200 import _lib
from ndnet_b200.engine import default_engine


def ndt_preprocessing(num_nds: int, points: torch.Tensor, classes: torch.Tensor = None, num_classes: int = None,
                      textbook_kl: bool = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """points: (batch, num_points, 3) float tensor; classes: (batch, num_points, num_classes+1) one-hot or None.
    textbook_kl (not in the reference's signature; default: environment NDNET_B200_TEXTBOOK_KL=1, else off) selects the
    algorithm the reference's README documents instead of the behaviour of its compiled core.

    Returns (points_new [B,num_nds,3], covs_new [B,num_nds,9], classes_new [B,num_nds,num_classes+1] or None),
    float32 on points.device, NaN/inf replaced by 0 (ndtnet_preprocessing.py:66-69)."""
    if points.dim() != 3 or points.shape[2] != 3:
        raise ValueError("points must be shaped (batch, num_points, 3)")
    out_device = points.device
    if points.is_cuda:
        dev = points.device
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("ndt_preprocessing needs a CUDA device: ndnet_b200 has no CPU path")
        dev = torch.device("cuda", torch.cuda.current_device())
    pts = points.to(device=dev)
    if pts.dtype not in (torch.float32, torch.float64):
        pts = pts.float()            # the reference widens whatever it gets to float64 (:30)
    labels = None
    ncls = 0
    if classes is not None:
        ncls = int(num_classes)
        # one-hot -> tag (:34, numpy argmax: the first maximal index), by the library's own kernel
        onehot = classes.to(device=dev, dtype=torch.float32).contiguous()
        if onehot.dim() != 3 or onehot.shape[:2] != pts.shape[:2]:
            raise ValueError("classes must be shaped (batch, num_points, num_classes + 1)")
        labels = torch.empty(onehot.shape[:2], dtype=torch.int16, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().ndnet_b200_onehot_to_labels(onehot.data_ptr(), onehot.shape[0] * onehot.shape[1], onehot.shape[2],
                                                       labels.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_onehot_to_labels failed ({rc})")
    eng = default_engine(dev)
    if textbook_kl is None:
        textbook_kl = os.environ.get("NDNET_B200_TEXTBOOK_KL", "") == "1"
    out = eng.downsample(pts, int(num_nds), labels, ncls, nan_to_num=True, want_info=False, textbook_kl=bool(textbook_kl))
    points_new = out.feat[:, :, 0:3].contiguous()
    covs_new = out.feat[:, :, 3:12].contiguous()
    classes_new = None
    if classes is not None:
        classes_new = torch.empty((pts.shape[0], int(num_nds), ncls + 1), dtype=torch.float32, device=dev)   # (:55-57)
        with torch.cuda.device(dev):
            rc = _lib.lib().ndnet_b200_labels_to_onehot(out.labels.data_ptr(), pts.shape[0] * int(num_nds), ncls + 1,
                                                       classes_new.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"ndnet_b200_labels_to_onehot failed ({rc})")
        classes_new = classes_new.to(out_device)
    return points_new.to(out_device), covs_new.to(out_device), classes_new
