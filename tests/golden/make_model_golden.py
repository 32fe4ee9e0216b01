"""Generates tests/golden/model_ref_golden.npz by importing the REFERENCE's torch modules
(/root/reference/ndnet/models/ndtnet.py, with ndnet.preprocessing.ndt_legacy stubbed because it loads
/usr/local/lib/libndnet.so at import, ndtnet.py:5) in this container, loading the name-keyed deterministic
weights of ndnet_b200.model.deterministic_state_dict into them, and recording their eval-mode outputs on
seeded inputs.  Run here (needs /root/reference):   python tests/golden/make_model_golden.py"""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.path.join(ROOT, "ndt-net_b200") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
from ndnet_b200.model import deterministic_state_dict  # noqa: E402


def load_reference_models():
    saved = {k: v for k, v in sys.modules.items() if k == "ndnet" or k.startswith("ndnet.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    stub = types.ModuleType("ndnet.preprocessing.ndt_legacy")
    stub.NDT_Sampler = object
    try:
        import ndnet  # noqa: F401  (the reference package)
        sys.modules["ndnet.preprocessing.ndt_legacy"] = stub
        ref = importlib.import_module("ndnet.models.ndtnet")
        ref.pointnet = importlib.import_module("ndnet.models.pointnet")
        assert ref.__file__.startswith("/root/reference"), ref.__file__
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "ndnet" or k.startswith("ndnet.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref


def inputs(seed, B, N):
    rng = np.random.default_rng(seed)
    pts = rng.normal(0, 10, (B, N, 3)).astype(np.float32)
    cov = (rng.normal(0, 1, (B, N, 9)) * rng.choice([0.0, 0.1, 1.0, 30.0], (B, N, 1))).astype(np.float32)
    return pts, cov


def main():
    ref = load_reference_models()
    out = {}
    with torch.no_grad():
        seg = ref.NDTNetSegmentation(num_classes=28, feature_dim=1024)
        seg.load_state_dict(deterministic_state_dict(seg, 0)); seg.eval()
        p, c = inputs(1, 2, 200)
        out["seg_out"] = seg(torch.from_numpy(p), torch.from_numpy(c)).numpy()
        cls = ref.NDTNetClassification()
        cls.load_state_dict(deterministic_state_dict(cls, 1)); cls.eval()
        p, c = inputs(2, 3, 130)
        out["cls_out"] = cls(torch.from_numpy(p), torch.from_numpy(c)).numpy()
        # PointNet (ndnet/models/pointnet.py) on 12-D points and on plain xyz
        pseg = ref.pointnet.PointNetSegmentation(point_dim=12, num_classes=28, feature_dim=768)
        pseg.load_state_dict(deterministic_state_dict(pseg, 2)); pseg.eval()
        p, c = inputs(3, 2, 150)
        out["pn_seg_out"] = pseg(torch.from_numpy(np.concatenate([p, c], 2) * 0.3)).numpy()
        pcls = ref.pointnet.PointNetClassification(point_dim=3, num_classes=40, feature_dim=768)
        pcls.load_state_dict(deterministic_state_dict(pcls, 3)); pcls.eval()
        p, c = inputs(4, 3, 140)
        out["pn_cls_out"] = pcls(torch.from_numpy(p * 0.05)).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
