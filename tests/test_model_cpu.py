"""CPU test: the drop-in torch modules (ndt-net_b200/ndnet/models/ndtnet.py) reproduce the outputs the
reference's own modules gave (tests/golden/model_ref_golden.npz, made by make_model_golden.py) and keep the
reference's state_dict key set."""
import os

import numpy as np
import torch

from ndnet.models.ndtnet import NDTNetClassification, NDTNetSegmentation
from ndnet_b200.model import deterministic_state_dict
from tests.golden.make_model_golden import inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "model_ref_golden.npz")


def test_drop_in_modules_match_reference_outputs():
    g = np.load(GOLDEN)
    with torch.no_grad():
        seg = NDTNetSegmentation(num_classes=28, feature_dim=1024)
        seg.load_state_dict(deterministic_state_dict(seg, 0)); seg.eval()
        p, c = inputs(1, 2, 200)
        out = seg(torch.from_numpy(p), torch.from_numpy(c)).numpy()
        assert out.shape == g["seg_out"].shape == (2, 200, 29)
        assert np.allclose(out, g["seg_out"], rtol=1e-4, atol=1e-4)
        cls = NDTNetClassification()
        cls.load_state_dict(deterministic_state_dict(cls, 1)); cls.eval()
        p, c = inputs(2, 3, 130)
        out = cls(torch.from_numpy(p), torch.from_numpy(c)).numpy()
        assert out.shape == g["cls_out"].shape == (3, 512, 1)
        assert np.allclose(out, g["cls_out"], rtol=1e-4, atol=1e-6)


def test_pointnet_drop_in_modules_match_reference_outputs():
    from ndnet.models.pointnet import PointNetClassification, PointNetSegmentation
    g = np.load(GOLDEN)
    with torch.no_grad():
        pseg = PointNetSegmentation(point_dim=12, num_classes=28, feature_dim=768)
        pseg.load_state_dict(deterministic_state_dict(pseg, 2)); pseg.eval()
        p, c = inputs(3, 2, 150)
        out = pseg(torch.from_numpy(np.concatenate([p, c], 2) * 0.3)).numpy()
        assert np.allclose(out, g["pn_seg_out"], rtol=1e-4, atol=1e-4)
        pcls = PointNetClassification(point_dim=3, num_classes=40, feature_dim=768)
        pcls.load_state_dict(deterministic_state_dict(pcls, 3)); pcls.eval()
        p, c = inputs(4, 3, 140)
        out = pcls(torch.from_numpy(p * 0.05)).numpy()
        assert np.allclose(out, g["pn_cls_out"], rtol=1e-4, atol=1e-6)


def test_state_dict_keys_follow_the_reference_layout():
    keys = set(NDTNetSegmentation(num_classes=28).state_dict())
    for k in ["feature_extractor.t1.conv1.weight", "feature_extractor.t2.fc3.bias", "feature_extractor.bn3.running_var",
              "feature_extractor.conv3.weight", "conv4.weight", "bn3.running_mean", "feature_extractor.t1.bn5.weight"]:
        assert k in keys, k
    assert NDTNetSegmentation(num_classes=28).state_dict()["conv1.weight"].shape == (512, 1088, 1)
    assert NDTNetClassification().state_dict()["conv3.weight"].shape == (512, 256, 1)


def test_drop_in_modules_match_reference_in_training_mode():
    """The GPU training tests compare the library with these modules' forward_torch in train() mode; here that definition is
    pinned to the reference's own modules in the same mode (tests/golden/model_ref_train_golden.npz, made by
    make_model_train_golden.py from /root/reference): forward on batch statistics, the loss of tools/train.py:74, every
    parameter gradient and the BatchNorm running statistics after the step, for all four networks."""
    from ndnet.models.pointnet import PointNetClassification, PointNetSegmentation
    from tests.golden.make_model_train_golden import FULL_GRADS, cases, run_case
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_ref_train_golden.npz"))
    mods = {"NDTNetSegmentation": NDTNetSegmentation, "NDTNetClassification": NDTNetClassification,
            "PointNetSegmentation": PointNetSegmentation, "PointNetClassification": PointNetClassification}
    torch.manual_seed(0)
    for name, net, args, target, seed in cases(mods):
        r = run_case(net, args, target, seed, forward=net.forward_torch)
        assert r["out"].shape == g[f"{name}.out"].shape, name
        assert np.allclose(r["out"], g[f"{name}.out"], rtol=2e-4, atol=2e-5), (name, np.abs(r["out"] - g[f"{name}.out"]).max())
        assert abs(r["loss"] - float(g[f"{name}.loss"])) <= 1e-5 * abs(float(g[f"{name}.loss"])), name
        assert list(r["grad_names"]) == list(g[f"{name}.grad_names"]), name          # same parameters, same order
        scale = np.maximum(g[f"{name}.grad_abs_sum"], 1e-12)
        # a bias in front of a train-mode BatchNorm has an analytically zero gradient: what both sides hold there is rounding
        # noise, hence the absolute term (1e-6 of the network's largest gradient sum)
        noise = 1e-6 * float(g[f"{name}.grad_abs_sum"].max())
        assert np.all(np.abs(r["grad_abs_sum"] - g[f"{name}.grad_abs_sum"]) <= 2e-3 * scale + noise), name
        assert np.all(np.abs(r["grad_sum"] - g[f"{name}.grad_sum"]) <= 2e-3 * scale + noise), name
        for k in FULL_GRADS:
            if f"{name}.grad.{k}" in g:
                ref = g[f"{name}.grad.{k}"]
                assert np.allclose(r["grad." + k], ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max() + 1e-9), (name, k)
        assert list(r["bn_names"]) == list(g[f"{name}.bn_names"]), name
        assert np.allclose(r["bn_sum"], g[f"{name}.bn_sum"], rtol=1e-4, atol=1e-5), name
