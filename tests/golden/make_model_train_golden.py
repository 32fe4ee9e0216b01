"""Generates tests/golden/model_ref_train_golden.npz: the REFERENCE's torch modules (/root/reference/ndnet/models/ndtnet.py,
pointnet.py, imported as in make_model_golden.py) in TRAINING mode - BatchNorm on batch statistics - on seeded inputs with the
name-keyed deterministic weights: the forward output, the loss of /root/reference/tools/train.py:74 (cross_entropy against
seeded one-hot / class targets), the gradient of every parameter (sum and sum of absolute values; a few whole tensors) and the
BatchNorm running statistics after the step.  The GPU training tests compare the library with the drop-in modules'
forward_torch in train mode; this file pins that torch definition to the reference's own modules in the same mode.
Run here (needs /root/reference):   python tests/golden/make_model_train_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "ndt-net_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from ndnet_b200.model import deterministic_state_dict  # noqa: E402
from tests.golden.make_model_golden import inputs, load_reference_models  # noqa: E402

FULL_GRADS = ("feature_extractor.t1.fc3.weight", "feature_extractor.conv1.weight", "feature_extractor.t2.fc3.bias")


def cases(mods):
    """(name, module, call arguments, target) for the four networks; `mods` has the four classes."""
    out = []
    p, c = inputs(11, 3, 96)
    rng = np.random.default_rng(5)
    seg = mods["NDTNetSegmentation"](num_classes=28, feature_dim=256)
    tgt = torch.zeros((3, 96, 29)).scatter_(2, torch.from_numpy(rng.integers(0, 29, (3, 96, 1))), 1.0)
    out.append(("seg", seg, (torch.from_numpy(p * 0.1), torch.from_numpy(c * 0.1)), tgt, 10))
    p, c = inputs(12, 4, 80)
    cls = mods["NDTNetClassification"]()
    out.append(("cls", cls, (torch.from_numpy(p * 0.1), torch.from_numpy(c * 0.1)), torch.from_numpy(rng.integers(0, 512, (4, 1))), 11))
    p, c = inputs(13, 3, 72)
    pseg = mods["PointNetSegmentation"](point_dim=12, num_classes=28, feature_dim=256)
    tgt = torch.zeros((3, 72, 29)).scatter_(2, torch.from_numpy(rng.integers(0, 29, (3, 72, 1))), 1.0)
    out.append(("pn_seg", pseg, (torch.from_numpy(np.concatenate([p, c], 2) * 0.1),), tgt, 12))
    p, c = inputs(14, 4, 64)
    pcls = mods["PointNetClassification"](point_dim=3, num_classes=40, feature_dim=256)
    out.append(("pn_cls", pcls, (torch.from_numpy(p * 0.05),), torch.from_numpy(rng.integers(0, 40, (4, 1))), 13))
    return out


def run_case(net, args, target, seed, forward=None):
    """one training-mode pass: output, loss, parameter gradients, BatchNorm buffers afterwards"""
    net.load_state_dict(deterministic_state_dict(net, seed))
    net.train()
    out = (forward or net)(*args)
    loss = torch.nn.functional.cross_entropy(out, target)              # tools/train.py:74
    net.zero_grad()
    loss.backward()
    res = {"out": out.detach().numpy(), "loss": np.float64(loss.item())}
    names, sums, asums = [], [], []
    for k, prm in net.named_parameters():
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        names.append(k); sums.append(g.double().sum().item()); asums.append(g.double().abs().sum().item())
        if k in FULL_GRADS:
            res["grad." + k] = g.numpy().copy()
    res["grad_names"] = np.array(names)
    res["grad_sum"] = np.array(sums)
    res["grad_abs_sum"] = np.array(asums)
    bn = {k: v.numpy().copy() for k, v in net.state_dict().items() if k.endswith("running_mean") or k.endswith("running_var")}
    keys = sorted(bn)
    res["bn_names"] = np.array(keys)
    res["bn_sum"] = np.array([bn[k].astype(np.float64).sum() for k in keys])
    return res


def main():
    ref = load_reference_models()
    mods = {"NDTNetSegmentation": ref.NDTNetSegmentation, "NDTNetClassification": ref.NDTNetClassification,
            "PointNetSegmentation": ref.pointnet.PointNetSegmentation, "PointNetClassification": ref.pointnet.PointNetClassification}
    out = {}
    torch.manual_seed(0)
    for name, net, args, target, seed in cases(mods):
        for k, v in run_case(net, args, target, seed).items():
            out[f"{name}.{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_ref_train_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: float(out[k]) for k in out if k.endswith(".loss")})


if __name__ == "__main__":
    main()
