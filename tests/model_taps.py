"""Test infrastructure: the activations of the torch fp32 NDT-Net forward at the points where the CUDA library exposes
taps (ndnet_b200_model_tap), named alike, so that each layer of the CUDA path can be compared on its own.
Reference lines: /root/reference/ndnet/models/ndtnet.py:33-62 (TNet), :112-164 (NDTNet), :218-243 (segmentation head)."""
import ctypes as C

import torch


def torch_taps(net, points, covs):
    """net: eval-mode NDTNetSegmentation / NDTNetClassification (the drop-in modules).  -> dict name -> fp32 tensor"""
    taps = {}
    fe = net.feature_extractor
    B, N, d = points.shape

    def tnet(t, x, name):
        for conv, bn in ((t.conv1, t.bn1), (t.conv2, t.bn2), (t.conv3, t.bn3)):
            x = torch.relu(bn(conv(x)))
        x = x.amax(dim=2)
        taps[name + ".pool"] = x
        x = torch.relu(t.bn4(t.fc1(x)))
        x = torch.relu(t.bn5(t.fc2(x)))
        x = t.fc3(x) + torch.eye(t.in_dim, device=x.device, dtype=x.dtype).reshape(1, -1)
        taps[name] = x.reshape(-1, t.in_dim, t.in_dim)
        return taps[name]

    with torch.no_grad():
        t1 = tnet(fe.t1, points.transpose(1, 2), "t1")
        p = torch.bmm(points, t1.transpose(1, 2))
        cov = torch.matmul(t1.unsqueeze(1), covs.reshape(B, N, d, d))
        x = torch.cat((p, cov.reshape(B, N, d * d)), dim=2).transpose(1, 2)
        x = fe.bn1(fe.conv1(x))
        taps["trunk.l1"] = x.transpose(1, 2)
        t2 = tnet(fe.t2, x, "t2")
        x_t2 = torch.bmm(x.transpose(1, 2), t2).transpose(1, 2)
        taps["trunk.xt2"] = x_t2.transpose(1, 2)
        x = fe.bn2(fe.conv2(x_t2))
        taps["trunk.l2"] = x.transpose(1, 2)
        x = fe.bn3(fe.conv3(x))
        taps["trunk.pool"] = x.amax(dim=2)
        if hasattr(net, "conv4"):
            g = x.amax(dim=2, keepdim=True).expand(-1, -1, N)
            h = torch.cat((x_t2, g), dim=1)
            h = torch.relu(net.bn1(net.conv1(h))); taps["head.l1"] = h.transpose(1, 2)
            h = torch.relu(net.bn2(net.conv2(h))); taps["head.l2"] = h.transpose(1, 2)
            h = torch.relu(net.bn3(net.conv3(h))); taps["head.l3"] = h.transpose(1, 2)
            taps["out"] = torch.nn.functional.log_softmax(net.conv4(h), dim=1).transpose(1, 2)
    return {k: v.contiguous().float() for k, v in taps.items()}


def library_tap(model, name, like):
    """One tap of the LAST forward of `model` (ndnet_b200.model.B200Model) as a tensor shaped like `like`."""
    from ndnet_b200 import _lib
    L = _lib.lib()
    out = torch.empty(like.numel(), dtype=torch.float32, device=like.device)
    n = L.ndnet_b200_model_tap(model.engine.handle, model._h, name.encode(), out.data_ptr(), out.numel(),
                               torch.cuda.current_stream(like.device).cuda_stream)
    if n == -307:
        return None            # not materialised by this configuration (the fused head keeps "head.l1" on chip)
    assert n == like.numel(), (name, n, like.numel())
    return out.reshape(like.shape)
