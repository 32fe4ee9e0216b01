"""Developer probe: CUDA forward vs torch fp32, with error statistics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
import numpy as np, torch
from ndnet.models.ndtnet import NDTNetClassification, NDTNetSegmentation
from ndnet_b200.model import deterministic_state_dict
from tests.golden.make_model_golden import inputs

net = NDTNetSegmentation(num_classes=28, feature_dim=1024); net.load_state_dict(deterministic_state_dict(net, 0)); net = net.cuda().eval()
for (B, N, sc) in [(2, 200, 1.0), (2, 200, 0.05), (1, 1000, 0.05), (64, 1000, 0.05)]:
    p, c = inputs(3, B, N); p, c = torch.from_numpy(p).cuda() * sc, torch.from_numpy(c).cuda() * sc
    with torch.no_grad():
        ref = net.forward_torch(p, c); got = net.forward_b200(p, c)
        torch.cuda.synchronize()
        t = time.time(); 
        for _ in range(5): got = net.forward_b200(p, c)
        torch.cuda.synchronize(); dt = (time.time() - t) / 5
        t = time.time()
        for _ in range(5): ref = net.forward_torch(p, c)
        torch.cuda.synchronize(); dt2 = (time.time() - t) / 5
    d = (got - ref).abs()
    print("   ref absmax", ref.abs().max().item(), "rel err (max/absmax)", (d.max() / ref.abs().max()).item())
    print("seg", B, N, "finite", bool(torch.isfinite(got).all()), "max", d.max().item(), "mean", d.mean().item(),
          "argmax agree", (got.argmax(-1) == ref.argmax(-1)).float().mean().item(), "ours %.3f ms torch %.3f ms" % (dt * 1e3, dt2 * 1e3))
cls = NDTNetClassification(); cls.load_state_dict(deterministic_state_dict(cls, 1)); cls = cls.cuda().eval()
for (B, N) in [(3, 130), (32, 512)]:
    p, c = inputs(4, B, N); p, c = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda()
    with torch.no_grad():
        ref = cls.forward_torch(p, c); got = cls.forward_b200(p, c)
    d = (got - ref).abs()
    print("cls", B, N, "max", d.max().item(), "ref max prob", ref.max().item(), "argmax agree", (got.argmax(1) == ref.argmax(1)).float().mean().item())
