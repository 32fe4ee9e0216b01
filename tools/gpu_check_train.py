"""Diagnostic: gradient error of the training kernels and of torch fp32 autograd, both against torch fp64 autograd,
over several seeds (the network is ill-conditioned in fp32: ReLU-mask and arg-max flips make the error a random
variable, so one draw says little).   python tools/gpu_check_train.py [B N [seeds]]"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
sys.path.insert(0, ROOT)
from tests.test_train_gpu import _batch, _net  # noqa: E402
from ndnet_b200.train import SegTrainer, reference_loss  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 200)
seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 6
verbose = len(sys.argv) > 4
for seed in range(seeds):
    net32 = _net(768, 28, seed)
    net64 = copy.deepcopy(net32).double()
    netb = copy.deepcopy(net32)
    pts, cov, gt = _batch(seed * 77 + B * 1000 + N, B, N, 28)
    o32 = net32(pts, cov)
    reference_loss(o32, gt).backward()
    o64 = net64(pts.double(), cov.double())
    reference_loss(o64, gt.double()).backward()
    out = SegTrainer(netb)(pts, cov)
    reference_loss(out, gt).backward()
    n32 = nb = nr = 0.0
    worst32 = worstb = 0.0
    for (name, p64), (_, p32), (_, pb) in zip(net64.named_parameters(), net32.named_parameters(), netb.named_parameters()):
        g = p64.grad.flatten()
        if g.norm().item() < 1e-6:
            continue
        e32 = (p32.grad.double().flatten() - g).norm().item()
        eb = (pb.grad.double().flatten() - g).norm().item()
        n32 += e32 ** 2; nb += eb ** 2; nr += g.norm().item() ** 2
        worst32 = max(worst32, e32 / g.norm().item()); worstb = max(worstb, eb / g.norm().item())
        if verbose:
            print(f"  {name:45s} |g|={g.norm().item():.3e}  torch32 {e32 / g.norm().item():.2e}  ours {eb / g.norm().item():.2e}")
    print(f"seed {seed}: fwd err torch32 {(o32.double() - o64).abs().max().item():.2e} ours {(out.double() - o64).abs().max().item():.2e} | "
          f"grad rel err (all params) torch32 {(n32 / nr) ** 0.5:.2e} ours {(nb / nr) ** 0.5:.2e} | worst param torch32 {worst32:.2e} ours {worstb:.2e}")
