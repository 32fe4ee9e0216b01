"""BASELINE config 3 — the README training shape: batch 16 synthetic 100k-point LiDAR scans, n_desired_nds=1000,
NDTNetSegmentation(num_classes=28, feature_dim=768) forward + backward, the batch sharded over the GPUs (2 scans per
GPU on 8).  One step = tools/train.py:58-80 of the reference: ndt_preprocessing -> model -> cross_entropy -> backward
-> (gradient all-reduce) -> Adam step.

    python tools/bench_train.py [--clouds-per-gpu 2] [--steps 10]            # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_train.py

Prints one JSON line: clouds/s through our kernels (NDT + train.cu), the same step with torch autograd for the network
(library path, for comparison only), and the reference on the host (oracle/_ref C core + torch CPU autograd) on a
bounded sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
sys.path.insert(0, ROOT)

N_POINTS, N_NDS, N_CLASSES, FEATURE_DIM = 100_000, 1000, 28, 768


def make_batch(n, seed0, device):
    from ndnet_b200.synth import lidar_batch
    pts, lab = lidar_batch(n, N_POINTS, seed0=seed0, with_labels=True, num_classes=N_CLASSES)
    gt = torch.zeros((n, N_POINTS, N_CLASSES + 1), dtype=torch.float32)
    gt.scatter_(2, torch.from_numpy(lab.astype(np.int64)).unsqueeze(-1), 1.0)
    return torch.from_numpy(pts).to(device), gt.to(device)


def build(device):
    from ndnet.models.ndtnet import NDTNetSegmentation
    from ndnet_b200.model import deterministic_state_dict
    net = NDTNetSegmentation(num_classes=N_CLASSES, feature_dim=FEATURE_DIM)
    net.load_state_dict(deterministic_state_dict(net, 0))
    return net.to(device).train()


def cpu_reference_step(n_clouds, threads):
    """The reference's step on the host: its C core per scan (oracle/_ref, 8 pthreads) + torch CPU autograd."""
    from oracle import ref_ctypes
    torch.set_num_threads(threads)
    lib = ref_ctypes.load(ref_ctypes.ref_lib_path("threaded"))
    net = build("cpu")
    opt = torch.optim.Adam(net.parameters(), lr=0.034)
    pts, gt = make_batch(n_clouds, 5000, "cpu")
    lab = gt.argmax(2).numpy().astype(np.uint16)
    t0 = time.perf_counter()
    means = np.zeros((n_clouds, N_NDS, 3), np.float32)
    covs = np.zeros((n_clouds, N_NDS, 9), np.float32)
    cls = np.zeros((n_clouds, N_NDS), np.int64)
    for b in range(n_clouds):
        r = ref_ctypes.downsample(lib, pts[b].numpy().astype(np.float64), N_NDS, lab[b], N_CLASSES, out_rows=N_NDS + 256)
        k = min(len(r.points), N_NDS)
        means[b, :k], covs[b, :k], cls[b, :k] = r.points[:k], r.covs[:k], r.classes[:k]
    m = torch.nan_to_num(torch.from_numpy(means), nan=0.0, posinf=0.0, neginf=0.0)
    c = torch.nan_to_num(torch.from_numpy(covs), nan=0.0, posinf=0.0, neginf=0.0)
    g = torch.zeros((n_clouds, N_NDS, N_CLASSES + 1)).scatter_(2, torch.from_numpy(cls).unsqueeze(-1), 1.0)
    loss = torch.nn.functional.cross_entropy(net(m, c), g)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds-per-gpu", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-clouds", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fp32", action="store_true", help="train.cu GEMMs on the fp32 FMA kernel instead of tcgen05 TF32")
    ap.add_argument("--only-ours", action="store_true", help="skip the torch-autograd comparison run (for ncu launch lists)")
    args = ap.parse_args()
    real_stdout = os.dup(1)          # NCCL's version banner goes to fd 1: keep stdout for the JSON line
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    import torch.distributed as dist
    from ndnet.preprocessing.ndtnet_preprocessing import ndt_preprocessing
    from ndnet_b200 import _lib
    from ndnet_b200.train import allreduce_gradients, reference_loss

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.clouds_per_gpu
    sets = [make_batch(B, 1000 * rank + 10 * s, dev) for s in range(3)]
    L = _lib.lib()

    def run(mode):
        net = build(dev)
        net.b200_tf32 = not args.fp32
        opt = torch.optim.Adam(net.parameters(), lr=0.034)              # tools/train.py:108,147

        def step(i):
            pcl, gt = sets[i % len(sets)]
            p, c, g = ndt_preprocessing(N_NDS, pcl, gt, N_CLASSES)      # :69
            pred = net(p, c) if mode == "ours" else net.forward_torch(p, c)   # :71
            loss = reference_loss(pred, g)                               # :74
            opt.zero_grad()
            loss.backward()                                              # :78
            allreduce_gradients(net)
            opt.step()                                                   # :83
            return loss

        for i in range(args.warmup):
            step(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        launches0 = L.ndnet_b200_launch_count()
        e0.record()
        for i in range(args.steps):
            loss = step(i)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / args.steps, float(loss.detach()), L.ndnet_b200_launch_count() - launches0

    ms_ours, loss_ours, ndt_launches = run("ours")
    ms_torch, loss_torch, _ = (float("nan"), float("nan"), 0) if args.only_ours else run("torch")
    if rank == 0:
        line = {"metric": "clouds/sec (NDT + NDTNetSegmentation fwd+bwd + Adam), 100k-point scans, D=1000, F=768",
                "config": {"workload": "config3: README training shape", "clouds_per_gpu_per_step": B, "n_gpus": world,
                           "global_batch": B * world, "bn": "per-replica batch statistics"},
                "value": B * world / (ms_ours * 1e-3), "unit": "clouds/s", "ms_per_step": ms_ours, "loss_last": loss_ours,
                "network_on_torch_autograd": {"value": B * world / (ms_torch * 1e-3), "ms_per_step": ms_torch, "loss_last": loss_torch,
                                              "note": "same NDT kernels, network fwd/bwd through torch (cuBLAS/cuDNN) instead of train.cu"},
                "ndt_launches_per_step": ndt_launches / args.steps, "steps": args.steps, "warmup": args.warmup,
                "dtype": "f64 (NDT) + " + ("f32" if args.fp32 else "tf32 tensor-core GEMMs, f32 accumulate/elementwise") + " (network fwd/bwd)", "data": "synthetic"}
        if not args.no_cpu and world == 1:
            threads = os.cpu_count() or 1
            cpu_reference_step(2, threads)
            dt = cpu_reference_step(args.cpu_clouds, threads)
            line["cpu_reference"] = {"value": args.cpu_clouds / dt, "unit": "clouds/s", "cores": threads,
                                     "sample": f"one step of {args.cpu_clouds} scans: reference C core (-O0, 8 pthreads) + torch CPU autograd + Adam"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
