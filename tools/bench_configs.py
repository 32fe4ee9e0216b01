"""Timings of the BASELINE.json configurations that bench.py does not quote (they are parity-test cases; this is a
courtesy measurement):  config 2 (ModelNet-shaped 2048-point clouds, batch 32, n_desired_nds=512, classification
forward) and config 5 (multiscale NDT at 4096/1024/256 per cloud, 120k-point scans).  One GPU, CUDA events, inputs
resident in HBM and rotated over sets larger than L2.  Prints one JSON line per config.
    python tools/bench_configs.py [--steps 10]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup=3):
    for i in range(warmup):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--scans", type=int, default=64, help="config 5: scans per step")
    args = ap.parse_args()
    from ndnet.models.ndtnet import NDTNetClassification
    from ndnet_b200.engine import NdtEngine
    from ndnet_b200.model import KIND_CLS, B200Model, deterministic_state_dict
    from ndnet_b200.synth import lidar_batch, modelnet_cloud
    dev = torch.device("cuda", 0)
    eng = NdtEngine(0)

    # ---- config 2: batch 32 x 2048 points -> 512 distributions -> NDTNetClassification() forward
    B, N, D = 32, 2048, 512
    n_sets = 160                                                           # 160 x 32 x 24 KB = 126 MB of points
    sets = [torch.from_numpy(np.stack([modelnet_cloud(N, s * B + b) for b in range(B)])).to(dev) for s in range(n_sets)]
    net = NDTNetClassification()
    net.load_state_dict(deterministic_state_dict(net, 1))
    model = B200Model(net.to(dev).eval(), KIND_CLS, dev)
    chk = eng.downsample(sets[0], D)
    ok = int((chk.info["status"] == 0).sum())

    def step2(i):
        f = eng.downsample(sets[i % n_sets], D, nan_to_num=True, want_info=False).feat
        return model(f)

    ms = timed(step2, max(args.steps, 50))
    print(json.dumps({"config": "config2: ModelNet-shaped 2048-point clouds, batch 32, n_desired_nds=512, NDTNetClassification forward",
                      "value": B / (ms * 1e-3), "unit": "clouds/s", "ms_per_step": ms, "clouds_per_step": B, "converged": f"{ok}/{B}",
                      "l2": f"inputs rotate over {n_sets} resident batches"}))

    # ---- config 5: multiscale NDT, 4096 / 1024 / 256 distributions per 120k-point scan
    S = args.scans
    pts = [torch.from_numpy(lidar_batch(S, 120_000, seed0=1000 * k)).to(dev) for k in range(2)]
    res = eng.downsample_multiscale(pts[0], (4096, 1024, 256))
    conv = [int((r.info["status"] == 0).sum()) for r in res]

    def step5(i):
        return eng.downsample_multiscale(pts[i % 2], (4096, 1024, 256), nan_to_num=True, want_info=False)

    ms = timed(step5, args.steps)
    print(json.dumps({"config": "config5: multiscale NDT at n_desired_nds 4096/1024/256 per cloud, 120k-point scans",
                      "value": S / (ms * 1e-3), "unit": "clouds/s (all three resolutions)", "ms_per_step": ms, "clouds_per_step": S,
                      "converged": {"4096": f"{conv[0]}/{S}", "1024": f"{conv[1]}/{S}", "256": f"{conv[2]}/{S}"},
                      "l2": "inputs rotate over 2 resident batches (2 x 92 MB)"}))


if __name__ == "__main__":
    main()
