#!/usr/bin/env python
"""Device-resident throughput (bench.py's `value` leg) over pipeline configurations, one process, inputs generated once:

    python tools/sweep_pipeline.py --configs 512:0,512:1,256:0,256:1 [--batch 2048] [--steps 6] [--lanes 8]

each configuration is `device_chunk:stagger`; prints one JSON line per configuration (clouds/s, ms per step).  Timing rules as
in bench.py: three warm-up steps, CUDA events on the launching stream, inputs rotate over two resident batches (> L2).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="512:0,512:1,256:0,256:1,128:0,128:1")
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--host", action="store_true", help="also time the host-buffer path (ndnet_b200_infer_host_u8)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from ndnet_b200.model import B200Model, KIND_SEG

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B = args.batch
    sets, host = [], []
    for s in range(2):
        p, l = bench.make_scans(B, seed0=bench.ndist_seed(0, B, s))
        sets.append((torch.from_numpy(p).to(dev), torch.from_numpy(l.astype(np.int16)).to(dev)))
        if args.host:
            host.append((torch.from_numpy(p).pin_memory(), torch.from_numpy(l.astype(np.uint8)).pin_memory()))
    model = B200Model(bench.build_network(dev), KIND_SEG, dev)
    out_host = torch.empty((B, bench.N_NDS, bench.N_CLASSES + 1), dtype=torch.float32).pin_memory() if args.host else None
    stream = torch.cuda.current_stream(dev)
    for cfg in args.configs.split(","):
        chunk, stagger = (int(x) for x in cfg.split(":"))
        model.set_pipeline(args.lanes, 64, chunk, stagger=bool(stagger))

        def step(i):
            return model.infer_device(sets[i % 2][0], bench.N_NDS, sets[i % 2][1], bench.N_CLASSES)

        def step_host(i):
            return model.infer_host(host[i % 2][0], bench.N_NDS, host[i % 2][1], bench.N_CLASSES, out_host)

        res = {"device_chunk": chunk, "stagger": stagger, "batch": B, "lanes": args.lanes}
        for name, fn in (("device", step),) + ((("host", step_host),) if args.host else ()):
            for i in range(3):
                fn(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record(stream)
            for i in range(args.steps):
                fn(i)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / args.steps
            res[name] = {"clouds_per_s": B / (ms * 1e-3), "ms_per_step": ms}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
