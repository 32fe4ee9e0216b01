#!/usr/bin/env python
"""Kernel timeline of the pipelined device-resident step (bench.py's `value` leg), from CUPTI through torch.profiler (nsys is
not in the image).  Shows which kernels of which lane run side by side: per kernel the share of its run time during which a
kernel of ANOTHER stream was running, the time-weighted number of kernels in flight, and front (limits / search / voxel
assignment) against back (statistics / divergences / selection / network) concurrency.

    python tools/timeline.py [--batch 2048] [--device-chunk 512] [--lanes 8] [--stagger 0|1] --out gpurun_out/r2_timeline

writes <out>.md (summary) and <out>.json (every kernel interval: name, stream, start us, duration us).
A run under the profiler is not a benchmark: the times here only say what overlaps what.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))

FRONT = ("k_init_limits", "k_limits", "k_decide", "k_count", "k_rank", "k_tile_prefix", "k_offsets", "k_scatter")


def short(name: str) -> str:
    n = name.split("(")[0]
    for p in ("void ", "ndt::", "mlp::", "(anonymous namespace)::"):
        n = n.replace(p, "")
    return n.strip()


def summarise(ev, out_md, title):
    """ev: list of (name, stream, ts_us, dur_us)"""
    ev = sorted(ev, key=lambda e: e[2])
    t0 = ev[0][2]
    t1 = max(e[2] + e[3] for e in ev)
    # sweep: points where the set of running kernels changes
    pts = []
    for i, (n, s, ts, d) in enumerate(ev):
        pts.append((ts, 1, i)); pts.append((ts + d, 0, i))
    pts.sort()
    running = set()
    conc_time = {}
    other_stream = [0.0] * len(ev)       # per kernel: time with a kernel of another stream in flight
    front_back = {"front only": 0.0, "back only": 0.0, "front + back": 0.0, "idle": 0.0}
    last = t0
    for t, kind, i in pts:
        dt = t - last
        if dt > 0:
            k = len(running)
            conc_time[k] = conc_time.get(k, 0.0) + dt
            streams = {}
            for j in running:
                streams[ev[j][1]] = streams.get(ev[j][1], 0) + 1
            for j in running:
                if len(streams) > 1:
                    other_stream[j] += dt
            f = any(short(ev[j][0]).startswith(FRONT) for j in running)
            b = any(not short(ev[j][0]).startswith(FRONT) for j in running)
            front_back["front + back" if f and b else "front only" if f else "back only" if b else "idle"] += dt
        if kind == 1:
            running.add(i)
        else:
            running.discard(i)
        last = t
    span = t1 - t0
    per = {}
    for i, (n, s, ts, d) in enumerate(ev):
        a = per.setdefault(short(n), [0, 0.0, 0.0])
        a[0] += 1; a[1] += d; a[2] += other_stream[i]
    lines = [f"# {title}", "",
             f"span {span / 1e3:.3f} ms, {len(ev)} kernels on {len({e[1] for e in ev})} streams, sum of kernel durations "
             f"{sum(e[3] for e in ev) / 1e3:.3f} ms (= {sum(e[3] for e in ev) / span:.2f} kernels in flight on average)", "",
             "| kernels in flight | share of the span |", "|---:|---:|"]
    for k in sorted(conc_time):
        lines.append(f"| {k} | {100 * conc_time[k] / span:.1f} % |")
    lines += ["", "| front = limits, search, voxel assignment; back = statistics, divergences, selection, network | share of the span |", "|---|---:|"]
    for k, v in front_back.items():
        lines.append(f"| {k} | {100 * v / span:.1f} % |")
    lines += ["", "| kernel | launches | total ms | share of its run time beside a kernel of another stream |", "|---|---:|---:|---:|"]
    for n, (c, d, o) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n}` | {c} | {d / 1e3:.3f} | {100 * o / max(d, 1e-9):.0f} % |")
    with open(out_md, "w") as f:
        f.write("\n".join(lines) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--device-chunk", type=int, default=512)
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--stagger", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_timeline"))
    args = ap.parse_args()
    if args.stagger >= 0:
        os.environ["NDNET_B200_STAGGER"] = str(args.stagger)
    import numpy as np
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    from ndnet_b200.model import B200Model, KIND_SEG

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B = args.batch
    sets = []
    for s in range(2):
        p, l = bench.make_scans(B, seed0=bench.ndist_seed(0, B, s))
        sets.append((torch.from_numpy(p).to(dev), torch.from_numpy(l.astype(np.int16)).to(dev)))
    model = B200Model(bench.build_network(dev), KIND_SEG, dev)
    model.set_pipeline(args.lanes, 64, args.device_chunk)

    def step(i):
        return model.infer_device(sets[i % 2][0], bench.N_NDS, sets[i % 2][1], bench.N_CLASSES)

    for i in range(3):
        step(i)
    torch.cuda.synchronize(dev)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize(dev)
    trace = args.out + ".trace.json"
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        tr = json.load(f)
    ev = [(e["name"], e.get("args", {}).get("stream", e.get("tid")), float(e["ts"]), float(e["dur"]))
          for e in tr["traceEvents"] if e.get("cat") == "kernel" and "dur" in e]
    os.remove(trace)
    if not ev:
        raise SystemExit("the profiler recorded no kernels")
    title = (f"timeline of {args.steps} device-resident steps: {B} scans per step, chunks of {args.device_chunk}, {args.lanes} lanes, "
             f"stagger={os.environ.get('NDNET_B200_STAGGER', 'library default')} (CUPTI via torch.profiler; not a benchmark)")
    summarise(ev, args.out + ".md", title)
    names = sorted({short(e[0]) for e in ev})
    t0 = min(e[2] for e in ev)
    with open(args.out + ".json", "w") as f:
        json.dump({"title": title, "names": names,
                   "kernels": [[names.index(short(n)), s, round(ts - t0, 2), round(d, 2)] for n, s, ts, d in ev]}, f)
    print(open(args.out + ".md").read())


if __name__ == "__main__":
    main()
