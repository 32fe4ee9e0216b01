"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md>      per-kernel launch counts / device time / share
  python tools/summarize_ncu.py kernel <report.ncu-rep> <kernel> <batch> <out.md>   key raw metrics + profiles/traffic.json
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(path)})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command; times are cold-cache and\n"
                "serialised (compare shares, not absolutes).\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal device time in the list: {tot / 1e3:.2f} ms\n")
    print("wrote", out)


def raw_page(rep):
    """rows of the raw page: from a .ncu-rep, or from the CSV `ncu -i <rep> --page raw --csv` wrote on the GPU box"""
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    return list(csv.reader(lines[start:]))


def kernel(rep, name, batch, out):
    rows = raw_page(rep)
    H, U = rows[0], rows[1]
    kn = H.index("Kernel Name")
    base = lambda r: r[kn].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()   # noqa: E731
    V = next((r for r in rows[2:] if len(r) > kn and base(r) == name), rows[2])       # the report may hold several kernels
    keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_tensor.sum", "launch__shared_mem_per_block_dynamic",
            "launch__shared_mem_per_block_static", "sm__cycles_active.avg"]
    vals = {h: (v, u) for h, u, v in zip(H, U, V)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full: `{name}` (batch {batch})\n\nfrom `{os.path.basename(rep)}` (one launch, `--clock-control none`).\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in keep:
            if k in vals:
                f.write(f"| {k} | {vals[k][0]} | {vals[k][1]} |\n")
    def to_bytes(v, u):
        x = float(v.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t[name] = {"batch": int(batch), "dram_bytes_per_launch": traffic, "source": os.path.basename(out)}
    json.dump(t, open(tpath, "w"), indent=1)
    print("wrote", out, "traffic", traffic)


def _main():
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "kernel":
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])


def table(rep, out, title):
    """One row per captured launch of a multi-kernel report."""
    rows = raw_page(rep)
    H = rows[0]
    ix = {h: i for i, h in enumerate(H)}
    cols = [("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
            ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("smsp__inst_executed.sum", "warp instr")]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nfrom `{os.path.basename(rep)}` (`ncu --set full --clock-control none`, one row per captured launch; units as ncu "
                "reports them, row 2 of the raw page).\n\n")
        f.write("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        units = rows[1]
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
            vals = []
            for k, _ in cols:
                if k in ix:
                    v, u = r[ix[k]], units[ix[k]]
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    vals.append(f"{v} {u}".strip() if u not in ("%", "") else v)
                else:
                    vals.append("-")
            f.write(f"| `{name}` | " + " | ".join(vals) + " |\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "table":
        table(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        _main()
