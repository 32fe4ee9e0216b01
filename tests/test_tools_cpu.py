"""CPU tests of the measurement helpers: the kernel-timeline summary (tools/timeline.py), the ncu-traffic lookup of bench.py
and the committed captures it reads (profiles/traffic.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_timeline_summary_counts_overlap(tmp_path):
    import timeline
    ev = [("void ndt::k_limits<float>(a)", 7, 0.0, 10.0),          # front kernel on stream 7
          ("void ndt::k_stats<float, false>(b)", 8, 5.0, 20.0),    # back kernel on stream 8, overlaps the first for 5 us
          ("mlp::k_gemm_chmax(c)", 8, 30.0, 5.0)]                  # alone, after a 5 us gap
    out = tmp_path / "t.md"
    timeline.summarise(ev, str(out), "t")
    text = out.read_text()
    assert "span 0.035 ms, 3 kernels on 2 streams" in text
    rows = {ln.split("|")[1].strip(): [c.strip() for c in ln.split("|")[2:-1]] for ln in text.splitlines() if ln.startswith("| ")}
    assert rows["0"] == ["14.3 %"] and rows["1"] == ["71.4 %"] and rows["2"] == ["14.3 %"]
    assert rows["front + back"] == ["14.3 %"] and rows["front only"] == ["14.3 %"] and rows["idle"] == ["14.3 %"]
    assert rows["`k_limits<float>`"][-1] == "50 %" and rows["`k_stats<float, false>`"][-1] == "25 %"
    assert rows["`k_gemm_chmax`"][-1] == "0 %"


def test_bench_traffic_lookup_matches_the_committed_captures():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for k in ("k_limits", "k_count", "k_rank", "k_tile_prefix", "k_offsets", "k_scatter", "k_stats", "k_stats_light", "k_kl", "k_select"):
        assert t[k]["batch"] == 512 and t[k]["dram_bytes_per_launch"] > 0, k
        assert os.path.exists(os.path.join(ROOT, t[k]["raw"])), k          # the raw ncu page the number comes from is committed
    assert bench.ncu_traffic("k_stats", 512) == t["k_stats"]["dram_bytes_per_launch"]
    assert bench.ncu_traffic("k_stats+k_stats_light", 512) == t["k_stats"]["dram_bytes_per_launch"] + t["k_stats_light"]["dram_bytes_per_launch"]
    assert bench.ncu_traffic("k_stats", 2048) is None and bench.ncu_traffic("k_nope", 512) is None
    # the statistics stage moves about 1.2 x the algorithmic bytes of the launch, not multiples of it
    algo = bench.ALGO_BYTES_PER_CLOUD * 512
    assert 0.9 * algo < bench.ncu_traffic("k_stats+k_stats_light", 512) < 1.5 * algo
