#!/bin/bash
# One GPU-box call that refreshes the evidence for the current tree: GPU tests, smoke, the default bench line, the launch list
# of the bench command and two `ncu --set full` captures (NDT kernels, network kernels) at 512 scans per launch.
# usage: tools/gpu_round_capture.sh <tag>      (outputs under gpurun_out/, every step under its own timeout)
tag=${1:-x}
o=gpurun_out
mkdir -p $o
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $o/r2_gputest_$tag.log 2>&1
timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2_smoke_$tag.log 2>&1
timeout 400 python bench.py > $o/r2_bench_$tag.json 2> $o/r2_bench_$tag.err
NDT='regex:k_(limits|count|rank|tile_prefix|offsets|scatter|stats|kl|select)'
NET='regex:k_(gemm|fc|head12|tnet|trunk|softmax)'
timeout 500 ncu --set full --clock-control none --import-source on -k "$NDT" --launch-skip 24 -c 24 -f -o $o/prof_ndt_r2_$tag \
    python bench.py --profile-stage --batch 512 > $o/ncu_ndt_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k "$NET" --launch-skip 26 -c 26 -f -o $o/prof_net_r2_$tag \
    python bench.py --profile-stage --batch 512 > $o/ncu_net_$tag.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $o/r2_launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --batch 512 --device-chunk 128 --no-train --no-cpu-baseline > $o/ncu_launches_$tag.log 2>&1
tail -3 $o/r2_gputest_$tag.log
