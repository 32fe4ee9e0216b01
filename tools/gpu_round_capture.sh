#!/bin/bash
# GPU-box calls that refresh the evidence for the current tree (outputs under gpurun_out/, every step under its own timeout;
# .ncu-rep files are turned into their raw-page CSV on the box and dropped: gpurun brings back at most 64 MiB).
#   tools/gpu_round_capture.sh check <tag>    GPU tests, smoke, the default bench line, training step with/without the NDT graph
#   tools/gpu_round_capture.sh ncu <tag>      `ncu --set full` of the NDT and network kernels at 512 scans per launch
#   tools/gpu_round_capture.sh launches <tag> launch list of the bench command
mode=${1:-check}
tag=${2:-x}
o=gpurun_out
mkdir -p $o
stamp() { echo "$(date +%T) $*" >> $o/progress_$tag.log; }
if [ "$mode" = check ]; then
    stamp tests
    ( time timeout 900 python -m pytest tests -m gpu -q --maxfail=8 ) > $o/r2_gputest_$tag.log 2>&1
    stamp smoke
    timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/r2_smoke_$tag.log 2>&1
    stamp bench
    timeout 400 python bench.py > $o/r2_bench_$tag.json 2> $o/r2_bench_$tag.err
    stamp train
    NDNET_B200_NDT_GRAPH=0 timeout 200 python tools/bench_train.py --no-cpu > $o/r2_train_${tag}_direct.json 2> $o/r2_train_$tag.err
    timeout 200 python tools/bench_train.py --no-cpu > $o/r2_train_${tag}_graph.json 2>> $o/r2_train_$tag.err
    timeout 200 python tools/bench_train.py --no-cpu --clouds-per-gpu 16 > $o/r2_train_${tag}_graph16.json 2>> $o/r2_train_$tag.err
    stamp done
    tail -3 $o/r2_gputest_$tag.log
elif [ "$mode" = ncu ]; then
    stamp tests
    ( time timeout 600 python -m pytest tests -m gpu -q --maxfail=8 ) > $o/r2_gputest_$tag.log 2>&1
    cap() {   # cap <name> <kernel regex> <skip> <count>
        stamp "ncu $1"
        timeout 420 ncu --set full --clock-control none -k "$2" --launch-skip $3 -c $4 -f -o $o/prof_$1_$tag \
            python bench.py --profile-stage --batch 512 > $o/ncu_$1_$tag.log 2>&1
        ncu -i $o/prof_$1_$tag.ncu-rep --page raw --csv > $o/prof_$1_$tag.raw.csv 2>> $o/ncu_$1_$tag.log
        rm -f $o/prof_$1_$tag.ncu-rep
    }
    cap ndt 'regex:k_(limits|rank|tile_prefix|offsets|scatter|stats|kl|select)' 9 9
    cap count 'regex:k_count' 15 4
    cap net 'regex:k_(gemm|fc|head12|tnet|trunk|softmax)' 26 26
    stamp done
else
    stamp launches
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $o/r2_launches_$tag.csv \
        python bench.py --steps 1 --warmup 3 --batch 512 --device-chunk 128 --no-train --no-cpu-baseline > $o/ncu_launches_$tag.log 2>&1
    stamp done
fi
