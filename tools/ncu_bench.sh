#!/bin/bash
# ncu launch list of bench.py's own timed region (the resident `value` steps), under gpurun.
# usage: tools/ncu_bench.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
WL=${2:-yolo11n}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > $OUT/bench_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/bench_plain_$TAG.log; exit 1; }
B2Y_CUDA_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/bench_launches_$TAG.csv $CMD > $OUT/ncu_bench_$TAG.log 2>&1
echo "ncu rc $?"; grep -c conv_tc $OUT/bench_launches_$TAG.csv
