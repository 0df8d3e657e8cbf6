#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one eager pass + full captures of selected launches.
# usage: tools/ncu_capture.sh <tag> "<launch indices for --set full>"
set -u
TAG=${1:-r01}
SEL=${2:-"1 6 69 73 26 76"}
OUT=gpurun_out
CMD="python tools/profile_step.py yolo11n 64 1"
B2Y_PROFILE_TABLE=$OUT/optable_$TAG.json $CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --profile-from-start off --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
echo "list rc $?"
[ "$SEL" = "none" ] && SEL=""
for L in $SEL; do
  ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip $L --launch-count 1 \
      -f -o $OUT/prof_${TAG}_L$L $CMD > $OUT/ncu_full_${TAG}_L$L.log 2>&1
  echo "full L$L rc $?"
done
ls -la $OUT | tail -20
