#!/bin/bash
# Round-end evidence run (under gpurun, one B200): full GPU suite, smoke, both bench arms with all sub-records, per-layer
# tables, ncu launch lists of bench.py's timed regions (yolo11n, BS-YOLO-n), per-launch counters of one eager pass
# (yolo11n, yolo11x) and two `--set full` captures.
# usage: tools/final_evidence.sh <tag> "<launch indices of the yolo11n pass for --set full>"
TAG=${1:-r02}
SEL=${2:-"1 62"}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_audit.jsonl
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc $?"; tail -4 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc $?"; tail -1 $OUT/smoke_$TAG.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 --layers-out $OUT/layers_yolo11n_$TAG.json ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc $?"
timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out $OUT/layers_bsyolo_$TAG.json > $OUT/bench_bsyolo_$TAG.json 2> /dev/null; echo "bsyolo rc $?"
timeout 600 python bench.py --workload yolo11x --no-extras --no-cpu-baseline --steps 5 --warmup 3 --layers-out $OUT/layers_11x_$TAG.json > $OUT/bench_11x_$TAG.json 2> /dev/null; echo "11x rc $?"
timeout 900 tools/ncu_capture.sh $TAG "$SEL" > $OUT/ncu_$TAG.out 2>&1; echo "ncu capture rc $?"
timeout 600 tools/ncu_bench.sh $TAG yolo11n > $OUT/ncu_bench_$TAG.out 2>&1; echo "ncu bench rc $?"
timeout 600 tools/ncu_bench.sh ${TAG}_bsyolo bsyolo_n > $OUT/ncu_bench_${TAG}_bsyolo.out 2>&1; echo "ncu bench bsyolo rc $?"
# yolo11x: per-launch time, DRAM bytes and tensor-pipe activity of one eager pass at batch 8 (the (6400,192,1728) GEMMs included)
CMDX="python tools/profile_step.py yolo11x 8 1"
B2Y_PROFILE_TABLE=$OUT/optable_${TAG}_11x.json $CMDX > $OUT/plain_${TAG}_11x.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --profile-from-start off --csv --log-file $OUT/launches_${TAG}_11x.csv $CMDX > $OUT/ncu_list_${TAG}_11x.log 2>&1
echo "11x list rc $?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("y11n", round(d["value"]), round(d["e2e"]["value"]), d["stages"], round(d["roofline"]["frac"],3), round(d["roofline"]["network_roofline_frac"],3), d["clocks"], d["cpu_baseline"])
for k,v in d.get("workloads",{}).items(): print("  ",k, v.get("value"), (v.get("e2e") or {}).get("value"), v.get("error"), v.get("bench_wall_s"), v.get("host_staging_ceiling"), (v.get("roofline") or {}).get("frac"))
PY
ls -la $OUT | tail -25
