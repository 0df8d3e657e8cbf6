#!/bin/bash
# Round-2 closing evidence run (one B200, trimmed to the GPU budget that is left): full GPU suite, smoke, both bench arms
# with all sub-records, BS-YOLO per-layer table, ncu launch list of bench.py's timed region (yolo11n).
# The `--set full` captures and the per-launch DRAM/tensor-pipe passes of tools/final_evidence.sh were taken earlier this
# round (profiles/r02_ncu_full_*.txt, r02_launches_*.csv) on the same conv kernels and are not repeated here.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_audit.jsonl
( time timeout 1200 python -m pytest tests -q -m gpu ) > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc $?"; tail -4 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc $?"; tail -1 $OUT/smoke_$TAG.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --layers-out $OUT/layers_yolo11n_$TAG.json ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc $?"
timeout 300 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out $OUT/layers_bsyolo_$TAG.json > $OUT/bench_bsyolo_$TAG.json 2> /dev/null; echo "bsyolo rc $?"
timeout 300 tools/ncu_bench.sh $TAG yolo11n > $OUT/ncu_bench_$TAG.out 2>&1; echo "ncu bench rc $?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("y11n", round(d["value"]), round(d["e2e"]["value"]), d["stages"], round(d["roofline"]["frac"],3), round(d["roofline"]["network_roofline_frac"],3), d["clocks"], d["cpu_baseline"])
for k,v in d.get("workloads",{}).items(): print("  ",k, v.get("value"), (v.get("e2e") or {}).get("value"), v.get("error"), v.get("bench_wall_s"), (v.get("roofline") or {}).get("frac"))
PY
