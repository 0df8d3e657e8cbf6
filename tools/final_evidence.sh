#!/bin/bash
# Round-end evidence run (under gpurun): full GPU suite, both bench arms, per-layer table, ncu launch lists + full captures.
# usage: tools/final_evidence.sh <tag> "<launch indices for --set full>"
TAG=${1:-r01}
SEL=${2:-"1 62"}
OUT=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc $?"; tail -2 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --layers-out $OUT/layers_$TAG.json > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc $?"
timeout 300 python bench.py --workload bsyolo_n --no-cpu-baseline > $OUT/bench_bsyolo_$TAG.json 2> /dev/null; echo "bsyolo rc $?"
timeout 300 python bench.py --workload yolo11x --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_11x_$TAG.json 2> /dev/null; echo "11x rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc $?"
timeout 900 tools/ncu_capture.sh $TAG "$SEL" > $OUT/ncu_$TAG.out 2>&1; echo "ncu rc $?"
timeout 600 tools/ncu_bench.sh $TAG > $OUT/ncu_bench_$TAG.out 2>&1; echo "ncu bench rc $?"
