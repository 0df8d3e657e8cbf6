#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_audit.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
for dn in 0 1; do
  B2Y_DUAL_NET=$dn timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_b_y11n_dual$dn.json 2> gpurun_out/bench_b_y11n_dual$dn.err
  B2Y_DUAL_NET=$dn timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_b$dn.json > gpurun_out/bench_b_bsyolo_dual$dn.json 2> gpurun_out/bench_b_bsyolo_dual$dn.err
done
tail -5 gpurun_out/pytest_gpu_b.log
for f in gpurun_out/bench_b_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
except Exception as e: print("ERR", e)
PY
done
