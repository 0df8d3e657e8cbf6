#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "pmsfa or msca or ela or layerwise or end_to_end or heads_vs or sppf or depthwise" ) > gpurun_out/pytest_gpu_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_d.json > gpurun_out/bench_d_bsyolo.json 2> gpurun_out/bench_d_bsyolo.err
B2Y_DW_DENSE_MAXC=256 timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_d256.json > gpurun_out/bench_d_bsyolo_dw256.json 2> gpurun_out/bench_d_bsyolo_dw256.err
B2Y_DW_DENSE_MAXC=512 timeout 900 python bench.py --workload yolo11x --no-extras --no-cpu-baseline --steps 5 --warmup 3 --layers-out gpurun_out/layers_11x_dw512.json > gpurun_out/bench_d_11x_dw512.json 2> gpurun_out/bench_d_11x_dw512.err
tail -4 gpurun_out/pytest_gpu_d.log
for f in gpurun_out/bench_d_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
except Exception as e: print("ERR", e)
PY
done
