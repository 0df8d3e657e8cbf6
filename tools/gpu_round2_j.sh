#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "msca or ela or depthwise or layerwise" ) > gpurun_out/pytest_gpu_j.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log
for v in 0 1 3; do
B2Y_DW_SIMPLE=$v timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_j$v.json > gpurun_out/bench_j_bsyolo_dw$v.json 2> gpurun_out/bench_j_bsyolo_dw$v.err
done
B2Y_DW_SIMPLE=3 timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 --layers-out gpurun_out/layers_y11n_j3.json > gpurun_out/bench_j_y11n_dw3.json 2> gpurun_out/bench_j_y11n_dw3.err
tail -4 gpurun_out/pytest_gpu_j.log
for f in gpurun_out/bench_j_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
except Exception as e: print("ERR", e)
PY
done
python - <<'PY'
import json
for v in (0,1,3):
    L=json.load(open(f"gpurun_out/layers_bsyolo_j{v}.json"))["layers"]
    print(v, [(l["name"], round(l["ms"]*1e3,1)) for l in L if l["kind"] in ("dwconv","msca_gate")])
L=json.load(open("gpurun_out/layers_y11n_j3.json"))["layers"]
print([(l["name"], round(l["ms"]*1e3,1)) for l in L if l["kind"] in ("dwconv",)])
PY
