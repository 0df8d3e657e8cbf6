#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "msca or ela or layerwise or end_to_end" ) > gpurun_out/pytest_gpu_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
timeout 300 python tools/sahi_probe.py 4 > gpurun_out/sahi_probe.log 2>&1
timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_h.json > gpurun_out/bench_h_bsyolo.json 2> gpurun_out/bench_h_bsyolo.err
tail -4 gpurun_out/pytest_gpu_h.log; cat gpurun_out/sahi_probe.log | tail -8
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_h_bsyolo.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
L=json.load(open("gpurun_out/layers_bsyolo_h.json"))["layers"]
for l in L:
    if l["kind"] in ("msca_branch","msca_gate","wsum4","rowcol_mean","ela_gate","ela_apply"): print(l["name"], l["kind"], round(l["ms"]*1e3,1))
PY
