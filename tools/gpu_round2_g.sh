#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "msca or ela or layerwise or end_to_end or sahi" ) > gpurun_out/pytest_gpu_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_yolo11n_g.json ) > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_g.json > gpurun_out/bench_g_bsyolo.json 2> gpurun_out/bench_g_bsyolo.err
tail -4 gpurun_out/pytest_gpu_g.log; tail -3 gpurun_out/bench_g.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[0])
print("y11n", d["value"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["clocks"])
for k,v in d.get("workloads",{}).items(): print("  ",k, v.get("value"), (v.get("e2e") or {}).get("value"), v.get("error"), v.get("bench_wall_s"), v.get("host_staging_ceiling"))
d=json.loads(open("gpurun_out/bench_g_bsyolo.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
L=json.load(open("gpurun_out/layers_bsyolo_g.json"))["layers"]
for l in L:
    if l["kind"] in ("pmsfa_dw","msca_branch","msca_gate","wsum4","rowcol_mean","ela_gate","ela_apply"): print(l["name"], l["kind"], round(l["ms"]*1e3,1))
PY
