#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "pmsfa or msca or ela or layerwise or end_to_end" ) > gpurun_out/pytest_gpu_e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_e.json > gpurun_out/bench_e_bsyolo.json 2> gpurun_out/bench_e_bsyolo.err
CMD="python tools/profile_step.py bsyolo_n 64 1"
$CMD > gpurun_out/plain_e.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"pmsfa_dw_kernel|msca_branch_kernel" -c 3 \
    -f -o gpurun_out/prof_r02_bsyolo_simt_v3 $CMD > gpurun_out/ncu_e.log 2>&1
echo "ncu rc=$?"
tail -4 gpurun_out/pytest_gpu_e.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_e_bsyolo.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
L=json.load(open("gpurun_out/layers_bsyolo_e.json"))["layers"]
for l in L:
    if l["kind"] in ("pmsfa_dw","msca_branch","rowcol_mean","ela_gate","ela_apply"): print(l["name"], l["kind"], round(l["ms"]*1e3,1))
PY
