#!/bin/bash
mkdir -p gpurun_out
( time B2Y_DUAL_NET=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "predict_stream or submit or sahi or graph_replay or dropin or end_to_end or ela" ) > gpurun_out/pytest_gpu_i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
for dn in 0 1; do
B2Y_DUAL_NET=$dn timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_i_y11n_dual$dn.json 2> gpurun_out/bench_i_y11n_dual$dn.err
B2Y_DUAL_NET=$dn timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_i$dn.json > gpurun_out/bench_i_bsyolo_dual$dn.json 2> gpurun_out/bench_i_bsyolo_dual$dn.err
done
tail -4 gpurun_out/pytest_gpu_i.log
for f in gpurun_out/bench_i_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
except Exception as e: print("ERR", e)
PY
done
python - <<'PY'
import json
L=json.load(open("gpurun_out/layers_bsyolo_i0.json"))["layers"]
for l in L:
    if l["kind"] in ("msca_gate","ela_gate","rowcol_mean"): print(l["name"], l["kind"], round(l["ms"]*1e3,1))
PY
