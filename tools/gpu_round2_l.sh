#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "depthwise or msca or ela or layerwise or end_to_end or heads_vs" ) > gpurun_out/pytest_gpu_l.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
for mc in 128 0; do
B2Y_DW_DENSE_MAXC=$mc timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 --layers-out gpurun_out/layers_y11n_l$mc.json > gpurun_out/bench_l_y11n_mc$mc.json 2> gpurun_out/bench_l_y11n_mc$mc.err
B2Y_DW_DENSE_MAXC=$mc timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo_l$mc.json > gpurun_out/bench_l_bsyolo_mc$mc.json 2> gpurun_out/bench_l_bsyolo_mc$mc.err
done
B2Y_DW_DENSE_MAXC=0 timeout 600 python bench.py --workload yolo11x --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_l_11x_mc0.json 2> gpurun_out/bench_l_11x_mc0.err
timeout 600 python bench.py --workload yolo11x --no-extras --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_l_11x_mc128.json 2> gpurun_out/bench_l_11x_mc128.err
tail -4 gpurun_out/pytest_gpu_l.log
for f in gpurun_out/bench_l_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"], d["roofline"]["frac"], d["roofline"]["network_roofline_frac"], d["roofline"].get("step_roofline_frac"))
except Exception as e: print("ERR", e)
PY
done
python - <<'PY'
import json
for f in ("layers_y11n_l128","layers_y11n_l0","layers_bsyolo_l128","layers_bsyolo_l0"):
    L=json.load(open(f"gpurun_out/{f}.json"))["layers"]
    print(f, [(l["name"].replace("model.",""), round(l["ms"]*1e3,1)) for l in L if l["kind"]=="dwconv" or ("cv3" in l["name"] and l["name"].endswith(".0") ) or l["name"] in ("model.5.cv2","model.24.cv2") or "pe[" in l["name"]])
PY
