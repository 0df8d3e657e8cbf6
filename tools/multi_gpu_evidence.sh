#!/bin/bash
# multi-GPU pass: NCCL tests (2 ranks), H2D ceiling probe, bench at N ranks (headline + BS-YOLO strong scaling + SAHI)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
( nvidia-smi topo -m; lscpu | grep -i -E 'numa|socket|model name|^cpu\(s\)'; nproc; free -g | head -2 ) > gpurun_out/topo_n$N.txt 2>&1
( time timeout 420 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/pytest_multi_n$N.log 2>&1
if [ "$N" = "2" ]; then ( time timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "msca or ela or layerwise" ) > gpurun_out/pytest_gates_n$N.log 2>&1; tail -3 gpurun_out/pytest_gates_n$N.log; fi
timeout 300 $TR tools/h2d_probe.py > gpurun_out/h2d_probe_n$N.json 2> gpurun_out/h2d_probe_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --extras bsyolo_n,sahi > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
NUMA=$(lscpu | grep -i "NUMA node(s)" | awk '{print $NF}')
if [ "$NUMA" != "1" ]; then   # binding / split uploads can only matter on a box that exposes more than one NUMA node
B2Y_BIND=0 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n${N}_nobind.json 2> gpurun_out/bench_n${N}_nobind.err
B2Y_H2D_SPLIT=2 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n${N}_split.json 2> gpurun_out/bench_n${N}_split.err
fi
tail -3 gpurun_out/pytest_multi_n$N.log; cat gpurun_out/h2d_probe_n$N.json
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nobind.json gpurun_out/bench_n${N}_split.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["config"].get("cpu_binding"))
    for k,v in d.get("workloads",{}).items(): print("   ",k, v.get("value"), v.get("e2e",{}).get("value"), v.get("error"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
done
