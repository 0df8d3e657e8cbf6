#!/bin/bash
# first GPU pass of round 2: tests, smoke, bench with sub-records, per-op tables, topology
mkdir -p gpurun_out
rm -f gpurun_out/parity_audit.jsonl
( nvidia-smi topo -m; lscpu | grep -i -E 'numa|socket|model name|^cpu\(s\)'; nproc; free -g | head -2 ) > gpurun_out/topo.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/smoke.log 2>&1
( time timeout 1200 python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_yolo11n.json ) > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
( time timeout 600 python bench.py --workload bsyolo_n --no-extras --no-cpu-baseline --steps 10 --warmup 3 --layers-out gpurun_out/layers_bsyolo.json ) > gpurun_out/bench_bsyolo.json 2> gpurun_out/bench_bsyolo.err
timeout 120 python tools/h2d_probe.py > gpurun_out/h2d_probe_n1.json 2> gpurun_out/h2d_probe_n1.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; head -c 600 gpurun_out/bench_a.json
