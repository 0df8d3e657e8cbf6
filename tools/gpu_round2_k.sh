#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/profile_step.py bsyolo_n 256 1"
$CMD > gpurun_out/plain_k.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"msca_gate_kernel|ela_gate_kernel|ela_stats_kernel|wsum4_band_kernel" -c 7 \
    -f -o gpurun_out/prof_r02_bsyolo_glue $CMD > gpurun_out/ncu_k.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_k.log; tail -3 gpurun_out/ncu_k.log
