#!/bin/bash
# ncu --set full of the new BS-YOLO SIMT kernels (one eager pass of bsyolo_n at batch 64)
mkdir -p gpurun_out
CMD="python tools/profile_step.py bsyolo_n 64 1"
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"pmsfa_dw_kernel|msca_branch_kernel|ela_stats_kernel|ela_apply_band_kernel|dwconv4_kernel|ela_gate_kernel" -c 14 \
    -f -o gpurun_out/prof_r02_bsyolo_simt $CMD > gpurun_out/ncu_c.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_c.log; tail -5 gpurun_out/ncu_c.log; ls -la gpurun_out/*.ncu-rep
