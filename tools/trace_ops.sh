#!/bin/bash
# Pipeline-event traces (CTA 0, %globaltimer ns) of selected network ops, eager launches.  usage: tools/trace_ops.sh "61 5 4 0"
for op in $1; do
  echo "=== op $op" >> gpurun_out/trace_ops.log
  B2Y_TRACE_OP=$op python tools/profile_step.py yolo11n 64 2 2>&1 | grep "b2y trace" | tail -4 >> gpurun_out/trace_ops.log
done
