#!/bin/bash
# GEMM probe on the GPU box: all cases in one worker process, restarted after a crash.
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $LOG 2>&1
timeout 900 python tools/gemm_probe.py --all >> $LOG 2>&1
if grep -q FAIL $LOG; then
  for v in "4096 1024" "512 4096" "1024 4096" "4096 256" "128 512"; do
    set -- $v
    echo "--- NPM_MN_LBO=$1 NPM_MN_SBO=$2 ---" >> $LOG
    for c in km:tf32:128x64x32 mk:tf32:128x64x32 mm:tf32:128x128x64; do
      NPM_MN_LBO=$1 NPM_MN_SBO=$2 timeout 120 python tools/gemm_probe.py $c 2>&1 | head -3 >> $LOG
    done
  done
fi
grep -E "^CASE|WORKER|ERROR|^---" $LOG | cut -c1-175 | tail -70
