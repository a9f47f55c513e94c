#!/bin/bash
# Round-2 starting point: stage profiles of the BASELINE throughput configs at HEAD of round 1.
mkdir -p gpurun_out
{
for cfg in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8" "1024 1280 4 128 1" "1024 1280 4 128 4"; do
  echo "== $cfg"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
} > gpurun_out/r2_baseline_stages.log 2>&1
cat gpurun_out/r2_baseline_stages.log
nvidia-smi --query-gpu=name,memory.total --format=csv
