#!/bin/bash
mkdir -p gpurun_out
{
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -8)
for cfg in "1024 1280 4 128 1" "1024 1280 4 128 4"; do
  echo "== $cfg"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
} > gpurun_out/r2_step1.log 2>&1
cat gpurun_out/r2_step1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_b8.csv python tools/stage_cfg.py 512 640 1 64 8 > gpurun_out/r2_ncu_b8.log 2>&1
tail -2 gpurun_out/r2_ncu_b8.log
