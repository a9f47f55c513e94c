#!/bin/bash
# Last pass of round 2 (after the mask-upsampling change): GPU tests, smoke, bench line, stage profiles of the batch
# configurations, ncu of the mask-upsampling launches at batch 8.
mkdir -p gpurun_out
T="timeout -s KILL"
($T 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu.log
($T 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
$T 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_cfg2.json 2> gpurun_out/bench.err
: > gpurun_out/r2d_timing.log
for c in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8" "1024 1280 4 128 4"; do
  B200MVS_STAGE_PROFILE=1 STEPS=3 $T 100 python tools/stage_cfg.py $c 2>&1 | tail -2 >> gpurun_out/r2d_timing.log
done
STEPS=1 $T 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:upsample_mask -c 8 --csv \
  --log-file gpurun_out/r2d_ncu_mask_upsampling.csv python tools/stage_cfg.py 512 640 1 64 8 > gpurun_out/ncu_mask.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 300 gpurun_out/r2d_bench_cfg2.json; echo; cat gpurun_out/r2d_timing.log; tail -12 gpurun_out/r2d_ncu_mask_upsampling.csv | cut -c1-250
