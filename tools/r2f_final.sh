#!/bin/bash
# Evidence pass after the sliding-window Conv3d filter (r2f): GPU tests, smoke, bench line, stage profiles of the batch
# configurations, ncu --set full of one warm cvf_tc_kernel launch at cfg3.
mkdir -p gpurun_out
T="timeout -s KILL"
($T 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu.log
($T 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
$T 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/bench.err
: > gpurun_out/r2f_timing.log
for c in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8" "1024 1280 4 128 4"; do
  B200MVS_STAGE_PROFILE=1 STEPS=3 $T 100 python tools/stage_cfg.py $c 2>&1 | tail -2 >> gpurun_out/r2f_timing.log
done
STEPS=1 $T 200 ncu --set full --clock-control none --import-source on -k regex:cvf_tc_kernel -s 8 -c 1 -o gpurun_out/r2f_cvf \
  python tools/stage_cfg.py 512 640 4 64 8 > gpurun_out/ncu_cvf.log 2>&1
ncu -i gpurun_out/r2f_cvf.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_cvf_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 300 gpurun_out/r2f_bench_cfg2.json; echo; cat gpurun_out/r2f_timing.log; tail -2 gpurun_out/ncu_cvf.log
