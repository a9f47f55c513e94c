#!/bin/bash
# Round-2 closing evidence (one GPU): tests, smoke, bench (+ reference arm), warm timings and stage profiles of the four
# configurations, launch lists, ncu --set full of the hot kernels of a cfg2 forward and of the wide sweep (cfg3).
mkdir -p gpurun_out
T="timeout -s KILL"
($T 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log
($T 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
$T 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/bench.err
$T 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2>> gpurun_out/bench.err
$T 300 python tools/gpu_timing.py > gpurun_out/r2c_timing.log 2>&1
for c in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8" "1024 1280 4 128 4"; do
  B200MVS_STAGE_PROFILE=1 STEPS=3 $T 100 python tools/stage_cfg.py $c 2>&1 | tail -2 >> gpurun_out/r2c_timing.log
done
SWEEP_PROF=1 STEPS=2 $T 100 python tools/stage_cfg.py 1024 1280 4 128 4 2>&1 | tail -1 >> gpurun_out/r2c_timing.log
$T 100 python tools/eval_target.py 2>&1 | tail -1 >> gpurun_out/r2c_timing.log
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_forward_cfg2.csv python tools/ncu_target.py > gpurun_out/ncu.log 2>&1
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c_launches_bench_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/bench_under_ncu.log 2>&1
K='conv3x3_ws_kernel|recurrence_kernel|l4_tail_kernel|cvf_tc_kernel|refine_head_pre_kernel|conv3x3_tc_kernel|conv5x5s2|gather_plan_kernel|mask_vote_kernel'
FORWARDS=1 $T 600 ncu --set full --clock-control none -k "regex:$K" -c 80 -f -o gpurun_out/r2c_full \
    python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r2c_full.ncu-rep --page raw --csv > gpurun_out/r2c_full.csv 2>/dev/null
rm -f gpurun_out/r2c_full.ncu-rep
STEPS=1 $T 300 ncu --set full --clock-control none -k regex:sweep_wide -s 1 -c 1 -f -o gpurun_out/r2c_wide \
    python tools/stage_cfg.py 512 640 4 64 8 > gpurun_out/ncu_wide.log 2>&1
ncu -i gpurun_out/r2c_wide.ncu-rep --page raw --csv > gpurun_out/r2c_wide.csv 2>/dev/null
rm -f gpurun_out/r2c_wide.ncu-rep
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 400 gpurun_out/r2c_bench_cfg2.json; echo; head -c 300 gpurun_out/r2c_bench_reference_arm.json; echo; cat gpurun_out/r2c_timing.log; tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_wide.log; ls -la gpurun_out | grep r2c
