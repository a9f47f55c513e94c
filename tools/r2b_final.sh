#!/bin/bash
# Round-2 final evidence (one GPU): tests, smoke, bench (+ reference arm), warm timing + sweep profile, stage profile,
# launch lists, ncu --set full of the hot kernels, sweep timeline.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_cfg2.json 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2>> gpurun_out/bench.err
timeout 300 python tools/gpu_timing.py > gpurun_out/r2_final_timing.log 2>&1
B200MVS_STAGE_PROFILE=1 FORWARDS=5 timeout 200 python tools/ncu_target.py 2>&1 | grep "stage profile" | tail -1 >> gpurun_out/r2_final_timing.log
timeout 100 python tools/eval_target.py 2>&1 | tail -1 >> gpurun_out/r2_final_timing.log
BATCH=8 VIEWS=4 NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2 | sed 's/^/cfg3 (B=8, V=4): /' >> gpurun_out/r2_final_timing.log
BATCH=8 VIEWS=1 NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2 | sed 's/^/cfg4 per GPU (B=8, V=1): /' >> gpurun_out/r2_final_timing.log
DEBUG=0 timeout 200 python tools/rec_trace.py > gpurun_out/r2_final_sweep_timeline.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches_forward_cfg2.csv python tools/ncu_target.py > gpurun_out/ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_final_launches_bench_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/bench_under_ncu.log 2>&1
K='conv3x3_ws_kernel|recurrence_kernel|l4_tail_kernel|cvf_tc_kernel|refine_head_pre_kernel|conv3x3_tc_kernel|conv5x5s2|gather_plan_kernel|mask_vote_kernel'
FORWARDS=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 80 -o gpurun_out/r2_final_full \
    python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r2_final_full.ncu-rep --page raw --csv > gpurun_out/r2_final_full.csv 2>/dev/null
if [ -f gpurun_out/r2_final_full.ncu-rep ] && [ $(stat -c %s gpurun_out/r2_final_full.ncu-rep) -gt 30000000 ]; then rm gpurun_out/r2_final_full.ncu-rep; fi
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 1500 gpurun_out/r2_final_bench_cfg2.json; echo; cat gpurun_out/r2_final_bench_reference_arm.json; cat gpurun_out/r2_final_timing.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | tail -20
