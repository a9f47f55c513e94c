#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/gpu_timing.py > gpurun_out/timing.log 2>&1
B200MVS_STAGE_PROFILE=1 FORWARDS=5 timeout 200 python tools/ncu_target.py 2>&1 | grep "stage profile" | tail -1 >> gpurun_out/timing.log
BATCH=8 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_b8.csv python tools/ncu_target.py > gpurun_out/ncu_b8.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/timing.log
