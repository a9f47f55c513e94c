#!/bin/bash
# One GPU-box pass over HEAD: gpu tests, bench line, warm timing + recurrence phase profile, ncu launch list.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/gpu_timing.py > gpurun_out/timing.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/timing.log
