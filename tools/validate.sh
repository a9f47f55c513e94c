#!/bin/bash
# One GPU-box pass over HEAD: gpu tests, smoke, bench line (+ reference arm), warm timing + recurrence phase profile,
# stage-boundary profile, post-processing kernel timing, ncu launch list of a cfg2 forward.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 300 python tools/gpu_timing.py > gpurun_out/timing.log 2>&1
B200MVS_STAGE_PROFILE=1 FORWARDS=5 timeout 200 python tools/ncu_target.py 2>&1 | grep "stage profile" | tail -1 >> gpurun_out/timing.log
timeout 100 python tools/eval_target.py 2>&1 | tail -1 >> gpurun_out/timing.log
BATCH=8 VIEWS=4 NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2 | sed 's/^/cfg3 (B=8, V=4): /' >> gpurun_out/timing.log
BATCH=8 VIEWS=1 NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2 | sed 's/^/cfg4 per GPU (B=8, V=1): /' >> gpurun_out/timing.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_reference.json; cat gpurun_out/timing.log
