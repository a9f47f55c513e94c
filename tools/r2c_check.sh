#!/bin/bash
# Round-2 (last session) check on one GPU: GPU tests, smoke, bench line with the configs block.
mkdir -p gpurun_out
(timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log
(timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
timeout -s KILL 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 600 gpurun_out/r2c_bench_cfg2.json; echo; tail -3 gpurun_out/bench.err
