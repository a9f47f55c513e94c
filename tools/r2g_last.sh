#!/bin/bash
# Closing pass of round 2 at HEAD: GPU tests, smoke, bench line.
mkdir -p gpurun_out
T="timeout -s KILL"
($T 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu.log
($T 100 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/smoke.log
$T 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_cfg2.json 2> gpurun_out/bench.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 400 gpurun_out/r2g_bench_cfg2.json; echo
