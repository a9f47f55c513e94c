#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recurrence_kernel -s 1 -c 1 -o gpurun_out/r2_rec python tools/ncu_target.py > gpurun_out/r2_ncu_rec.log 2>&1
tail -3 gpurun_out/r2_ncu_rec.log
ls -la gpurun_out/r2_rec.ncu-rep
timeout 300 python tools/gpu_timing.py 2>&1 | tail -4
