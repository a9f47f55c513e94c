#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or stagewise" 2>&1 | tail -4)
NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2
DEBUG=0,4,8,12 timeout 300 python tools/rec_trace.py 2>&1 | tee gpurun_out/rec_trace.log | tail -150
B200MVS_REC_OCC=1 BATCH=8 NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | grep -E "occupancy|async" | tail -12
