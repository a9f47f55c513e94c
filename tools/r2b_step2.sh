#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or stagewise or cfg3" 2>&1 | tail -4)
NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2
DEBUG=${DEBUG:-0,4,20,36} timeout 300 python tools/rec_trace.py 2>&1 | tee gpurun_out/rec_trace.log | grep -v "^rank5" | tail -130
