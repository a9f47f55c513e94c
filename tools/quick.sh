#!/bin/bash
# Fast GPU check after a kernel change: parity tests (-k filter as $1, default: all of test_parity_gpu) + warm timing.
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -6)
NOPROF=${NOPROF:-1} timeout 300 python tools/gpu_timing.py 2>&1 | tail -6
B200MVS_STAGE_PROFILE=1 FORWARDS=5 timeout 200 python tools/ncu_target.py 2>&1 | grep "stage profile" | tail -1
