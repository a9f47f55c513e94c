#!/bin/bash
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4)
B200MVS_REC_OCC=1 STEPS=1 timeout 100 python tools/stage_cfg.py 512 640 1 64 1 2>&1 | grep "occupancy" | sort -u
NOPROF=1 timeout 300 python tools/gpu_timing.py 2>&1 | tail -2
B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py 512 640 1 64 1 2>&1 | grep -E "stage profile" | tail -1
B200MVS_TC_PROFILE=1 FORWARDS=2 timeout 200 python tools/ncu_target.py 2>&1 | grep "^tc TH" | tail -4
