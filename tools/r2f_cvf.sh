#!/bin/bash
# Parity tests + stage profiles of the configurations the Conv3d filter dominates (after a cvf_tc.cu change).
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4)
for cfg in "512 640 1 64 1" "512 640 4 64 8" "1024 1280 4 128 4"; do
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 200 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|event mean" | tail -2
done
