#!/bin/bash
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -4)
for cfg in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8" "1024 1280 4 128 4"; do
  echo "== $cfg"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
