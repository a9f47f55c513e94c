#!/bin/bash
(timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -12)
B200MVS_WS_PROFILE=1 FORWARDS=2 timeout 200 python tools/ncu_target.py 2>&1 | grep "^ws TH" | tail -12
for cfg in "512 640 1 64 1" "512 640 1 64 8"; do
  echo "== $cfg"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
  B200MVS_WS_NOPOLY=1 B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
