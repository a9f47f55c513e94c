#!/bin/bash
for cfg in "512 640 1 64 8" "512 640 4 64 8"; do
  echo "== $cfg"
  B200MVS_LANE_TRACE=1 STEPS=2 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "^lane|depthmaps/s" | tail -3
  STEPS=5 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "depthmaps/s" | tail -1
  B200MVS_LANES=1 STEPS=5 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "depthmaps/s" | tail -1
done
