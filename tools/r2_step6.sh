#!/bin/bash
(B200MVS_LANES=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "cfg3_item_multiview or batch8" 2>&1 | tail -3)
for cfg in "512 640 1 64 8" "512 640 4 64 8"; do
  echo "== $cfg"
  B200MVS_LANE_TRACE=1 STEPS=2 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "^lane|depthmaps/s" | tail -9
done
