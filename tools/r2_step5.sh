#!/bin/bash
echo "== lanes=1"
(B200MVS_LANES=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "cfg3_item_multiview or batch8" 2>&1 | tail -30)
echo "== lanes default"
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -30)
for cfg in "512 640 1 64 1" "512 640 1 64 8" "512 640 4 64 8"; do
  echo "== $cfg"
  STEPS=5 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "depthmaps/s" | tail -1
  B200MVS_LANES=1 STEPS=5 timeout 300 python tools/stage_cfg.py $cfg 2>&1 | grep -E "depthmaps/s" | tail -1
done
