#!/bin/bash
for o in "" "overlap=0" "conv0_precompute=0" "left_late=0" "lanes=2"; do
  echo "== B=8 V=1 OPTS=$o"
  STEPS=5 timeout 300 python tools/stage_cfg.py 512 640 1 64 8 "$o" 2>&1 | grep -E "depthmaps/s" | tail -1
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py 512 640 1 64 8 "$o" 2>&1 | grep -E "stage profile" | tail -1
done
for o in "" "overlap=0" "lanes=2"; do
  echo "== B=8 V=4 OPTS=$o"
  STEPS=5 timeout 300 python tools/stage_cfg.py 512 640 4 64 8 "$o" 2>&1 | grep -E "depthmaps/s" | tail -1
done
