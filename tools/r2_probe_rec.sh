#!/bin/bash
for o in "" "overlap=0" "conv0_precompute=0" "left_late=0"; do
  echo "== B=8 V=1 OPTS=$o"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py 512 640 1 64 8 "$o" 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
for o in "" "overlap=0"; do
  echo "== B=8 V=4 OPTS=$o"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py 512 640 4 64 8 "$o" 2>&1 | grep -E "stage profile|depthmaps/s" | tail -2
done
for b in 2 4 6 7; do
  echo "== B=$b V=1 overlap=0"
  B200MVS_STAGE_PROFILE=1 STEPS=3 timeout 300 python tools/stage_cfg.py 512 640 1 64 $b "overlap=0" 2>&1 | grep -E "stage profile" | tail -1
done
B200MVS_TC_PROFILE=1 FORWARDS=2 timeout 200 python tools/ncu_target.py 2>&1 | grep "^tc TH" | tail -45
