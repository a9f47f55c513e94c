#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_ws_kernel -s 18 -c 4 -o gpurun_out/r2_ws python tools/ncu_target.py > gpurun_out/r2_ncu_ws.log 2>&1
tail -3 gpurun_out/r2_ncu_ws.log
ls -la gpurun_out/r2_ws.ncu-rep
B200MVS_WS_PROFILE=1 FORWARDS=2 timeout 200 python tools/ncu_target.py 2>&1 | grep -A9 "^ws TH" | tail -75 > gpurun_out/r2_ws_timeline.log
tail -44 gpurun_out/r2_ws_timeline.log
