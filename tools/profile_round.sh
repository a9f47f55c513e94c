#!/bin/bash
# Profile artefacts of a round (run on the GPU box; outputs under gpurun_out/, summaries are copied to profiles/ by hand):
#   launches_bench.csv   ncu launch list of the bench command (serialised, cold caches: compare shares)
#   full.csv             ncu --set full raw page of the hot kernels of one cfg2 forward
#   eval_full.csv        the same for the post-processing kernel
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
K='conv3x3_ws_kernel|recurrence_kernel|cvf_tc_kernel|refine_head_pre_kernel|conv3x3_tc_kernel|conv5x5s2'
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 66 -c 66 -o gpurun_out/${TAG}_full \
    python tools/ncu_target.py > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:depth_metrics_kernel -s 2 -c 1 -o gpurun_out/${TAG}_eval \
    python tools/eval_target.py > gpurun_out/${TAG}_ncu_eval.log 2>&1
ncu -i gpurun_out/${TAG}_eval.ncu-rep --page raw --csv > gpurun_out/${TAG}_eval_full.csv 2>/dev/null
timeout 100 python tools/eval_target.py 2>&1 | tail -1 > gpurun_out/${TAG}_eval_timing.log
ls -la gpurun_out/ | tail -12
du -sh gpurun_out/${TAG}_full.ncu-rep
# keep the merged-back payload small: the CSV pages carry what the summaries need
if [ $(stat -c %s gpurun_out/${TAG}_full.ncu-rep) -gt 30000000 ]; then rm gpurun_out/${TAG}_full.ncu-rep; fi
tail -3 gpurun_out/${TAG}_ncu_full.log; cat gpurun_out/${TAG}_eval_timing.log
