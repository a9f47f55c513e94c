#!/bin/bash
# Stage-boundary timing of the cfg2 forward (B200MVS_STAGE_PROFILE hook in api.cu) + one bench line.
mkdir -p gpurun_out
B200MVS_STAGE_PROFILE=1 FORWARDS=6 timeout 200 python tools/ncu_target.py 2>&1 | grep "stage profile" | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['step_ms_rank0'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['latency_kernel']['kernel_ms'])"
