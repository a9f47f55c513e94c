#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_pytest_gpu.log
cat gpurun_out/r2_pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -3 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],'rec us/step',d['latency_kernel']['us_per_step'])
for k,v in (d.get('configs') or {}).items():
    print(k, round(v['ms_per_step'],3),'ms', round(v['depthmaps_per_s'],1),'dm/s frac', round(v['frac_of_roofline'],3), v['stage_us_rank0'])
PY
