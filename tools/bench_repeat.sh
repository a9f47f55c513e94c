#!/bin/bash
# Repeats the default bench line N times and prints value / per-step times (outlier hunting).
for i in $(seq 1 ${1:-3}); do timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), d['clocks'], 'max step', max(d['step_ms_rank0']), 'first', d['step_ms_rank0'][:3], 'e2e', round(d['e2e']['ms_per_step'],4))"; done
