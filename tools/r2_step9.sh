#!/bin/bash
echo "== poly"; timeout 300 python tools/cfg_report.py 500 636 1 12 2>&1 | grep -E "idepth|raw|mask|left_feature" 
echo "== nopoly"; B200MVS_WS_NOPOLY=1 timeout 300 python tools/cfg_report.py 500 636 1 12 2>&1 | grep -E "idepth|raw|mask|left_feature"
echo "== no ws"; python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from multi_view_stereonet_b200 import synthetic
from tests import _gpu_util
sd, _ = bench.load_state()
net = _gpu_util.make_net(sd)
net.set_option("warp_specialized", 0)
rep, _, _ = _gpu_util.run_case(net, sd, synthetic.make_inputs(500, 636, 1, 1, smooth=True), 12)
print({k: v for k, v in rep.items() if k.startswith("idepth")})
PY
