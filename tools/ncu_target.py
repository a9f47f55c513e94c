"""Two cfg2 forwards for ncu captures, e.g.
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 30 -c 3 -o gpurun_out/prof python tools/ncu_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic  # noqa: E402

sd, _ = bench.load_state()
net = MultiViewStereoNet()
net.load_state_dict(sd)
net = net.cuda().eval()
inp = synthetic.to_device(synthetic.make_inputs(512, 640, int(os.environ.get("VIEWS", "1")), int(os.environ.get("BATCH", "1"))), "cuda")
with torch.no_grad():
    for _ in range(int(os.environ.get("FORWARDS", "2"))):
        net(*inp, 64, True, [True] * 5)
torch.cuda.synchronize()
