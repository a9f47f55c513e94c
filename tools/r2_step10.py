import sys, torch
sys.path.insert(0, '.')
import bench
from multi_view_stereonet_b200 import synthetic
from tests import _gpu_util
sd, _ = bench.load_state()
def run(rows, cols, hyps, smooth, **opts):
    net = _gpu_util.make_net(sd)
    for k, v in opts.items():
        net.set_option(k, v)
    rep, _, _ = _gpu_util.run_case(net, sd, synthetic.make_inputs(rows, cols, 1, 1, smooth=smooth), hyps, stages=False)
    print(rows, cols, hyps, smooth, opts, {k: f"{v:.1e}" for k, v in rep.items() if k.startswith("idepth")}, flush=True)
run(500, 636, 12, True)
run(500, 636, 12, True, half_activations=0)
run(500, 636, 12, True, tensor_cores=0)
run(512, 640, 12, True)
run(512, 640, 12, True, tensor_cores=0)
run(504, 632, 12, True)
run(500, 636, 64, True)
run(250, 318, 12, True)
