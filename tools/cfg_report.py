"""Stage-by-stage parity report of one image group of a configuration (tests/_gpu_util.run_case):
    python tools/cfg_report.py ROWS COLS VIEWS HYPS"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import synthetic
from tests import _gpu_util
rows, cols, views, hyps = [int(x) for x in sys.argv[1:5]]
sd, _ = bench.load_state()
net = _gpu_util.make_net(sd)
rep, _, _ = _gpu_util.run_case(net, sd, synthetic.make_inputs(rows, cols, views, 1), hyps)
print(_gpu_util.format_report(rep))
