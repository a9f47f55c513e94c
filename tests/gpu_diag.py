"""Prints per-stage CUDA-vs-oracle errors for a few cases (run on the GPU box):

    python -m tests.gpu_diag [case ...]
"""
import sys
import time
import traceback

import torch

from multi_view_stereonet_b200 import synthetic
from tests._gpu_util import format_report, make_net, run_case
from tests._util import load_gta_state

CASES = {
    "cfg1": dict(rows=64, cols=80, views=1, hyps=8, batch=1, smooth=False),
    "cfg1_smooth": dict(rows=64, cols=80, views=1, hyps=8, batch=1, smooth=True),
    "mv_small": dict(rows=96, cols=128, views=2, hyps=6, batch=2, smooth=True),
    "odd_small": dict(rows=68, cols=90, views=3, hyps=5, batch=1, smooth=True),
    "mid": dict(rows=256, cols=320, views=2, hyps=16, batch=2, smooth=False),
    "cfg3_b2": dict(rows=512, cols=640, views=4, hyps=64, batch=2, smooth=False),
    "cfg2": dict(rows=512, cols=640, views=1, hyps=64, batch=1, smooth=False),
    "cfg2_smooth": dict(rows=512, cols=640, views=1, hyps=64, batch=1, smooth=True),
}


def main():
    names = sys.argv[1:] or ["cfg1", "cfg1_smooth", "mv_small", "odd_small", "mid", "cfg2"]
    state = load_gta_state()
    net = make_net(state)
    for name in names:
        c = CASES[name]
        print(f"== {name} {c}", flush=True)
        try:
            inputs = synthetic.make_inputs(c["rows"], c["cols"], c["views"], c["batch"], smooth=c["smooth"])
            t = time.time()
            rep, _, _ = run_case(net, state, inputs, c["hyps"])
            print(format_report(rep))
            print(f"  ({time.time() - t:.1f}s, {net.last_launch_count()} launches)", flush=True)
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
