"""Runs one BASELINE configuration (or one item of it) on the GPU, checks it against the CPU oracle and times it.

    python tools/cfg_check.py ROWS COLS VIEWS HYPS BATCH [--no-oracle]
"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
from oracle import mvsnet_oracle as oracle
from tests._util import rel_linf

rows, cols, views, hyps, batch = [int(x) for x in sys.argv[1:6]]
sd, _ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
cpu = synthetic.make_inputs(rows, cols, views, batch)
dev = synthetic.to_device(cpu, "cuda")
flags = (hyps, True, [True] * 5)
with torch.no_grad():
    out = net(*dev, *flags); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): out = net(*dev, *flags)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
print(f"{rows}x{cols} V={views} D={hyps} B={batch}: {ms:.2f} ms per forward = {batch / ms * 1e3:.1f} depthmaps/s, "
      f"{net.last_launch_count()} launches", flush=True)
if "--no-oracle" not in sys.argv:
    one = synthetic.make_inputs(rows, cols, views, 1)     # item 0 (seeded per item)
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oracle.forward(sd, *one, *flags[:2], tuple(flags[2]))
    print(f"oracle item 0: {time.perf_counter() - t0:.1f} s on {torch.get_num_threads()} threads")
    for lvl in range(5):
        err = rel_linf(out["left_idepthmap_pyr"][lvl][:1].cpu(), ref["left_idepthmap_pyr"][lvl])
        mism = int((out["left_idepthmap_mask_pyr"][lvl][:1].cpu() != ref["left_idepthmap_mask_pyr"][lvl]).sum())
        print(f"  level {lvl}: rel L-inf {err:.2e}, mask mismatches {mism}")
