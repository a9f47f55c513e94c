"""Warm CUDA-event timing + stage-boundary profile of one configuration (B200MVS_STAGE_PROFILE hook in api.cu):

    B200MVS_STAGE_PROFILE=1 python tools/stage_cfg.py ROWS COLS VIEWS HYPS BATCH [OPTS like pdl=0,overlap=0]

The stage profile is printed by the library to stderr for every forward; the last lines are the warm ones.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic  # noqa: E402

rows, cols, views, hyps, batch = [int(x) for x in sys.argv[1:6]]
sd, _ = bench.load_state()
net = MultiViewStereoNet()
net.load_state_dict(sd)
net = net.cuda().eval()
net.mask_mode = os.environ.get("MASK_MODE", "dense")
for kv in filter(None, (sys.argv[6] if len(sys.argv) > 6 else "").split(",")):
    k, v = kv.split("=")
    net.set_option(k, int(v))
inp = synthetic.to_device(synthetic.make_inputs(rows, cols, views, batch), "cuda")
flags = (hyps, True, [True] * 5)
steps = int(os.environ.get("STEPS", "5"))
with torch.no_grad():
    for _ in range(2):
        net(*inp, *flags)
    torch.cuda.synchronize()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t0 = time.perf_counter()
    for i in range(steps):
        st[i].record()
        net(*inp, *flags)
        en[i].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
ms = [a.elapsed_time(b) for a, b in zip(st, en)]
mac, byt, _ = bench.algorithmic_work(rows, cols, views, hyps)
print(f"{rows}x{cols} V={views} D={hyps} B={batch}: event mean {sum(ms) / steps:.3f} ms (min {min(ms):.3f}), "
      f"wall/step {1e3 * wall / steps:.3f} ms, {batch / (sum(ms) / steps) * 1e3:.1f} depthmaps/s, "
      f"launches {net.last_launch_count()}, mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB torch", flush=True)
if os.environ.get("PROBE"):
    # event pairs around every launch of one kernel class (b200mvs_probe_select), e.g. PROBE=cvf_conv32
    net.probe_select(os.environ["PROBE"])
    with torch.no_grad():
        net(*inp, *flags)
    torch.cuda.synchronize()
    total_ms, launches = net.probe_read()
    net.probe_select("none")
    print(f"probe {os.environ['PROBE']}: {launches} launches, {1e3 * total_ms / max(1, launches):.1f} us each", flush=True)
if os.environ.get("SWEEP_PROF"):
    # per-phase cycle totals of CTA (0, 0) of the wide sweep (sweep_wide.cu) or rank 0 of the cluster kernel
    net.set_option("recurrence_profile", 1)
    with torch.no_grad():
        net(*inp, *flags)
    torch.cuda.synchronize()
    prof = net.get_stage("recurrence_profile", torch.int64).cpu()[:12].tolist()
    print("sweep phases (cycles per step): " + " ".join(f"{v / max(1, hyps - 1):.0f}" for v in prof)
          + f" | total {sum(prof[:10]) / max(1, hyps - 1):.0f}", flush=True)
