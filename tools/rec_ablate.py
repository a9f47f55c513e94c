"""Timing ablations of the persistent recurrence kernel (results are wrong while a flag is set)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
sd, _ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
inp = synthetic.to_device(synthetic.make_inputs(512, 640, 1, 1), "cuda")
def t(label):
    with torch.no_grad():
        for _ in range(3): net(*inp, 64, True, [True] * 5)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): net(*inp, 64, True, [True] * 5)
        b.record(); torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b) / 10:.3f} ms per forward", flush=True)
t("baseline")
for flag, name in ((1, "skip MMAs"), (2, "skip gathers"), (4, "skip halo pushes"), (7, "skip all three")):
    net.set_option("recurrence_debug", flag); t(name)
net.set_option("recurrence_debug", 0)
