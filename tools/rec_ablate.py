"""Timing ablations of the persistent recurrence kernel (results are wrong while a flag is set).
Prints the per-phase cycle profile (rank 5 of the cluster) for each flag combination."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
sd, _ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
inp = synthetic.to_device(synthetic.make_inputs(512, 640, 1, 1), "cuda")
names = ["W", "MMA0", "E0", "barA", "S1", "MMA1", "E1", "barC", "S2", "MMA2", "E2", "barE"]
net.set_option("recurrence_profile", 1)
def t(label):
    with torch.no_grad():
        for _ in range(3): net(*inp, 64, True, [True] * 5)
        torch.cuda.synchronize()
    prof = net.get_stage("recurrence_profile", torch.int64).view(16, 12).cpu()
    for r in (0, 5):
        print(f"{label:>18} r{r}:", " ".join(f"{n}={prof[r, i].item() / 63:.0f}" for i, n in enumerate(names)),
              f" total={prof[r].sum().item() / 63:.0f}", flush=True)
flags = [int(x) for x in os.environ.get("FLAGS", "0,1,2,4,8").split(",")]
for flag in flags:
    net.set_option("recurrence_debug", flag); t(f"debug={flag}")
net.set_option("recurrence_debug", 0)
