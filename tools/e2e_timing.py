"""Where the end-to-end (host tensors in, host tensors out) time goes.  B200MVS_HOST_PROFILE=1 prints the library's
own split (enqueue uploads / enqueue kernels / wait)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic
sd, _ = bench.load_state(); net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
cpu = synthetic.make_inputs(512, 640, 1, 1)
pin = lambda t: t.pin_memory()
host = ([pin(t) for t in cpu[0]], [pin(t) for t in cpu[1]], [pin(t) for t in cpu[2]], [[pin(t) for t in p] for p in cpu[3]])
out = {"left_idepthmap_pyr": [torch.empty((1, 1) + tuple(t.shape[-2:]), dtype=torch.float32).pin_memory() for t in cpu[0]]}
net.set_host_outputs(out)
flags = (64, True, [True] * 5)
with torch.no_grad():
    for _ in range(3): net(*host, *flags)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); net(*host, *flags); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e ms per call:", " ".join(f"{t:.3f}" for t in ts))
