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
# OPTS="early_d2h=0,left_late=1" sets library options; AB="early_d2h" alternates that option off / on in one process
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("="); net.set_option(k, int(v))
ab = os.environ.get("AB")
with torch.no_grad():
    for _ in range(3): net(*host, *flags)
    for rnd in range(6 if ab else 1):
        if ab: net.set_option(ab, rnd & 1)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); net(*host, *flags); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print((f"{ab}={rnd & 1}: " if ab else "") + f"e2e ms per call: median {ts[len(ts) // 2]:.3f} min {ts[0]:.3f} max {ts[-1]:.3f}")
