"""Times forwards of the cfg2 workload with CUDA events (async and per-step sync) and prints the
per-phase cycle profile of the persistent recurrence kernel.  Run on the GPU box:

    python tools/gpu_timing.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic  # noqa: E402


def main():
    sd, _ = bench.load_state()
    net = MultiViewStereoNet()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    batch = int(os.environ.get("BATCH", "1"))
    views = int(os.environ.get("VIEWS", "1"))
    inp = synthetic.to_device(synthetic.make_inputs(512, 640, views, batch), "cuda")
    flags = (64, True, [True] * 5)

    def run(label, sync_each, steps=20):
        with torch.no_grad():
            for _ in range(3):
                net(*inp, *flags)
            torch.cuda.synchronize()
            st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            t0 = time.perf_counter()
            for i in range(steps):
                st[i].record()
                net(*inp, *flags)
                en[i].record()
                if sync_each:
                    torch.cuda.synchronize()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        ms = [a.elapsed_time(b) for a, b in zip(st, en)]
        print(f"{label}: event mean {sum(ms) / steps:.3f} ms (min {min(ms):.3f} max {max(ms):.3f}), "
              f"wall/step {1e3 * wall / steps:.3f} ms, launches {net.last_launch_count()}", flush=True)

    # OPTS="pdl=0,overlap=0" toggles library options for A/B timing
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):
        k, v = kv.split("=")
        net.set_option(k, int(v))
    run("async", False)
    run("sync ", True)
    if os.environ.get("NOPROF"):
        return
    net.set_option("recurrence_profile", 1)
    with torch.no_grad():
        net(*inp, *flags)
    torch.cuda.synchronize()
    prof = net.get_stage("recurrence_profile", torch.int64).view(16, 12).cpu()
    names = ["W", "MMA0", "E0", "barA", "S1", "MMA1", "E1", "barC", "S2", "MMA2", "E2", "barE"]
    for r in (0, 5, 10):
        print("recurrence rank", r, " ".join(f"{n}={prof[r, i].item() / 63:.0f}" for i, n in enumerate(names)),
              " cycles/step", prof[r].sum().item() / 63)


if __name__ == "__main__":
    main()
