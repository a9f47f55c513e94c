"""Per-warp timeline of one step of the persistent recurrence kernel (option "recurrence_profile": clock64 of
lane 0 of every warp at ~26 points of step D/2; point 0 = right behind the previous step's cluster barrier,
the common origin of all CTAs).  Prints, per traced point, min / median / max over the warps of a CTA for
ranks 0, 5 and 10, and the slowest (rank, warp) over the cluster.

    DEBUG=0,4 python tools/rec_trace.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multi_view_stereonet_b200 import MultiViewStereoNet, synthetic

POINTS = ["start", "tma_done", "gathered", "sync_W", "plan/issue0", "conv0_done", "E0_ld", "E0_sent", "barA",
          "coef1", "own1", "halo1", "sync_S1", "conv1_done", "E1_ld", "E1_sent", "barC", "coef2", "own2", "halo2",
          "sync_S2", "conv2_done", "E2_ld", "E2_stored", "cl_barrier", "tma_issued", "E0_sts", "E0_fence", "E0_bulk", "E0_shfl"]


def main():
    sd, _ = bench.load_state()
    net = MultiViewStereoNet(); net.load_state_dict(sd); net = net.cuda().eval()
    inp = synthetic.to_device(synthetic.make_inputs(512, 640, 1, 1), "cuda")
    net.set_option("recurrence_profile", 1)
    names = ["W", "MMA0", "E0", "barA", "S1", "MMA1", "E1", "barC", "S2", "MMA2", "E2", "barE"]
    for dbg in [int(x) for x in os.environ.get("DEBUG", "0").split(",")]:
        net.set_option("recurrence_debug", dbg)
        with torch.no_grad():
            for _ in range(3):
                net(*inp, 64, True, [True] * 5)
            torch.cuda.synchronize()
        prof = net.get_stage("recurrence_profile", torch.int64).view(16, 12).cpu()
        tr = net.get_stage("recurrence_trace", torch.int64).view(16, 16, 32).cpu()
        print(f"=== debug={dbg}")
        for r in (0, 5, 10):
            print(f"rank {r}:", " ".join(f"{n}={prof[r, i].item() / 63:.0f}" for i, n in enumerate(names)),
                  f" total={prof[r].sum().item() / 63:.0f}")
        t0 = tr[:11, :, 0:1].clone()
        rel = tr[:11, :, :len(POINTS)] - t0          # [rank][warp][point], cycles since the CTA's own point 0
        print(f"{'point':>12} | " + " | ".join(f"rank {r:>2} min/med/max" for r in (0, 5, 10)) + " | cluster max (rank,warp)")
        for k, name in enumerate(POINTS):
            cols = []
            for r in (0, 5, 10):
                v = rel[r, :, k]
                cols.append(f"{v.min().item():6d} {v.median().item():6d} {v.max().item():6d}")
            flat = rel[:, :, k]
            idx = flat.argmax().item()
            cols.append(f"{flat.max().item():6d} ({idx // 16},{idx % 16})")
            print(f"{name:>12} | " + " | ".join(cols))
        # per-warp detail of rank 5 at the points where warps diverge
        for k in (1, 2, 4, 6, 26, 27, 28, 29, 7, 8, 9, 10, 11, 12, 22, 23):
            print(f"rank5 {POINTS[k]:>10}:", " ".join(f"{rel[5, w, k].item():5d}" for w in range(16)))
    net.set_option("recurrence_debug", 0)


if __name__ == "__main__":
    main()
