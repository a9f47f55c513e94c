"""A few evaluate_batch calls (512x640, batch 8) for ncu captures of depth_metrics_kernel and a device timing of it:
  ncu --set full --clock-control none -k regex:depth_metrics_kernel -s 2 -c 1 -o gpurun_out/eval python tools/eval_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_view_stereonet_b200 import evaluation as ev  # noqa: E402

B, H, W = 8, 512, 640
g = torch.Generator().manual_seed(3)
baseline = (torch.rand(B, generator=g) * 0.4 + 0.2).cuda()
truth = (torch.rand(B, 1, H, W, generator=g) * 30.0).cuda()
est = (1.0 / (truth * (0.8 + 0.4 * torch.rand(B, 1, H, W, generator=g).cuda()) + 1e-3)).contiguous()
for _ in range(4):
    ev.evaluate_batch(est, baseline, truth, "gta_sfm")
torch.cuda.synchronize()
tick, tock = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
import ctypes  # noqa: E402
from multi_view_stereonet_b200 import _lib  # noqa: E402
lib = _lib.load()
idepth, depth = torch.empty_like(est), torch.empty_like(est)
metrics = torch.empty((B, 8), dtype=torch.float64, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ms = []
for _ in range(10):
    flush.zero_()
    tick.record()
    lib.b200mvs_depth_metrics(est.data_ptr(), baseline.data_ptr(), truth.data_ptr(), 0, 0.0, 1e3, B, H * W,
                              idepth.data_ptr(), depth.data_ptr(), metrics.data_ptr(), stream)
    tock.record()
    torch.cuda.synchronize()
    ms.append(tick.elapsed_time(tock))
ms.sort()
byt = B * H * W * 4 * 4          # estimate + truth in, idepth + depth out
print(f"b200mvs_depth_metrics: {B}x{H}x{W}, L2 flushed, median {ms[len(ms) // 2] * 1e3:.1f} us (scratch alloc + memset + kernel), "
      f"{byt / ms[len(ms) // 2] / 1e6:.0f} GB/s algorithmic ({byt / 1e6:.1f} MB)")
