"""Golden fixture for the reprojection layers: runs the REFERENCE's own classes (imported read-only from
/root/reference/stereo/image_predictor.py) on seeded inputs and stores inputs + outputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_image_predictor.py        -> tests/golden/image_predictor_small.npz
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from multi_view_stereonet_b200 import synthetic  # noqa: E402
from stereo import image_predictor as ref  # noqa: E402  (the reference)


def make_inputs(n=3, rows=24, cols=32, channels=3, seed=4321):
    """K / poses of the synthetic camera rig, smooth positive idepth and disparity maps, a smooth right image."""
    inp = synthetic.make_inputs(rows * 16, cols * 16, n, 1, seed=seed, smooth=True)
    _, K_pyr, Ts, _ = inp
    K = K_pyr[4].repeat(n, 1, 1).contiguous()                 # intrinsics of the (rows, cols) level
    T = torch.cat(Ts, dim=0).contiguous()                     # n different relative poses
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, rows), torch.linspace(0, 1, cols), indexing="ij")
    idepth = torch.stack([0.008 + 0.04 * (0.5 + 0.5 * torch.sin(3.0 * xx + i) * torch.cos(2.0 * yy - i)) for i in range(n)])
    idepth = (idepth + 0.002 * torch.rand(idepth.shape, generator=g)).unsqueeze(1).contiguous()
    disparity = torch.stack([1.0 + 3.0 * (0.5 + 0.5 * torch.cos(2.5 * xx - i) * torch.sin(1.5 * yy + i)) for i in range(n)])
    disparity = (disparity + 0.05 * torch.rand(disparity.shape, generator=g)).unsqueeze(1).contiguous()
    image = torch.stack([torch.stack([torch.sin(6.0 * xx * (c + 1) + i) * torch.cos(5.0 * yy + c) for c in range(channels)])
                         for i in range(n)])
    image = (image + 0.1 * torch.rand(image.shape, generator=g)).contiguous()
    return K, T, idepth, disparity, image


def main():
    K, T, idepth, disparity, image = make_inputs()
    out = {"K": K, "T": T, "idepth": idepth, "disparity": disparity, "image": image}
    with torch.no_grad():
        out["d2i"] = ref.DisparityToIDepth()(K, T, disparity.clone())
        out["i2d"] = ref.IDepthToDisparity()(K, T, idepth.clone())
        px, ri, m = ref.IDepthmapProjector()(K, T, idepth.clone())
        out["proj_pixels"], out["proj_idepths"], out["proj_mask"] = px, ri, m
        out["idip_pred"], out["idip_mask"] = ref.IDepthImagePredictor()(K, T, idepth.clone(), image.clone())
        out["ip_pred"], out["ip_mask"] = ref.ImagePredictor()(K, T, disparity.clone(), image.clone())
        out["rect_pred"], out["rect_mask"] = ref.RectifiedImagePredictor()(K, T, disparity.clone(), image.clone())
    path = os.path.join(HERE, "image_predictor_small.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    for k, v in out.items():
        print(k, tuple(v.shape), v.dtype, float(v.float().abs().max()))
    print("masked fraction:", {k: float(out[k].float().mean()) for k in ("proj_mask", "idip_mask", "ip_mask", "rect_mask")})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
