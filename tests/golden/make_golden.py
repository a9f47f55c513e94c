"""Generates the golden fixtures in this directory by running the REFERENCE's own
eager model (imported read-only from /root/reference) on the seeded synthetic
inputs of `multi_view_stereonet_b200.synthetic`.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes
  gta_sfm_150epochs_state.npz   the reference's pretrained GTA-SfM weights, extracted
                                from its TorchScript archive (608,614 parameters)
  <case>.npz                    reference outputs (+ stage tensors for small cases)
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from multi_view_stereonet_b200 import synthetic, weights  # noqa: E402
from multi_view_stereonet.multi_view_stereonet import (  # noqa: E402  (the reference)
    MultiViewStereoNet, create_idepth_samples)

# name -> (rows, cols, views, hyps, batch, smooth, do_cvf, do_refiners, with_stages)
CASES = {
    "cfg1": (64, 80, 1, 8, 1, False, True, [True] * 5, True),
    "cfg1_smooth": (64, 80, 1, 8, 1, True, True, [True] * 5, True),
    "mv_small": (96, 128, 2, 6, 2, True, True, [True] * 5, True),
    "odd_small": (68, 90, 3, 5, 1, True, True, [True] * 5, False),
    "flags_nocvf": (64, 80, 2, 8, 1, True, False, [True, False, True, False, False], False),
    "cfg2": (512, 640, 1, 64, 1, False, True, [True] * 5, False),
    "cfg2_smooth": (512, 640, 1, 64, 1, True, True, [True] * 5, False),
    # one image group of BASELINE cfg3: 4 comparison views, 64 hypotheses
    "cfg3_item": (512, 640, 4, 64, 1, False, True, [True] * 5, False),
    # cfg2 at SURVEY.md 8d's original camera pitch (0.01 rad): the geometry with knife-edge mask pixels
    "cfg2_pitch001": (512, 640, 1, 64, 1, False, True, [True] * 5, False),
}
PITCH = {"cfg2_pitch001": 0.01}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    pt = os.path.join(REF, "pretrained/gta_sfm_150epochs/checkpoints/epoch0149/stereo_network.pt")
    sd = weights.load_torchscript_archive_weights(pt)
    net = MultiViewStereoNet().eval()
    net.load_state_dict(sd, strict=True)
    unique = {k: v for k, v in sd.items() if not k.startswith("right_feature_extractor.feature_extractor.")}
    assert sum(v.numel() for v in unique.values()) == 608614       # pretrained/gta_sfm_150epochs/logs.txt:4
    weights.save_state_npz(unique, os.path.join(HERE, "gta_sfm_150epochs_state.npz"))

    only = sys.argv[1:]
    for name, (rows, cols, views, hyps, batch, smooth, cvf, refiners, with_stages) in CASES.items():
        if only and name not in only:
            continue
        inp = synthetic.make_inputs(rows, cols, views, batch, smooth=smooth, pitch=PITCH.get(name))
        left_pyr, K_pyr, Ts, right_pyrs = inp
        with torch.no_grad():
            out = net(left_pyr, K_pyr, Ts, right_pyrs, hyps, cvf, refiners)
        big = rows * cols > 100000
        rec = {"meta": np.array([rows, cols, views, hyps, batch, int(smooth), int(cvf)] + [int(r) for r in refiners]),
               "input_checksum": np.array([float(left_pyr[0].double().sum()), float(right_pyrs[-1][0].double().sum()),
                                           float(left_pyr[4].double().abs().sum())]),
               "pitch": np.array([PITCH.get(name, synthetic.PITCH_RAD)])}
        for lvl in range(5):
            rec[f"idepth{lvl}"] = out["left_idepthmap_pyr"][lvl].numpy()
            m = out["left_idepthmap_mask_pyr"][lvl].numpy()
            rec[f"mask_count{lvl}"] = m.reshape(batch, hyps, -1).sum(-1)
            if not big or lvl == 4:
                rec[f"raw{lvl}"] = out["left_idepthmap_raw_pyr"][lvl].numpy()
                rec[f"mask{lvl}"] = np.packbits(m.reshape(-1))
        if with_stages:
            with torch.no_grad():
                lf = net.left_feature_extractor(left_pyr[0])
                for lvl in range(1, 5):
                    rec[f"left_feature{lvl}"] = lf[lvl].numpy()
                T = Ts[0].clone()
                baseline = T[:, :3, 3].pow(2).sum(1).sqrt()
                T[:, :3, 3] /= baseline.unsqueeze(1)
                samples = create_idepth_samples(T, K_pyr[-1], left_pyr[-1].shape[-2], left_pyr[-1].shape[-1], hyps)
                vol, mask = net.right_feature_extractor(T, K_pyr, right_pyrs[0], samples)
                cost = (lf[-1].unsqueeze(2) - vol).abs() * (~mask).unsqueeze(1)
                filt = net.volume_filter4(cost)
                rec["v0_idepth_samples"] = samples.numpy()
                rec["v0_right_feature_volume"] = vol.numpy()
                rec["v0_mask"] = np.packbits(mask.numpy().reshape(-1))
                rec["v0_cost_filtered"] = filt.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        o = out["left_idepthmap_pyr"][0]
        print(f"{name}: idepth0 range [{float(o.min()):.4f}, {float(o.max()):.4f}]  "
              f"{os.path.getsize(os.path.join(HERE, name + '.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
