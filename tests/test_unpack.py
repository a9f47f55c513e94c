"""Input preparation: the restated pyramid / intrinsics logic and the CUDA `multi_view_unpack_batch`
against the fixture produced by the reference's own function (tests/golden/make_golden_unpack.py)."""
import os

import numpy as np
import pytest
import torch

from multi_view_stereonet_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixture():
    z = np.load(os.path.join(HERE, "golden", "unpack_small.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_restated_pyramid_and_intrinsics_match_reference():
    """synthetic.build_image_pyramid / build_K_pyramid (what the parity inputs are made with) vs the reference."""
    f = load_fixture()
    batch = synthetic.make_raw_batch()
    pyr = synthetic.build_image_pyramid(batch["left_image"], 5)
    K_pyr = synthetic.build_K_pyramid(batch["K"].squeeze(1), [p.shape[-2:] for p in pyr])
    for lvl in range(5):
        assert torch.equal(pyr[lvl], f[f"left_image_pyr{lvl}"])
        assert float((K_pyr[lvl] - f[f"K_pyr{lvl}"]).abs().max()) <= 1e-5


@pytest.mark.gpu
def test_cuda_unpack_batch_matches_reference():
    from multi_view_stereonet_b200 import multi_view_stereonet_utils as snu
    f = load_fixture()
    batch = synthetic.make_raw_batch()
    V = len(batch["right_image"])
    inputs = snu.multi_view_unpack_batch(batch, torch.device("cuda:0"), 5)
    assert set(inputs) >= {"T_right_in_left", "T_left_in_right", "K_pyr", "left_image_pyr", "right_image_pyr", "baseline",
                           "left_depthmap_true", "left_idepthmap_true", "right_depthmap_true", "right_idepthmap_true"}
    assert float((inputs["baseline"].cpu() - f["baseline"]).abs().max()) <= 1e-6
    for lvl in range(5):
        got = inputs["left_image_pyr"][lvl].cpu()
        assert got.shape == f[f"left_image_pyr{lvl}"].shape
        assert float((got - f[f"left_image_pyr{lvl}"]).abs().max()) <= 1e-6, lvl     # odd sizes: 2x3 / 3x3 windows
        assert float((inputs["K_pyr"][lvl].cpu() - f[f"K_pyr{lvl}"]).abs().max()) <= 1e-5
        for v in range(V):
            assert float((inputs["right_image_pyr"][v][lvl].cpu() - f[f"right_image_pyr{v}_{lvl}"]).abs().max()) <= 1e-6
    for v in range(V):
        assert float((inputs["T_right_in_left"][v].cpu() - f[f"T_right_in_left{v}"]).abs().max()) <= 1e-6
        assert float((inputs["T_left_in_right"][v].cpu() - f[f"T_left_in_right{v}"]).abs().max()) <= 2e-6
        assert float((inputs["right_idepthmap_true"][v].cpu() - f[f"right_idepthmap_true{v}"]).abs().max()) <= 1e-4
    assert float((inputs["left_idepthmap_true"].cpu() - f["left_idepthmap_true"]).abs().max()) <= 1e-4
    # even sizes (2x2 windows) are bit exact against torch's area interpolation
    img = torch.rand(2, 3, 64, 80)
    pyr = snu.build_image_pyramid(img.cuda(), 5)
    for a, b in zip(pyr, synthetic.build_image_pyramid(img, 5)):
        assert torch.equal(a.cpu(), b)


@pytest.mark.gpu
def test_unpack_then_forward_is_the_reference_call_sequence():
    """test.py:188-214: unpack -> multi_view_forward -> dict with stereo_time_ms."""
    from multi_view_stereonet_b200 import MultiViewStereoNet, multi_view_stereonet_utils as snu
    from tests._util import load_gta_state
    net = MultiViewStereoNet()
    net.load_state_dict(load_gta_state(), strict=True)
    net = net.to("cuda:0").eval()
    batch = synthetic.make_raw_batch(B=1, V=1, rows=64, cols=80)
    inputs = snu.multi_view_unpack_batch(batch, torch.device("cuda:0"), net.num_levels)
    params = {"num_idepth_samples": 8, "cost_volume_filter": True, "refiners": [True] * 5}
    with torch.no_grad():
        out = snu.multi_view_forward(net, inputs, params)
    assert out["stereo_time_ms"] > 0 and len(out["left_idepthmap_pyr"]) == 5
    assert out["left_idepthmap_pyr"][0].shape == (1, 1, 64, 80) and bool(torch.isfinite(out["left_idepthmap_pyr"][0]).all())


@pytest.mark.gpu
def test_two_view_unpack_and_forward_match_reference():
    """The 2-view call sites (multi_view_stereonet_utils.py:406-539): `unpack_batch` against the fixture produced by
    the reference's own function, then `forward` with and without the right-view estimate."""
    from multi_view_stereonet_b200 import MultiViewStereoNet, multi_view_stereonet_utils as snu
    from tests._util import load_gta_state
    z = np.load(os.path.join(HERE, "golden", "unpack2_small.npz"))
    f = {k: torch.from_numpy(z[k]) for k in z.files}
    inputs = snu.unpack_batch(synthetic.make_raw_batch_two_view(), torch.device("cuda:0"), 5)
    assert torch.is_tensor(inputs["T_right_in_left"]) and torch.is_tensor(inputs["right_image_pyr"][0])
    for key, tol in (("baseline", 1e-6), ("T_right_in_left", 1e-6), ("T_left_in_right", 2e-6),
                     ("left_idepthmap_true", 1e-4), ("right_idepthmap_true", 1e-4), ("left_depthmap_true", 1e-5)):
        assert float((inputs[key].cpu() - f[key]).abs().max()) <= tol, key
    for lvl in range(5):
        assert float((inputs["K_pyr"][lvl].cpu() - f[f"K_pyr{lvl}"]).abs().max()) <= 1e-5
        assert float((inputs["left_image_pyr"][lvl].cpu() - f[f"left_image_pyr{lvl}"]).abs().max()) <= 1e-6
        assert float((inputs["right_image_pyr"][lvl].cpu() - f[f"right_image_pyr{lvl}"]).abs().max()) <= 1e-6

    net = MultiViewStereoNet()
    net.load_state_dict(load_gta_state(), strict=True)
    net = net.to("cuda:0").eval()
    b = synthetic.make_raw_batch(B=1, V=1, rows=64, cols=80)
    two = {"left_image": b["left_image"], "right_image": b["right_image"][0], "K": b["K"],
           "T_right_in_left": b["T_right_in_left"][0], "left_filename": ["l"], "right_filename": ["r"]}
    inputs = snu.unpack_batch(two, torch.device("cuda:0"), net.num_levels)
    params = {"num_idepth_samples": 8, "cost_volume_filter": True, "refiners": [True] * 5,
              "estimate_right_idepthmap": True}
    with torch.no_grad():
        out = snu.forward(net, inputs, params)
        multi = snu.multi_view_forward(net, snu.multi_view_unpack_batch(b, torch.device("cuda:0"), 5), params)
    assert set(out) >= {"left_idepthmap_pyr", "right_idepthmap_pyr", "right_idepthmap_raw_pyr",
                        "right_idepthmap_mask_pyr", "stereo_time_ms"}
    # the left estimate is the multi-view call with one comparison view
    for lvl in range(5):
        assert float((out["left_idepthmap_pyr"][lvl] - multi["left_idepthmap_pyr"][lvl]).abs().max()) <= 1e-5
    assert out["right_idepthmap_pyr"][0].shape == (1, 1, 64, 80)
    assert bool(torch.isfinite(out["right_idepthmap_pyr"][0]).all())
    params["estimate_right_idepthmap"] = False
    with torch.no_grad():
        assert "right_idepthmap_pyr" not in snu.forward(net, inputs, params)
