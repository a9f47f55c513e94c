"""CUDA reprojection layers (multi_view_stereonet_b200.image_predictor -> b200mvs_reproject) against the
reference-generated fixture and against the oracle on larger seeded inputs."""
import pytest
import torch

from oracle import image_predictor_oracle as ipo
from tests.test_image_predictor_oracle import close, load_fixture, masks_agree

pytestmark = pytest.mark.gpu


def cuda(f, *names):
    return [f[n].cuda() for n in names]


def test_layers_match_reference_fixture():
    from multi_view_stereonet_b200 import image_predictor as ip
    f = load_fixture()
    K, T, idepth, disparity, image = cuda(f, "K", "T", "idepth", "disparity", "image")
    assert close(ip.DisparityToIDepth()(K, T, disparity).cpu(), f["d2i"], 2e-5)
    assert close(ip.IDepthToDisparity()(K, T, idepth).cpu(), f["i2d"], 2e-5)
    px, ri, m = ip.IDepthmapProjector()(K, T, idepth)
    assert px.shape == f["proj_pixels"].shape and m.dtype == torch.bool
    assert close(px.cpu(), f["proj_pixels"], 2e-5) and close(ri.cpu(), f["proj_idepths"], 2e-5)
    assert masks_agree(m.cpu(), f["proj_mask"], f["proj_pixels"])
    for layer, arg, key in ((ip.IDepthImagePredictor(), idepth, "idip"), (ip.ImagePredictor(), disparity, "ip"),
                            (ip.RectifiedImagePredictor(), disparity, "rect")):
        pred, m = layer(K, T, arg, image)
        assert close(pred.cpu(), f[key + "_pred"], 1e-4), key
        assert int((m.cpu() != f[key + "_mask"]).sum()) == 0, key


def test_image_predictor_matches_oracle_at_full_resolution():
    from multi_view_stereonet_b200 import image_predictor as ip, synthetic
    rows, cols, n = 512, 640, 2
    _, K_pyr, Ts, right = synthetic.make_inputs(rows, cols, n, 1, smooth=True)
    K = K_pyr[0].repeat(n, 1, 1)
    T = torch.cat(Ts, 0)
    image = torch.cat([r[0] for r in right], 0)
    g = torch.Generator().manual_seed(7)
    disparity = 2.0 + 20.0 * torch.rand(n, 1, rows, cols, generator=g)
    ref_pred, ref_mask = ipo.image_predictor(K, T, disparity, image)
    truth, truth_mask = ipo.image_predictor(K, T, disparity, image, dtype=torch.float64)
    pred, mask = ip.ImagePredictor()(K.cuda(), T.cuda(), disparity.cuda(), image.cuda())
    ref_px, _, _ = ipo.idepthmap_projector(K, T, ipo.disparity_to_idepth(K, T, disparity))
    assert masks_agree(mask.cpu(), ref_mask, ref_px)
    # Two float32 evaluations of the same chain differ by ~1e-4 px in the sampling coordinate at 640 px; judge both
    # against the float64 evaluation: the kernel must be as close to it as the float32 restatement is.
    ok = ~(mask.cpu() | ref_mask | truth_mask)
    err_cuda = float(((pred.cpu().double() - truth) * ok).abs().max())
    err_ref = float(((ref_pred.double() - truth) * ok).abs().max())
    assert err_cuda <= max(4.0 * err_ref, 1e-4), (err_cuda, err_ref)
    assert float(((pred.cpu() - ref_pred) * ok).abs().mean()) <= 1e-5
    with pytest.raises(RuntimeError):
        ip.ImagePredictor()(K, T, disparity, image)                       # no CPU path
