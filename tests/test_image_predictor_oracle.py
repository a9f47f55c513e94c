"""The reprojection-layer oracle against the fixture produced by the reference's own classes
(tests/golden/make_golden_image_predictor.py)."""
import os

import numpy as np
import torch

from oracle import image_predictor_oracle as ipo

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixture():
    z = np.load(os.path.join(HERE, "golden", "image_predictor_small.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def close(a, b, tol):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def masks_agree(mask, ref_mask, pixels):
    """Masks are hard thresholds on |coordinate| > 1: a disagreement is only acceptable on the threshold itself."""
    bad = mask != ref_mask
    if not bool(bad.any()):
        return True
    margin = (pixels.abs() - 1.0).abs().amin(dim=-1).unsqueeze(1)
    return bool((margin[bad] < 1e-5).all())


def test_disparity_to_idepth_matches_reference():
    f = load_fixture()
    assert close(ipo.disparity_to_idepth(f["K"], f["T"], f["disparity"]), f["d2i"], 2e-5)


def test_idepth_to_disparity_matches_reference():
    f = load_fixture()
    assert close(ipo.idepth_to_disparity(f["K"], f["T"], f["idepth"]), f["i2d"], 2e-5)


def test_idepthmap_projector_matches_reference():
    f = load_fixture()
    px, ri, m = ipo.idepthmap_projector(f["K"], f["T"], f["idepth"])
    assert close(px, f["proj_pixels"], 2e-5)
    assert close(ri, f["proj_idepths"], 2e-5)
    assert masks_agree(m, f["proj_mask"], f["proj_pixels"])
    assert 0.02 < float(f["proj_mask"].float().mean()) < 0.6     # the fixture exercises both sides of the mask


def test_image_predictors_match_reference():
    f = load_fixture()
    pred, m = ipo.idepth_image_predictor(f["K"], f["T"], f["idepth"], f["image"])
    assert close(pred, f["idip_pred"], 1e-4) and int((m != f["idip_mask"]).sum()) == 0
    pred, m = ipo.image_predictor(f["K"], f["T"], f["disparity"], f["image"])
    assert close(pred, f["ip_pred"], 1e-4) and int((m != f["ip_mask"]).sum()) == 0
    pred, m = ipo.rectified_image_predictor(f["K"], f["T"], f["disparity"], f["image"])
    assert close(pred, f["rect_pred"], 1e-4) and int((m != f["rect_mask"]).sum()) == 0
