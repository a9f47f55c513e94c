"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the
reference-generated golden fixtures.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from multi_view_stereonet_b200 import synthetic
from tests._util import REL_LINF_TOL, load_case, rel_linf, unpack_mask

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(gta_state):
    from tests._gpu_util import make_net
    return make_net(gta_state)


def _assert_report(rep):
    from tests._gpu_util import format_report
    txt = format_report(rep)
    # any mask disagreement with the oracle must sit on a knife-edge coordinate
    assert rep["mask_flips_not_knife_edge"] == 0, txt
    for k, v in rep.items():
        if k.startswith(("idepth", "raw")):
            assert v <= REL_LINF_TOL, f"{k}\n{txt}"
        if k.startswith("mask_mismatch"):
            assert v == 0, f"{k}\n{txt}"


@pytest.mark.parametrize("name", ["cfg1", "cfg1_smooth", "mv_small", "odd_small", "flags_nocvf", "cfg2", "cfg2_smooth"])
def test_golden_fixture(name, net, gta_state):
    """CUDA path vs outputs of the reference's own model (tests/golden)."""
    z, inputs, hyps, cvf, refiners = load_case(name)
    net.keep_stages(False)
    with torch.no_grad():
        out = net(*synthetic.to_device(inputs, "cuda"), hyps, cvf, refiners)
    for lvl in range(5):
        got = out["left_idepthmap_pyr"][lvl].cpu()
        assert rel_linf(got, z[f"idepth{lvl}"]) <= REL_LINF_TOL, (name, lvl)
        m = out["left_idepthmap_mask_pyr"][lvl].cpu().numpy()
        assert m.dtype == np.bool_
        batch = m.shape[0]
        np.testing.assert_array_equal(m.reshape(batch, hyps, -1).sum(-1), z[f"mask_count{lvl}"])
        if f"mask{lvl}" in z:
            np.testing.assert_array_equal(m, unpack_mask(z[f"mask{lvl}"], m.shape))
        if f"raw{lvl}" in z:
            assert rel_linf(out["left_idepthmap_raw_pyr"][lvl].cpu(), z[f"raw{lvl}"]) <= REL_LINF_TOL, (name, lvl)


@pytest.mark.parametrize("rows,cols,views,hyps,batch,smooth", [
    (64, 80, 1, 8, 1, False),          # cfg1
    (96, 128, 2, 6, 2, True),
    (68, 90, 3, 5, 1, True),           # odd pyramid sizes
    (256, 320, 2, 16, 2, False),
    (512, 640, 1, 64, 1, False),       # cfg2
    (512, 640, 1, 64, 1, True),
])
def test_stagewise_vs_oracle(rows, cols, views, hyps, batch, smooth, net, gta_state):
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(rows, cols, views, batch, smooth=smooth)
    rep, _, _ = run_case(net, gta_state, inputs, hyps)
    _assert_report(rep)
    assert rep["v0/right_feature_volume"] <= REL_LINF_TOL
    assert rep["left_feature4"] <= 1e-4


def test_flag_variants(net, gta_state):
    """do_cost_volume_filter=False and partially disabled refiners, including the
    reference's double baseline division when do_refiners[4] is False."""
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(64, 80, 2, 1, smooth=True)
    for cvf, refiners in [(False, (True, False, True, False, False)), (True, (False,) * 5), (False, (True,) * 5)]:
        rep, _, _ = run_case(net, gta_state, inputs, 8, cvf, refiners, stages=False)
        _assert_report(rep)


def test_cfg3_item_multiview(net, gta_state):
    """One image group of BASELINE cfg3 (4 comparison views, 64 hypotheses) against the oracle, and a batch of
    two groups: items are independent, so item 0 of the batch must reproduce the batch-1 run of the same item.
    (Work decomposition and atomics order depend on the batch size, so the match is to float32 noise -- which the
    63-step recurrence amplifies to ~1e-4 on these inputs -- not bit-exact.)"""
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(512, 640, 4, 1)
    rep, _, _ = run_case(net, gta_state, inputs, 64, stages=False)
    _assert_report(rep)
    one = synthetic.make_inputs(512, 640, 4, 1, smooth=True)
    both = synthetic.make_inputs(512, 640, 4, 2, smooth=True)
    with torch.no_grad():
        out1 = net(*synthetic.to_device(one, "cuda"), 64, True, [True] * 5)
        out1b = net(*synthetic.to_device(one, "cuda"), 64, True, [True] * 5)
        out2 = net(*synthetic.to_device(both, "cuda"), 64, True, [True] * 5)
    for lvl in range(5):
        a = out1["left_idepthmap_pyr"][lvl][0]
        b = out2["left_idepthmap_pyr"][lvl][0]
        # run-to-run: only the order of the float64 statistic atomics differs
        assert rel_linf(out1b["left_idepthmap_pyr"][lvl][0].cpu(), a.cpu()) <= 1e-5
        assert rel_linf(b.cpu(), a.cpu()) <= REL_LINF_TOL / 2
        assert bool((out1["left_idepthmap_mask_pyr"][lvl][0] == out2["left_idepthmap_mask_pyr"][lvl][0]).all())


def test_cfg5_item_large_image(net, gta_state):
    """One image group of BASELINE cfg5 (1024x1280, 4 comparison views, 128 hypotheses): the 1/16-scale image
    (64x80 = 41 M-tiles) is beyond the persistent recurrence kernel's cluster, so this exercises the multi-launch
    fallback of the depth sweep, and the full-resolution warp has a knife-edge mask pixel on these inputs
    (tests/_gpu_util.py re-runs the oracle with the CUDA path's tie-break after checking it is one)."""
    from tests._gpu_util import run_case
    rep, _, _ = run_case(net, gta_state, synthetic.make_inputs(1024, 1280, 4, 1), 128, stages=False)
    _assert_report(rep)


def test_homography_image_predictor(net):
    from multi_view_stereonet_b200 import HomographyImagePredictor
    from oracle import mvsnet_oracle as oracle
    g = torch.Generator().manual_seed(5)
    img = torch.rand(3, 5, 37, 53, generator=g)
    H = torch.eye(3).repeat(3, 1, 1) + 0.01 * torch.randn(3, 3, 3, generator=g)
    H[:, 0, 2] += torch.tensor([3.3, -7.1, 0.4])
    H[:, 2, :2] *= 0.01
    ref, rmask = oracle.homography_warp(H, img)
    pred, mask = HomographyImagePredictor()(H.cuda(), img.cuda())
    assert mask.dtype == torch.bool and tuple(mask.shape) == (3, 1, 37, 53)
    assert int((mask.cpu() != rmask).sum()) == 0
    assert float((pred.cpu() - ref).abs().max()) <= 1e-5


def test_host_entry_matches_device_entry(net, gta_state):
    """forward() on CPU tensors goes through b200mvs_forward_host."""
    inputs = synthetic.make_inputs(64, 80, 1, 1)
    with torch.no_grad():
        a = net(*inputs, 8, True, [True] * 5)
        b = net(*synthetic.to_device(inputs, "cuda"), 8, True, [True] * 5)
    assert a["left_idepthmap_pyr"][0].device.type == "cpu"
    for lvl in range(5):
        assert rel_linf(a["left_idepthmap_pyr"][lvl], b["left_idepthmap_pyr"][lvl].cpu()) <= 1e-5
        assert bool((a["left_idepthmap_mask_pyr"][lvl] == b["left_idepthmap_mask_pyr"][lvl].cpu()).all())
    assert net.last_h2d_bytes > 0 and net.last_d2h_bytes > 0


def test_inputs_not_modified_and_errors(net):
    inputs = synthetic.to_device(synthetic.make_inputs(64, 80, 1, 1), "cuda")
    T_before = inputs[2][0].clone()
    with torch.no_grad():
        net(*inputs, 8, True, [True] * 5)
    assert torch.equal(T_before, inputs[2][0])          # the reference clones T (multi_view_stereonet.py:566)
    with pytest.raises(AssertionError):
        net(inputs[0][:4], inputs[1], inputs[2], inputs[3], 8, True, [True] * 5)   # :548-549
