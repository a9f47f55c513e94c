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


@pytest.mark.parametrize("name", ["cfg3_item", "cfg2_pitch001"])
def test_reference_fixture_with_knife_edge_geometry(name, net, gta_state):
    """Reference-generated fixtures whose geometry is NOT tuned away from the mask threshold: one image group of
    BASELINE cfg3 (4 comparison views) and cfg2 at SURVEY.md 8d's original camera pitch (0.01 rad).  Protocol
    ("Knife-edge mask pixels"): every mask disagreement with the oracle must sit within float32
    resolution of the threshold, the oracle re-run with the CUDA path's tie-breaks must match to the parity bar --
    and when no tie broke differently the CUDA path must match the reference's own output directly."""
    from tests._gpu_util import run_case
    z, inputs, hyps, cvf, refiners = load_case(name)
    rep, out, _ = run_case(net, gta_state, inputs, hyps, cvf, tuple(refiners), stages=False)
    _assert_report(rep)
    print(name, "mask flips vs oracle: level 0", rep["mask_flips_l0"], "level 4", rep["mask_flips_l4"])
    if rep["mask_flips_l0"] == 0 and rep["mask_flips_l4"] == 0:
        for lvl in range(5):
            assert rel_linf(out["left_idepthmap_pyr"][lvl].cpu(), z[f"idepth{lvl}"]) <= REL_LINF_TOL, (name, lvl)
            m = out["left_idepthmap_mask_pyr"][lvl].cpu().numpy()
            np.testing.assert_array_equal(m.reshape(m.shape[0], hyps, -1).sum(-1), z[f"mask_count{lvl}"])


def test_batch8_matches_single_items(net, gta_state):
    """The per-GPU share of BASELINE cfg4 (batch 8, one comparison view, 64 hypotheses): image groups are independent
    (multi_view_stereonet.py never mixes batch items), so every item of the batch must reproduce its own batch-1
    run -- which test_golden_fixture / test_stagewise_vs_oracle pin to the reference for item 0 -- and item 5 is
    checked against the oracle directly."""
    from oracle import mvsnet_oracle as oracle
    batch = synthetic.make_inputs(512, 640, 1, 8)
    net.keep_stages(False)
    with torch.no_grad():
        out8 = net(*synthetic.to_device(batch, "cuda"), 64, True, [True] * 5)
        for i in (0, 3, 7):
            one = synthetic.make_inputs(512, 640, 1, 1, first_item=i)
            out1 = net(*synthetic.to_device(one, "cuda"), 64, True, [True] * 5)
            for lvl in range(5):
                assert rel_linf(out8["left_idepthmap_pyr"][lvl][i].cpu(), out1["left_idepthmap_pyr"][lvl][0].cpu()) \
                    <= REL_LINF_TOL / 2, (i, lvl)
                assert bool((out8["left_idepthmap_mask_pyr"][lvl][i] == out1["left_idepthmap_mask_pyr"][lvl][0]).all())
        one = synthetic.make_inputs(512, 640, 1, 1, first_item=5)
        ref = oracle.forward(gta_state, *one, 64, True, (True,) * 5)
    for lvl in range(5):
        assert rel_linf(out8["left_idepthmap_pyr"][lvl][5].cpu(), ref["left_idepthmap_pyr"][lvl][0]) <= REL_LINF_TOL, lvl
        assert int((out8["left_idepthmap_mask_pyr"][lvl][5].cpu() != ref["left_idepthmap_mask_pyr"][lvl][0]).sum()) == 0


def test_cost_filter_decompositions_agree(net):
    """The Conv3d filter (cvf_tc.cu) cuts its work differently at batch 1 (every (row tile, strip) column in 13 depth
    chunks of 5 slices: accumulator windows never wrap around the tensor-memory ring) and at batch 8 (equal ranges of 39
    output slices that cross column boundaries: several segments per CTA, wrapped windows, the N = 64 + 32 path).  Every
    output slice adds its products in the same order either way, so the filtered cost volumes agree to the float64
    GroupNorm statistics' summation order."""
    batch = synthetic.make_inputs(512, 640, 1, 8)
    net.keep_stages(True)
    try:
        with torch.no_grad():
            net(*synthetic.to_device(batch, "cuda"), 64, True, [True] * 5)
            cost8 = net.get_stage("cost_filtered", torch.float32).view(8, 64, 32, 40).clone()
            for i in (0, 6):
                one = synthetic.make_inputs(512, 640, 1, 1, first_item=i)
                net(*synthetic.to_device(one, "cuda"), 64, True, [True] * 5)
                cost1 = net.get_stage("cost_filtered", torch.float32).view(64, 32, 40)
                assert rel_linf(cost8[i].cpu(), cost1.cpu()) <= 1e-5, i
    finally:
        net.keep_stages(False)


def test_many_chains_take_the_wide_sweep(net):
    """From the third round of clusters on (more than 14 (image group, view) chains on B200) the automatic rule hands
    the depth sweep to the wide kernel (one co-resident grid for all chains).  Batch 8 with two views = 16 chains:
    same results as the cluster kernel (option sweep = 0 with lanes = 1 and a forced "sweep" = 2 step-by-step run are
    the two references), and a forced two-lane call never starts two cooperative grids (it falls back to clusters)."""
    inputs = synthetic.to_device(synthetic.make_inputs(512, 640, 2, 8), "cuda")
    try:
        with torch.no_grad():
            auto = net(*inputs, 16, True, [True] * 5)
            net.set_option("sweep", 2)
            steps = net(*inputs, 16, True, [True] * 5)
            n_steps = net.last_launch_count()
            net.set_option("sweep", 0)
            net.set_option("lanes", 2)
            lanes = net(*inputs, 16, True, [True] * 5)
        assert n_steps > net.last_launch_count() / 2 + 15 * 3      # 15 steps x (warp + three convs) instead of one launch
        for lvl in range(5):
            for other in (steps, lanes):
                assert rel_linf(auto["left_idepthmap_pyr"][lvl].cpu(), other["left_idepthmap_pyr"][lvl].cpu()) <= REL_LINF_TOL / 2
                assert bool((auto["left_idepthmap_mask_pyr"][lvl] == other["left_idepthmap_mask_pyr"][lvl]).all())
    finally:
        net.set_option("sweep", 0)
        net.set_option("lanes", 0)


def test_lanes_option_gives_the_same_results(net):
    """Option "lanes" = 2 cuts a call whose depth sweep needs more than one round of clusters into two concurrent lanes
    of whole image groups (1 = never; 0 = automatic, the default: two lanes when the last round would hold one or two
    clusters, e.g. batch 8 with one view); results must not change."""
    inputs = synthetic.to_device(synthetic.make_inputs(512, 640, 1, 8), "cuda")
    try:
        with torch.no_grad():
            net.set_option("lanes", 1)
            one = net(*inputs, 64, True, [True] * 5)
            n1 = net.last_launch_count()
            net.set_option("lanes", 2)
            two = net(*inputs, 64, True, [True] * 5)
            n2 = net.last_launch_count()
            net.set_option("lanes", 0)
            auto = net(*inputs, 64, True, [True] * 5)
            assert net.last_launch_count() == n2        # 8 pairs, 7 co-resident clusters: the automatic rule splits
            assert rel_linf(auto["left_idepthmap_pyr"][0].cpu(), two["left_idepthmap_pyr"][0].cpu()) <= REL_LINF_TOL / 2
    finally:
        net.set_option("lanes", 0)
    assert n2 == 2 * n1
    for lvl in range(5):
        assert rel_linf(two["left_idepthmap_pyr"][lvl].cpu(), one["left_idepthmap_pyr"][lvl].cpu()) <= REL_LINF_TOL / 2
        assert bool((two["left_idepthmap_mask_pyr"][lvl] == one["left_idepthmap_mask_pyr"][lvl]).all())


def test_lazy_and_packed_mask_volumes(net):
    """mask_mode "lazy": forward produces the level-4 volume only; the finer levels and their bit-packed form come
    from the same MaskUpsampler chain on demand and equal the dense default bit for bit."""
    from multi_view_stereonet_b200.multi_view_stereonet import LazyMaskPyramid
    inputs = synthetic.to_device(synthetic.make_inputs(96, 136, 2, 2, smooth=True), "cuda")   # W/8 not integral at L2+
    try:
        with torch.no_grad():
            dense = net(*inputs, 6, True, [True] * 5)
            n_dense = net.last_launch_count()
            net.mask_mode = "lazy"
            lazy = net(*inputs, 6, True, [True] * 5)
            n_lazy = net.last_launch_count()
            host = net(*[[t.cpu() for t in inputs[0]], [t.cpu() for t in inputs[1]], [t.cpu() for t in inputs[2]],
                         [[t.cpu() for t in p] for p in inputs[3]]], 6, True, [True] * 5)
            net.mask_mode = "none"
            none = net(*inputs, 6, True, [True] * 5)
    finally:
        net.mask_mode = "dense"
    assert n_lazy == n_dense - 4                      # the four mask-upsampling launches are gone from forward
    pyr = lazy["left_idepthmap_mask_pyr"]
    assert isinstance(pyr, LazyMaskPyramid) and len(pyr) == 5
    assert none["left_idepthmap_mask_pyr"][:4] == [None] * 4
    for lvl in range(5):
        assert torch.equal(lazy["left_idepthmap_pyr"][lvl], dense["left_idepthmap_pyr"][lvl])
    for lvl in (4, 2, 0, 1, 3):                       # any order: the chain is built from the finest level present
        want = dense["left_idepthmap_mask_pyr"][lvl]
        got = pyr[lvl]
        assert got.dtype == torch.bool and torch.equal(got, want), lvl
        assert torch.equal(host["left_idepthmap_mask_pyr"][lvl], want.cpu()), lvl
    fresh = LazyMaskPyramid(dense["left_idepthmap_mask_pyr"][4].view(torch.uint8), [m.shape[-2:] for m in pyr])
    for lvl in (0, 3, 4):                             # packed straight from the coarser level (no dense volume of lvl)
        bits = fresh.packed(lvl).cpu().numpy()
        want = np.packbits(dense["left_idepthmap_mask_pyr"][lvl].cpu().numpy(), axis=-1)
        np.testing.assert_array_equal(bits, want)


@pytest.mark.parametrize("h,w,H,W", [(32, 40, 64, 80), (8, 16, 16, 32), (64, 80, 128, 160), (9, 13, 17, 25), (32, 40, 63, 80),
                                     (6, 12, 12, 24)])
def test_mask_upsampling_against_interpolate(h, w, H, W):
    """MaskUpsampler (multi_view_stereonet.py:389-396: float -> bilinear -> > 0.5) on random 0/1 volumes against
    torch's own interpolate: the exact-2x sizes take the byte-replication kernel (upsample_mask2x_kernel: for 0/1
    inputs the thresholded bilinear value IS the nearest input pixel), the others the general kernel."""
    from multi_view_stereonet_b200.multi_view_stereonet import LazyMaskPyramid
    g = torch.Generator().manual_seed(h * 1000 + W)
    for density in (0.5, 0.1, 0.9):
        m = (torch.rand(3, 5, h, w, generator=g) < density).to(torch.uint8).cuda()
        pyr = LazyMaskPyramid(m, [(H, W)] * 4 + [(h, w)])
        got = pyr._upsample(m, 3, False)
        val = torch.nn.functional.interpolate(m.float().cpu(), size=(H, W), mode="bilinear", align_corners=False)
        diff = got.bool().cpu() != (val > 0.5)
        if H == 2 * h and W == 2 * w:
            assert int(diff.sum()) == 0, (h, w, H, W, density)      # dyadic weights: no rounding anywhere
        else:   # general sizes: only values that sit on the threshold to float32 rounding may differ
            assert float((val[diff] - 0.5).abs().max()) <= 1e-6 if bool(diff.any()) else True, (h, w, H, W, density)


def test_weights_follow_in_place_updates(gta_state):
    """The native weight copy follows in-place parameter updates no module hook sees: a submodule's
    load_state_dict, torch.nn.init, an optimizer step."""
    from tests._gpu_util import make_net
    net = make_net(gta_state)
    inputs = synthetic.to_device(synthetic.make_inputs(64, 80, 1, 1, smooth=True), "cuda")
    run = lambda: net(*inputs, 8, True, [True] * 5)["left_idepthmap_pyr"][0].clone()
    with torch.no_grad():
        base = run()
        sub = {k: v.clone() for k, v in net.refiner0.state_dict().items()}
        sub["conv_final.bias"] += 0.05
        net.refiner0.load_state_dict(sub)
        a = run()
        assert not torch.equal(a, base)
        torch.nn.init.constant_(net.refiner0.conv_final.bias, float(gta_state["refiner0.conv_final.bias"]))
        assert torch.equal(run(), base)
    opt = torch.optim.SGD([net.refiner0.conv_final.bias], lr=1.0)
    net.refiner0.conv_final.bias.grad = torch.full_like(net.refiner0.conv_final.bias, -0.05)
    opt.step()
    with torch.no_grad():
        assert torch.equal(run(), a)
    with pytest.raises(RuntimeError, match="inference-only"):
        net([t.clone().requires_grad_(True) for t in inputs[0]], *inputs[1:], 8, True, [True] * 5)


@pytest.mark.parametrize("rows,cols,views,hyps,batch,smooth", [
    (64, 80, 1, 8, 1, False),          # cfg1
    (96, 128, 2, 6, 2, True),
    (68, 90, 3, 5, 1, True),           # odd pyramid sizes
    (256, 320, 2, 16, 2, False),
    (512, 640, 1, 64, 1, False),       # cfg2
    (512, 640, 1, 64, 1, True),
    (500, 636, 1, 64, 1, True),        # rows not a multiple of the dilations: ragged polyphase components at level 0
    (512, 640, 1, 12, 1, True),        # the reference's own 12 hypotheses: small idepth range, split-fp16 refiners
])
def test_stagewise_vs_oracle(rows, cols, views, hyps, batch, smooth, net, gta_state):
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(rows, cols, views, batch, smooth=smooth)
    rep, _, _ = run_case(net, gta_state, inputs, hyps)
    _assert_report(rep)
    assert rep["v0/right_feature_volume"] <= REL_LINF_TOL
    assert rep["left_feature4"] <= 1e-4


@pytest.mark.parametrize("rows,cols,hyps", [(64, 80, 2), (96, 128, 3), (256, 320, 2)])
def test_minimal_hypothesis_counts(rows, cols, hyps, net, gta_state):
    """Two hypotheses are the smallest sweep the reference supports (one step: no hand-off between CTAs ever happens,
    the last-step paths of the persistent kernel run first); three add exactly one hand-off."""
    from tests._gpu_util import run_case
    rep, _, _ = run_case(net, gta_state, synthetic.make_inputs(rows, cols, 1, 1, smooth=True), hyps, stages=False)
    _assert_report(rep)


@pytest.mark.parametrize("rows,cols,views,hyps,batch", [
    (512, 640, 1, 16, 1),    # 11 CTAs of one tile each
    (512, 640, 2, 8, 3),     # 6 chains
    (96, 128, 1, 5, 1),      # a single CTA
    (272, 400, 2, 6, 2),     # odd 1/16-scale image (17 x 25)
])
def test_wide_sweep_vs_oracle(rows, cols, views, hyps, batch, net, gta_state):
    """The wide sweep (sweep_wide.cu: one cooperative launch, chain barriers through L2) forced onto shapes the
    cluster kernel normally takes: same parity bar, and its feature volume agrees with the cluster kernel's far
    inside that bar (same split-fp16 arithmetic; only the order of the GroupNorm partial sums differs)."""
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(rows, cols, views, batch, smooth=True)
    net.keep_stages(True)
    with torch.no_grad():
        net(*synthetic.to_device(inputs, torch.device("cuda")), hyps, True, [True] * 5)
        torch.cuda.synchronize()
        vol_cluster = net.get_stage("right_feature_volume", torch.float32).clone()
    net.set_option("sweep", 1)
    try:
        rep, _, _ = run_case(net, gta_state, inputs, hyps, stages=True)
        vol_wide = net.get_stage("right_feature_volume", torch.float32)
        scale = float(vol_cluster.abs().max())
        assert float((vol_wide - vol_cluster).abs().max()) <= 2e-4 * scale
    finally:
        net.set_option("sweep", 0)
        net.keep_stages(False)
    _assert_report(rep)
    for v in range(views):
        assert rep[f"v{v}/right_feature_volume"] <= 1e-3


def test_wide_sweep_watchdog_reports_a_stuck_barrier(gta_state):
    """The wide sweep's chain barriers spin; a wait that cannot complete must end in an error, not in a hung GPU.
    Debug bit 32 makes one CTA skip an arrival: after ~1 s the watchdog raises the abort flag, the kernel ends, and the
    host entry reports it; the next forward (bit cleared) is fine again."""
    from tests._gpu_util import make_net
    net = make_net(gta_state)
    inputs = synthetic.make_inputs(512, 640, 1, 1, smooth=True)
    with torch.no_grad():
        good = net(*synthetic.to_device(inputs, "cuda"), 4, True, [True] * 5)["left_idepthmap_pyr"][0].cpu()
        net.set_option("sweep", 1)
        net.set_option("recurrence_debug", 32)
        with pytest.raises(RuntimeError, match="timed out"):
            net(*inputs, 4, True, [True] * 5)          # CPU tensors: b200mvs_forward_host synchronises and checks
        net.set_option("recurrence_debug", 0)
        again = net(*inputs, 4, True, [True] * 5)["left_idepthmap_pyr"][0]
    assert rel_linf(again, good) <= REL_LINF_TOL / 2


def test_flag_variants(net, gta_state):
    """do_cost_volume_filter=False and partially disabled refiners, including the
    reference's double baseline division when do_refiners[4] is False."""
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(64, 80, 2, 1, smooth=True)
    for cvf, refiners in [(False, (True, False, True, False, False)), (True, (False,) * 5), (False, (True,) * 5)]:
        rep, _, _ = run_case(net, gta_state, inputs, 8, cvf, refiners, stages=False)
        _assert_report(rep)


def test_large_incremental_motion_leaves_the_sweep_window(net, gta_state):
    """Forward + vertical camera motion: between consecutive hypotheses the corner pixels of the 1/16-scale image move
    by more than one row, so some taps of the depth sweep's warp lie outside the CTA's shared-memory window
    (recurrence.cu).  gather_plan_kernel must raise the per-(group, view) flag and the sweep must read those taps
    from global memory behind the progress flags; the ordinary geometry must not raise it."""
    from tests._gpu_util import run_case
    inputs = synthetic.make_inputs(512, 640, 2, 1, smooth=True, translation=(0.02, 0.22, 0.30))
    rep, _, _ = run_case(net, gta_state, inputs, 24, stages=False)
    _assert_report(rep)
    flags = net.get_stage("recurrence_flags", torch.int32).view(-1, 17).cpu()
    assert int(flags[:, 16].max()) == 1, flags
    assert int(flags[:, :11].min()) >= 22          # every CTA published up to the second-to-last hypothesis
    with torch.no_grad():
        net(*synthetic.to_device(synthetic.make_inputs(512, 640, 1, 1), "cuda"), 64, True, [True] * 5)
    flags = net.get_stage("recurrence_flags", torch.int32).view(-1, 17).cpu()
    assert int(flags[:, 16].max()) == 0 and int(flags[:, :16].max()) == 0, flags


def test_l4_chain_matches_per_layer_kernels(net):
    """The level-4 tail of the feature network as one cluster kernel (option l4_chain, default) against the seven
    per-layer launches: same arithmetic (split-fp16 MMAs, fp32 accumulation), different statistics summation order."""
    inputs = synthetic.to_device(synthetic.make_inputs(512, 640, 2, 2, smooth=True), "cuda")
    with torch.no_grad():
        a = net(*inputs, 64, True, [True] * 5)
        na = net.last_launch_count()
        net.set_option("l4_chain", 0)
        try:
            b = net(*inputs, 64, True, [True] * 5)
            nb = net.last_launch_count()
        finally:
            net.set_option("l4_chain", 1)
    assert nb - na == 12, (na, nb)                 # two feature networks x (7 launches -> 1)
    for lvl in range(5):
        # (float32 statistics summed in a fixed tree vs float64 atomics: ~1e-6 per layer, amplified by the 63-step sweep)
        assert rel_linf(a["left_idepthmap_pyr"][lvl].cpu(), b["left_idepthmap_pyr"][lvl].cpu()) <= REL_LINF_TOL / 2
        assert bool((a["left_idepthmap_mask_pyr"][lvl] == b["left_idepthmap_mask_pyr"][lvl]).all())


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
    (64x80 = 41 M-tiles) is beyond the persistent recurrence kernel's cluster, so this exercises the wide sweep
    (sweep_wide.cu; 4 chains of 21 two-tile CTAs), and the full-resolution warp has a knife-edge mask pixel on these inputs
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process(gta_state):
    """Two handles on different GPUs in one process: the > 48 KB shared-memory and non-portable-cluster
    opt-ins are per device, so the second GPU must get its own."""
    from tests._gpu_util import make_net
    inputs = synthetic.make_inputs(512, 640, 1, 1)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = make_net(gta_state, dev)
        with torch.no_grad():
            outs.append(net(*synthetic.to_device(inputs, dev), 64, True, [True] * 5)["left_idepthmap_pyr"][0].cpu())
    assert rel_linf(outs[1], outs[0]) <= 1e-5
