"""Post-processing row: depth conversion, validity mask and depth metrics after the forward.
Fixture tests/golden/evaluation_small.npz comes from the reference's own functions (make_golden_evaluation.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import evaluation_oracle as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SPLITS = {"gta": "gta_sfm_overlap0.5_test.txt", "demon": "demon_test.txt"}
# float32 per-pixel terms are identical; the reference averages them in float32 (numpy pairwise), we in float64
MEAN_TOL = 2e-6


def fixture():
    z = np.load(os.path.join(HERE, "golden", "evaluation_small.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", ["gta", "demon"])
def test_oracle_matches_reference(name):
    f = fixture()
    idepth, depth = oracle.idepthmap_to_depthmap(f[f"{name}_idepth"], f[f"{name}_baseline"])
    assert np.array_equal(idepth, f[f"{name}_idepth_est"]) and np.array_equal(depth, f[f"{name}_depth_est"])
    res = oracle.evaluate(f[f"{name}_idepth"], f[f"{name}_baseline"], f[f"{name}_truth_norm"], SPLITS[name])
    for b, (m, count) in enumerate(res):
        assert count == int(f[f"{name}_counts"][b])
        if count == 0:
            assert m is None
            continue
        ref = f[f"{name}_metrics"][b]
        for i, k in enumerate(oracle.METRIC_KEYS):
            assert abs(m[k] - ref[i]) <= MEAN_TOL * max(1.0, abs(ref[i])), (k, m[k], ref[i])


def test_metric_files_roundtrip(tmp_path):
    from multi_view_stereonet_b200 import evaluation as ev
    m = {k: 0.1 * (i + 1) for i, k in enumerate(ev.METRIC_KEYS)}
    path = str(tmp_path / "depth_metrics.txt")
    ev.write_metrics_header(path, m)
    ev.write_metrics(path, "a.png", m)
    ev.write_metrics(path, "b.png", {k: 3 * v for k, v in m.items()})
    avg = ev.compute_avg_metrics(path)
    assert avg["num_samples"] == 2 and abs(avg["abs_rel"] - 0.2) < 1e-12 and abs(avg["a3"] - 1.4) < 1e-12
    assert open(path).readline().split() == ["file"] + list(ev.METRIC_KEYS)      # header as test.py:124-132
    rt = str(tmp_path / "runtime_metrics.txt")
    ev.write_runtime_metrics(rt, "a.png", 1.5)
    ev.write_runtime_metrics(rt, "b.png", 2.5)
    assert np.loadtxt(rt, skiprows=1, usecols=1).mean() == 2.0                  # test.py:381-386


def test_load_params_defaults(tmp_path):
    from multi_view_stereonet_b200 import evaluation as ev
    p = tmp_path / "params.yaml"
    p.write_text("num_levels: 5\nnum_idepth_samples: 12\nsize: [480, 640]\n")    # the DeMoN file has no filter keys
    params = ev.load_params(str(p))
    assert params["cost_volume_filter"] is True and params["refiners"] == [True] * 5 and params["num_idepth_samples"] == 12
    assert ev.get_groundtruth_limits("demon_test.txt") == (0.5, 10.0) and ev.get_groundtruth_limits("x/gta_sfm_y") == (0.0, 1e3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gta", "demon"])
def test_cuda_evaluate_batch_matches_reference(name):
    from multi_view_stereonet_b200 import evaluation as ev
    f = fixture()
    dev = torch.device("cuda:0")
    t = lambda k: torch.from_numpy(f[f"{name}_{k}"]).to(dev)
    idepth, depth, metrics = ev.evaluate_batch(t("idepth"), t("baseline"), t("truth_norm"), SPLITS[name])
    assert np.array_equal(idepth.cpu().numpy(), f[f"{name}_idepth_est"])         # IEEE divisions: bit exact
    assert np.array_equal(depth.cpu().numpy(), f[f"{name}_depth_est"])
    i2, d2 = ev.idepthmap_to_depthmap(t("idepth"), t("baseline"))
    assert torch.equal(i2, idepth) and torch.equal(d2, depth)
    for b, m in enumerate(metrics):
        count = int(f[f"{name}_counts"][b])
        if count == 0:
            assert m is None
            continue
        assert m["num_valid"] == count                                             # the mask is bit exact
        ref = f[f"{name}_metrics"][b]
        for i, k in enumerate(ev.METRIC_KEYS):
            assert abs(m[k] - ref[i]) <= MEAN_TOL * max(1.0, abs(ref[i])), (k, m[k], ref[i])


@pytest.mark.gpu
def test_cuda_metrics_full_size_vs_oracle_and_masked_entry():
    """512x640 batch of 4 (odd pixel count variant too) against the oracle; get_depth_prediction_metrics on masked
    vectors equals the fused evaluation; repeated calls are bit-identical (fixed summation order)."""
    from multi_view_stereonet_b200 import evaluation as ev
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    for rows, cols in ((512, 640), (33, 41)):
        B = 4
        baseline = torch.rand(B, generator=g) * 0.4 + 0.2
        truth = torch.rand(B, 1, rows, cols, generator=g) * 30.0 + 0.01
        truth[torch.rand(B, 1, rows, cols, generator=g) < 0.2] = 0.0
        truth_norm = truth / baseline.view(-1, 1, 1, 1)
        est = torch.where(truth_norm > 0, 1.0 / (truth_norm * (0.8 + 0.4 * torch.rand(B, 1, rows, cols, generator=g))),
                          torch.zeros(()))
        ref = oracle.evaluate(est.numpy(), baseline.numpy(), truth_norm.numpy(), "gta_sfm")
        idepth, depth, metrics = ev.evaluate_batch(est.to(dev), baseline.to(dev), truth_norm.to(dev), "gta_sfm")
        _, _, again = ev.evaluate_batch(est.to(dev), baseline.to(dev), truth_norm.to(dev), "gta_sfm")
        assert metrics == again
        for b in range(B):
            assert metrics[b]["num_valid"] == ref[b][1]
            for k in ev.METRIC_KEYS:
                assert abs(metrics[b][k] - ref[b][0][k]) <= 1e-6 * max(1.0, abs(ref[b][0][k])), (k, b)
        # the reference's own calling convention: already-masked vectors
        t0 = (truth_norm[0] * baseline[0]).squeeze()
        d0 = depth[0].squeeze().cpu()
        mask = (t0 > 0.0) & (t0 < 1e3) & (d0 > 0.0) & (d0 < 1e3)
        m0 = ev.get_depth_prediction_metrics(t0[mask].to(dev), d0[mask].to(dev))
        for k in ev.METRIC_KEYS:
            assert abs(m0[k] - metrics[0][k]) <= 1e-9 * max(1.0, abs(m0[k])), k


@pytest.mark.gpu
def test_eval_loop_writes_reference_files(tmp_path):
    """test.py:188-281 with a synthetic loader: unpack -> forward -> metrics files."""
    from multi_view_stereonet_b200 import MultiViewStereoNet, evaluation as ev, synthetic
    from tests._util import load_gta_state
    net = MultiViewStereoNet()
    net.load_state_dict(load_gta_state(), strict=True)
    net = net.to("cuda:0").eval()
    loader = [synthetic.make_raw_batch(B=1, V=1, rows=64, cols=80, seed=s) for s in (1, 2)]
    for i, b in enumerate(loader):
        b["left_filename"] = ["scene/%04d.png" % i]
    params = {"num_idepth_samples": 8, "cost_volume_filter": True, "refiners": [True] * 5}
    out_dir = str(tmp_path / "output")
    n = ev.test("gta_sfm_test", torch.device("cuda:0"), net, loader, False, out_dir, params)
    assert n == 2
    avg = ev.compute_avg_metrics(os.path.join(out_dir, "depth_metrics.txt"))
    assert avg["num_samples"] == 2 and np.isfinite(avg["abs_rel"]) and 0.0 <= avg["a1"] <= 1.0
    assert len(open(os.path.join(out_dir, "runtime_metrics.txt")).readlines()) == 3
