"""What the reference's eval driver does around the hot path (test.py), on the device, with the reference's names:

  load_params                    test.py:336-344     params.yaml (+ the keys the DeMoN file lacks)
  load_models                    test.py:307-316     weights out of the reference's TorchScript archive
  get_groundtruth_limits         test.py:167-186     depth limits per split
  idepthmap_to_depthmap          test.py:211-213     idepth / baseline -> depth
  get_depth_prediction_metrics   test.py:41-71       abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
  evaluate_batch                 test.py:211-236, 258-263   all of the above for a batch, one kernel
  write_metrics_header / write_metrics / write_runtime_metrics / compute_avg_metrics   test.py:124-165, 265-271
  test                           test.py:188-281     the loop over a loader (without the training losses)

The reference copies every estimate to the host and evaluates it with numpy; here `b200mvs_depth_metrics` reads the
estimate and the ground truth once on the device and returns eight numbers per image.  There is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

METRIC_KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def load_params(params_file):
    """params.yaml as test.py:336-340 reads it; `cost_volume_filter` / `refiners` default to the full network where
    the file predates them (pretrained/demon_45epochs/params.yaml has neither key)."""
    import yaml
    with open(params_file, "r") as stream:
        params = yaml.load(stream, Loader=yaml.FullLoader)
    params.setdefault("num_levels", 5)
    params.setdefault("cost_volume_filter", True)
    params.setdefault("refiners", [True] * params["num_levels"])
    return params


def load_models(device, weights_dir, params=None):
    """test.py:307-316.  `torch.jit.load` of the shipped archives fails under torch >= 2; the weights
    are read out of the archive instead and loaded into the B200 module."""
    from . import weights
    from .multi_view_stereonet import MultiViewStereoNet
    state = weights.load_torchscript_archive_weights(os.path.join(weights_dir, "stereo_network.pt"))
    stereo_network = MultiViewStereoNet()
    stereo_network.load_state_dict(state, strict=True)
    stereo_network = stereo_network.to(device)
    stereo_network.eval()
    return stereo_network


def get_groundtruth_limits(split):
    """(min_depth, max_depth) of get_groundtruth_depthmap (test.py:167-186)."""
    if "gta_sfm" in split:
        return 0.0, 1e3
    if "demon" in split:
        return 0.5, 10.0   # limits from DPSNet
    raise AssertionError("unknown split: %s" % (split,))


def _call(est, baseline, depth_true, est_is_depth, lo, hi, want_maps, want_metrics):
    if est.device.type != "cuda":
        raise RuntimeError("evaluation (B200) needs CUDA tensors; there is no CPU path")
    lib = _lib.load()
    est = est.detach().to(torch.float32).contiguous()
    batch = est.shape[0] if est.dim() > 1 else 1
    pixels = est.numel() // batch
    dev = est.device
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    bl = None if baseline is None else baseline.detach().to(dev, torch.float32).contiguous().view(-1)
    if bl is not None:
        assert bl.numel() == batch
    gt = None if depth_true is None else depth_true.detach().to(dev, torch.float32).contiguous()
    if gt is not None:
        assert gt.numel() == est.numel()
    idepth = torch.empty_like(est) if want_maps else None
    depth = torch.empty_like(est) if want_maps else None
    metrics = torch.empty((batch, 8), dtype=torch.float64, device=dev) if want_metrics else None
    ptr = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(index):
        stream = ctypes.c_void_p(torch.cuda.current_stream(index).cuda_stream)
        _lib.check(lib.b200mvs_depth_metrics(ptr(est), ptr(bl), ptr(gt), int(est_is_depth), float(lo), float(hi), batch,
                                             pixels, ptr(idepth), ptr(depth), ptr(metrics), stream),
                   "b200mvs_depth_metrics")
    return idepth, depth, metrics


def idepthmap_to_depthmap(left_idepthmap, baseline):
    """test.py:211-213: (idepth / baseline, depth) with depth = 1 / idepth where positive, 0 elsewhere."""
    idepth, depth, _ = _call(left_idepthmap, baseline, None, False, 0.0, 0.0, True, False)
    return idepth, depth


def get_depth_prediction_metrics(depthmap_true, depthmap_est):
    """test.py:41-71 on CUDA tensors holding the already-masked depths ("assumes no invalid inputs")."""
    inf = float("inf")
    _, _, m = _call(depthmap_est.reshape(1, -1), None, depthmap_true.reshape(1, -1), True, -inf, inf, False, True)
    row = m[0].tolist()
    return {k: row[i] for i, k in enumerate(METRIC_KEYS)}


def evaluate_batch(left_idepthmap, baseline, left_depthmap_true, split):
    """Depth conversion, validity mask and metrics for a batch in one pass (test.py:211-236, 258-263).

    left_idepthmap (B,1,H,W) = outputs["left_idepthmap_pyr"][0]; baseline (B,) = inputs["baseline"];
    left_depthmap_true (B,1,H,W) = inputs["left_depthmap_true"] (baseline-normalised, as unpacked).
    Returns (idepth_est, depth_est, metrics) where metrics[b] is the reference's dict for item b, or None when no
    pixel has valid ground truth and a valid estimate (the reference skips such images, test.py:224-226), plus
    "num_valid"."""
    lo, hi = get_groundtruth_limits(split)
    idepth, depth, m = _call(left_idepthmap, baseline, left_depthmap_true, False, lo, hi, True, True)
    rows = m.tolist()          # the only device -> host copy: 8 doubles per image
    metrics = []
    for row in rows:
        if row[7] <= 0:
            metrics.append(None)
            continue
        d = {k: row[i] for i, k in enumerate(METRIC_KEYS)}
        d["num_valid"] = int(row[7])
        metrics.append(d)
    return idepth, depth, metrics


# ---- files (test.py:124-165, 265-271) ---------------------------------------------------------------------------
def write_metrics_header(output_file, metrics_dict):
    with open(output_file, "w") as ff:
        ff.write("file ")
        for key in METRIC_KEYS:
            if key in metrics_dict:
                ff.write("{} ".format(key))
        ff.write("\n")


def write_metrics(output_file, input_file, metrics_dict):
    with open(output_file, "a") as ff:
        ff.write("{} ".format(input_file))
        for key in METRIC_KEYS:
            if key in metrics_dict:
                ff.write("{} ".format(metrics_dict[key]))
        ff.write("\n")


def write_runtime_metrics(output_file, input_file, runtime_ms):
    if not os.path.exists(output_file):
        with open(output_file, "w") as stream:
            stream.write("file runtime_ms\n")
    with open(output_file, "a") as stream:
        stream.write("{} {}\n".format(input_file, runtime_ms))


def compute_avg_metrics(metrics_file):
    with open(metrics_file, "r") as ff:
        keys = ff.readline().split()[1:]   # skip the file name
    metrics = np.atleast_2d(np.loadtxt(metrics_file, skiprows=1, usecols=range(1, len(keys) + 1)))
    avg = np.mean(metrics, axis=0)
    out = {k: avg[i] for i, k in enumerate(keys)}
    out["num_samples"] = metrics.shape[0]
    return out


def test(split, device, stereo_network, loader, save_images, output_dir, params):
    """The reference's evaluation loop (test.py:188-281) without the training losses (`compute_losses`, out of
    scope) and the debug images: unpack -> timed forward -> depth + metrics on the device -> the reference's
    depth_metrics.txt / runtime_metrics.txt.  Returns the number of batches."""
    from . import multi_view_stereonet_utils as snu
    assert not save_images, "debug image dumps are not part of this build (DESIGN.md: out of scope)"
    stereo_network.eval()
    os.makedirs(output_dir, exist_ok=True)
    depth_metrics_file = os.path.join(output_dir, "depth_metrics.txt")
    runtime_metrics_file = os.path.join(output_dir, "runtime_metrics.txt")
    num_batches = 0
    with torch.no_grad():
        for batch in loader:
            inputs = snu.multi_view_unpack_batch(batch, device, stereo_network.num_levels)
            outputs = snu.multi_view_forward(stereo_network, inputs, params)
            num_batches += 1
            _, _, metrics = evaluate_batch(outputs["left_idepthmap_pyr"][0], inputs["baseline"],
                                           inputs["left_depthmap_true"], split)
            names = inputs.get("left_filename") or ["item%d" % i for i in range(len(metrics))]
            for idx, m in enumerate(metrics):
                left_file = names[idx]
                if m is None:
                    print("WARNING: No truth for image: {}".format(left_file))
                    continue
                if not os.path.exists(depth_metrics_file):
                    write_metrics_header(depth_metrics_file, m)
                write_metrics(depth_metrics_file, left_file, m)
                write_runtime_metrics(runtime_metrics_file, left_file, outputs["stereo_time_ms"])
    return num_batches
