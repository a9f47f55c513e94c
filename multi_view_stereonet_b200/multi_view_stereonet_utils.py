"""Host glue around the hot path, mirroring the reference's `multi_view_stereonet/multi_view_stereonet_utils.py`
(same function names, arguments and dictionary keys):

  build_image_pyramid       utils/image_utils.py:111-128      area pyramid, on the device
  multi_view_unpack_batch   :541-641                           pyramids, per-level K, inverse poses, baseline
  multi_view_forward        :643-662                           the timed call of the network
  unpack_batch, forward     :406-539                           the two-view variants of the same pair (one comparison
                                                               image, optional right-view estimate)

The pyramids, intrinsics and pose normalisation run in CUDA kernels (b200mvs_area_downsample,
b200mvs_prepare_cameras); ground-truth depth maps, which only the losses read, are rescaled with plain tensor ops.
"""
import ctypes

import torch

from . import _lib


def _stream(device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    return index, ctypes.c_void_p(torch.cuda.current_stream(index).cuda_stream)


def build_image_pyramid(image, num_levels):
    """Image pyramid by repeated area half-downsampling to ((h+1)//2, (w+1)//2) (utils/image_utils.py:111-128)."""
    assert len(image.shape) == 4
    if image.device.type != "cuda":
        raise RuntimeError("build_image_pyramid (B200) needs a CUDA tensor; there is no CPU path")
    lib = _lib.load()
    pyramid = [image.detach().to(torch.float32).contiguous()]
    index, stream = _stream(image.device)
    with torch.cuda.device(index):
        for _ in range(1, num_levels):
            src = pyramid[-1]
            n, c, h, w = src.shape
            dst = torch.empty((n, c, (h + 1) // 2, (w + 1) // 2), dtype=torch.float32, device=src.device)
            _lib.check(lib.b200mvs_area_downsample(src.data_ptr(), n * c, h, w, dst.data_ptr(), stream),
                       "b200mvs_area_downsample")
            pyramid.append(dst)
    return pyramid


def multi_view_unpack_batch(batch, device, num_levels):
    """Unpack a batch of data (multi_view_stereonet_utils.py:541-641)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("multi_view_unpack_batch (B200) needs a CUDA device; there is no CPU path")
    lib = _lib.load()
    left_image = batch["left_image"].to(device, non_blocking=True)
    right_image = [r.to(device, non_blocking=True) for r in batch["right_image"]]
    left_image_pyr = build_image_pyramid(left_image, num_levels)
    right_image_pyr = [build_image_pyramid(r, num_levels) for r in right_image]

    K = torch.squeeze(batch["K"].to(device, non_blocking=True), dim=1).to(torch.float32).contiguous()
    Ts = [torch.squeeze(t.to(device, non_blocking=True), dim=1).to(torch.float32).contiguous()
          for t in batch["T_right_in_left"]]
    B, V = K.shape[0], len(Ts)
    f32 = dict(dtype=torch.float32, device=device)
    K_all = torch.empty((num_levels, B, 4, 4), **f32)
    T_all = torch.empty((V, B, 4, 4), **f32)
    Tinv_all = torch.empty((V, B, 4, 4), **f32)
    baseline = torch.empty((B,), **f32)
    sizes = (ctypes.c_int32 * (2 * num_levels))()
    for lvl, img in enumerate(left_image_pyr):
        sizes[2 * lvl], sizes[2 * lvl + 1] = img.shape[-2], img.shape[-1]
    index, stream = _stream(device)
    with torch.cuda.device(index):
        _lib.check(lib.b200mvs_prepare_cameras(K.data_ptr(), _lib.ptr_array([t.data_ptr() for t in Ts]), B, V,
                                               num_levels, sizes, K_all.data_ptr(), T_all.data_ptr(),
                                               Tinv_all.data_ptr(), baseline.data_ptr(), stream),
                   "b200mvs_prepare_cameras")
    # the reference asserts every baseline is positive (:598-599)
    assert int(torch.sum(baseline > 0)) == baseline.shape[0]

    inputs = {"left_filename": batch.get("left_filename"),
              "right_filename": batch.get("right_filename"),
              "T_right_in_left": [T_all[v] for v in range(V)],
              "T_left_in_right": [Tinv_all[v] for v in range(V)],
              "K_pyr": [K_all[lvl] for lvl in range(num_levels)],
              "left_image_pyr": left_image_pyr,
              "right_image_pyr": right_image_pyr,
              "baseline": baseline}

    if "left_depthmap_true" in batch:      # ground truth for the losses only (:613-637)
        scale = baseline.view(-1, 1, 1, 1)
        inputs["left_depthmap_true"] = batch["left_depthmap_true"].to(device) / scale
        inv = inputs["left_depthmap_true"].clone()
        inv[inv > 0] = 1.0 / inv[inv > 0]
        inputs["left_idepthmap_true"] = inv
        inputs["right_depthmap_true"] = [r.to(device) / scale for r in batch["right_depthmap_true"]]
        inputs["right_idepthmap_true"] = []
        for r in inputs["right_depthmap_true"]:
            inv = r.clone()
            inv[inv > 0] = 1.0 / inv[inv > 0]
            inputs["right_idepthmap_true"].append(inv)

    assert inputs["left_image_pyr"][0].dtype == torch.float32
    return inputs


def multi_view_forward(stereo_network, inputs, params):
    """Forward pass with the reference's CUDA-event timer around it (multi_view_stereonet_utils.py:643-662,
    utils/pytorch_utils.py:31-48)."""
    torch.cuda.synchronize()
    tick, tock = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tick.record()
    left_outputs = stereo_network(inputs["left_image_pyr"], inputs["K_pyr"], inputs["T_right_in_left"],
                                  inputs["right_image_pyr"], params["num_idepth_samples"],
                                  params["cost_volume_filter"], params["refiners"])
    tock.record()
    torch.cuda.synchronize()
    return {"left_idepthmap_pyr": left_outputs["left_idepthmap_pyr"],
            "left_idepthmap_raw_pyr": left_outputs["left_idepthmap_raw_pyr"],
            "left_idepthmap_mask_pyr": left_outputs["left_idepthmap_mask_pyr"],
            "stereo_time_ms": tick.elapsed_time(tock)}


def unpack_batch(batch, device, num_levels):
    """Two-view unpack (multi_view_stereonet_utils.py:406-501): `right_image`, `T_right_in_left` (and the right
    ground truth) are single tensors instead of per-view lists; same kernels as `multi_view_unpack_batch`."""
    multi = dict(batch)
    multi["right_image"] = [batch["right_image"]]
    multi["T_right_in_left"] = [batch["T_right_in_left"]]
    if "left_depthmap_true" in batch:
        multi["right_depthmap_true"] = [batch["right_depthmap_true"]]
    inputs = multi_view_unpack_batch(multi, device, num_levels)
    for key in ("T_right_in_left", "T_left_in_right", "right_image_pyr", "right_depthmap_true", "right_idepthmap_true"):
        if key in inputs:
            inputs[key] = inputs[key][0]
    if "left_disparity_true" in batch:     # passed through to the device unchanged (:469-474)
        inputs["left_disparity_true"] = batch["left_disparity_true"].to(device)
        inputs["right_disparity_true"] = batch["right_disparity_true"].to(device)
    return inputs


def _timed_call(stereo_network, image_pyr, K_pyr, T, other_pyr, params):
    torch.cuda.synchronize()
    tick, tock = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tick.record()
    out = stereo_network(image_pyr, K_pyr, [T], [other_pyr], params["num_idepth_samples"],
                         params["cost_volume_filter"], params["refiners"])
    tock.record()
    torch.cuda.synchronize()
    return out, tick.elapsed_time(tock)


def forward(stereo_network, inputs, params):
    """Two-view forward (multi_view_stereonet_utils.py:503-539): the left estimate and, if
    params["estimate_right_idepthmap"], the right one with the roles of the two images swapped
    (`T_left_in_right`); `stereo_time_ms` is then the mean of the two calls."""
    left, ms = _timed_call(stereo_network, inputs["left_image_pyr"], inputs["K_pyr"], inputs["T_right_in_left"],
                           inputs["right_image_pyr"], params)
    outputs = {"left_idepthmap_pyr": left["left_idepthmap_pyr"],
               "left_idepthmap_raw_pyr": left["left_idepthmap_raw_pyr"],
               "left_idepthmap_mask_pyr": left["left_idepthmap_mask_pyr"],
               "stereo_time_ms": ms}
    if params["estimate_right_idepthmap"]:
        right, right_ms = _timed_call(stereo_network, inputs["right_image_pyr"], inputs["K_pyr"],
                                      inputs["T_left_in_right"], inputs["left_image_pyr"], params)
        outputs["right_idepthmap_pyr"] = right["left_idepthmap_pyr"]
        outputs["right_idepthmap_raw_pyr"] = right["left_idepthmap_raw_pyr"]
        outputs["right_idepthmap_mask_pyr"] = right["left_idepthmap_mask_pyr"]
        outputs["stereo_time_ms"] = 0.5 * (outputs["stereo_time_ms"] + right_ms)
    return outputs
