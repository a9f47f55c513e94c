"""Drop-ins for the layers of the reference's `stereo/image_predictor.py`: the homography warp on the hot path and
the reprojection layers next to it (`ImagePredictor` and its parts), same class names and `forward` signatures."""
import ctypes

import torch
import torch.nn as tnn

from . import _lib


class HomographyImagePredictor(tnn.Module):
    """Predicts an image from a source image and a homography
    (reference stereo/image_predictor.py:463-523): bilinear resampling with
    border clamping plus the out-of-image mask (True = invalid)."""

    def forward(self, H_left_in_right, right_image):
        assert len(H_left_in_right.shape) == 3
        assert H_left_in_right.shape[1] == 3
        assert H_left_in_right.shape[2] == 3
        if right_image.device.type != "cuda":
            raise RuntimeError("HomographyImagePredictor (B200) needs CUDA tensors; there is no CPU path")
        lib = _lib.load()
        n, c, rows, cols = right_image.shape
        assert H_left_in_right.shape[0] == n
        H = H_left_in_right.detach().to(torch.float32).contiguous()
        img = right_image.detach().to(torch.float32).contiguous()
        pred = torch.empty_like(img)
        mask = torch.empty((n, 1, rows, cols), dtype=torch.uint8, device=img.device)
        dev = img.device.index if img.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for i0 in range(0, n, 65535):
                i1 = min(n, i0 + 65535)
                _lib.check(lib.b200mvs_homography_warp(H[i0:i1].data_ptr(), img[i0:i1].data_ptr(), i1 - i0, c, rows,
                                                       cols, 0, pred[i0:i1].data_ptr(), mask[i0:i1].data_ptr(),
                                                       ctypes.c_void_p(stream)), "b200mvs_homography_warp")
        return pred, mask.view(torch.bool)


# map kinds of b200mvs_reproject (include/b200mvs.h)
_IDEPTH, _DISPARITY, _RECTIFIED = 0, 1, 2


def _check_KT(K, T_right_in_left):
    assert len(T_right_in_left.shape) == 3
    assert T_right_in_left.shape[1] == 4
    assert T_right_in_left.shape[2] == 4
    assert len(K.shape) == 3
    assert K.shape[1] == 4
    assert K.shape[2] == 4


def _reproject(K, T_right_in_left, left_map, kind, right_image=None, want=()):
    """One launch of the fused reprojection kernel; `want` names the optional outputs to produce."""
    _check_KT(K, T_right_in_left)
    if left_map.device.type != "cuda":
        raise RuntimeError("image_predictor (B200) needs CUDA tensors; there is no CPU path")
    lib = _lib.load()
    dev = left_map.device
    n = K.shape[0]
    rows, cols = left_map.shape[-2], left_map.shape[-1]
    f32 = dict(dtype=torch.float32, device=dev)
    Kc = K.detach().to(**f32).contiguous()
    Tc = T_right_in_left.detach().to(**f32).contiguous()
    mc = left_map.detach().to(**f32).contiguous()
    assert mc.numel() == n * rows * cols
    out = {}
    img = None
    channels = 0
    if right_image is not None:
        img = right_image.detach().to(**f32).contiguous()
        channels = img.shape[1]
        assert img.shape[0] == n and img.shape[-2] == rows and img.shape[-1] == cols
        out["pred"] = torch.empty_like(img)
    if "mask" in want:
        out["mask"] = torch.empty((n, 1, rows, cols), dtype=torch.uint8, device=dev)
    if "right_pixels" in want:
        out["right_pixels"] = torch.empty((n, rows, cols, 2), **f32)
    for name in ("right_idepths", "idepth", "disparity"):
        if name in want:
            out[name] = torch.empty(left_map.shape, **f32)

    def ptr(name):
        return out[name].data_ptr() if name in out else None

    index = dev.index if dev.index is not None else torch.cuda.current_device()
    with torch.cuda.device(index):
        stream = torch.cuda.current_stream(index).cuda_stream
        _lib.check(lib.b200mvs_reproject(Kc.data_ptr(), Tc.data_ptr(), mc.data_ptr(), kind,
                                         img.data_ptr() if img is not None else None, n, channels, rows, cols,
                                         ptr("pred"), ptr("mask"), ptr("right_pixels"), ptr("right_idepths"),
                                         ptr("idepth"), ptr("disparity"), ctypes.c_void_p(stream)),
                   "b200mvs_reproject")
    if "mask" in out:
        out["mask"] = out["mask"].view(torch.bool)
    return out


def disparity_to_idepth(K, T_right_in_left, left_disparity):
    """General (non-rectified) disparities -> inverse depths (reference stereo/image_predictor.py:120-209)."""
    return _reproject(K, T_right_in_left, left_disparity, _DISPARITY, want=("idepth",))["idepth"]


class DisparityToIDepth(tnn.Module):
    """reference stereo/image_predictor.py:211-218"""

    def forward(self, K, T_right_in_left, left_disparity):
        return disparity_to_idepth(K, T_right_in_left, left_disparity)


class IDepthToDisparity(tnn.Module):
    """Inverse depthmap -> (non-rectified) disparities (reference stereo/image_predictor.py:220-273)."""

    def forward(self, K, T_right_in_left, left_idepthmap):
        return _reproject(K, T_right_in_left, left_idepthmap, _IDEPTH, want=("disparity",))["disparity"]


class IDepthmapProjector(tnn.Module):
    """Projects a left inverse depthmap to the right frame: right pixels (normalised grid coordinates), right
    inverse depths, out-of-image mask (reference stereo/image_predictor.py:525-576)."""

    def forward(self, K, T_right_in_left, left_idepthmap):
        o = _reproject(K, T_right_in_left, left_idepthmap, _IDEPTH, want=("right_pixels", "right_idepths", "mask"))
        return o["right_pixels"], o["right_idepths"], o["mask"]


class IDepthImagePredictor(tnn.Module):
    """Predicts the left image from the right image and a left inverse depthmap
    (reference stereo/image_predictor.py:347-398)."""

    def forward(self, K, T_right_in_left, left_idepthmap, right_image):
        o = _reproject(K, T_right_in_left, left_idepthmap, _IDEPTH, right_image, want=("mask",))
        return o["pred"], o["mask"]


class ImagePredictor(tnn.Module):
    """Predicts the left image from the right image and a left disparity map
    (reference stereo/image_predictor.py:578-601)."""

    def forward(self, K, T_right_in_left, left_disparity, right_image):
        o = _reproject(K, T_right_in_left, left_disparity, _DISPARITY, right_image, want=("mask",))
        return o["pred"], o["mask"]


class RectifiedImagePredictor(tnn.Module):
    """Predicts the left image from the right image and a rectified disparity map in pixels
    (reference stereo/image_predictor.py:275-345)."""

    def forward(self, K, T_right_in_left, left_disparity, right_image):
        o = _reproject(K, T_right_in_left, left_disparity, _RECTIFIED, right_image, want=("mask",))
        return o["pred"], o["mask"]
