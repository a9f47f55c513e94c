"""Drop-ins for the on-path layers of the reference's `stereo/image_predictor.py`."""
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
