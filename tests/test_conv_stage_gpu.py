"""Kernel-level parity of the two 3x3 32->32 convolution kernels (fp32 FFMA and tcgen05 fp16-operand)
against torch.nn.functional.conv2d in fp32, through the C ABI stage entry b200mvs_conv3x3_c32."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# fp16 operands (11 significant bits) with fp32 accumulation over K = 288 products of O(1) values:
# error ~ 2^-11 * sqrt(288) * |x||w| relative to the output scale; 2e-3 of the max is a safe bound.
TC_TOL = 2e-3
FP32_TOL = 2e-5


def conv_stage(x_nhwc, w, bias, dilation, use_tc):
    from multi_view_stereonet_b200 import _lib
    lib = _lib.load()
    n, h, wd, c = x_nhwc.shape
    y = torch.empty_like(x_nhwc)
    wh = w.contiguous().cpu()
    bh = bias.contiguous().cpu() if bias is not None else None
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.b200mvs_conv3x3_c32(x_nhwc.data_ptr(), wh.data_ptr(), bh.data_ptr() if bh is not None else None,
                                 n, h, wd, dilation, int(use_tc), y.data_ptr(), ctypes.c_void_p(stream))
    _lib.check(rc, "b200mvs_conv3x3_c32")
    return y


@pytest.mark.parametrize("dilation", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(1, 128, 160), (2, 70, 131), (1, 16, 62)])
@pytest.mark.parametrize("use_tc", [0, 1, 2])
def test_conv3x3_c32(dilation, shape, use_tc):
    n, h, w = shape
    if use_tc == 2 and dilation == 8 and h > 100:
        pass
    g = torch.Generator().manual_seed(dilation * 100 + h)
    x = torch.randn(n, 32, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.1
    bias = torch.randn(32, generator=g)
    ref = F.conv2d(x, wt, bias, padding=dilation, dilation=dilation)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = conv_stage(x_nhwc, wt, bias, dilation, use_tc).permute(0, 3, 1, 2).cpu()
    err = float((y - ref).abs().max() / ref.abs().max())
    assert err <= (TC_TOL if use_tc == 1 else FP32_TOL), err
