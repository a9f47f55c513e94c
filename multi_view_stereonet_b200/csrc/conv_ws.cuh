// Persistent warp-specialised tcgen05 convolution for the refiner residual blocks (conv_ws.cu).
#pragma once
#include "conv.cuh"

namespace b200mvs {

// 3x3 (dilated) 32->32, fp16 activations in and out, input = lrelu(GN(y)) (+ resid); weights as packed by
// pack_conv3x3_tc_weights(..., split = false).  Supported only for layers with at least one tile per SM.
bool conv3x3_ws_supported(const ConvParams& p);
int launch_conv3x3_ws(const ConvParams& p, const uint8_t* w16, cudaStream_t stream);

}  // namespace b200mvs
