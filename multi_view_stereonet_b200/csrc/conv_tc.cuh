// tcgen05 tensor-core convolution entry points (conv_tc.cu).
#pragma once
#include <vector>

#include "conv.cuh"

namespace b200mvs {

// Packs a reference (32, 32, 3, 3) weight into the fp16 UMMA-canonical layout the kernel stages verbatim.
void pack_conv3x3_tc_weights(const float* w_oihw, std::vector<uint8_t>* out);
bool conv3x3_tc_supported(const ConvParams& p);
// Same contract as launch_conv(CONV_3x3, 32, ...) for a 32-channel source without extra planes.
int launch_conv3x3_tc(const ConvParams& p, const uint8_t* w16, cudaStream_t stream);

}  // namespace b200mvs
