// tcgen05 tensor-core convolution entry points (conv_tc.cu).
#pragma once
#include <vector>

#include "conv.cuh"

namespace b200mvs {

// Packs a reference (32, cin, 3, 3) weight into the fp16 UMMA-canonical blocks the kernel stages verbatim:
// [tap][k-step] blocks, k-steps = two over the 32 feature channels (reference input index feat_off + c) if
// has_feat, then one over the planar extra channels (reference indices extra_idx, at most 4).  split = blocks
// hold [W_hi | W_lo] (N = 64) for the split-precision kernel.
void pack_conv3x3_tc_weights(const float* w_oihw, int cin, bool has_feat, int feat_off,
                             const std::vector<int>& extra_idx, bool split, std::vector<uint8_t>* out);
bool conv3x3_tc_supported(const ConvParams& p);
// Same contract as launch_conv(CONV_3x3, 32, ...).  split selects hi/lo fp16 operands (fp32-class accuracy).
int launch_conv3x3_tc(const ConvParams& p, const uint8_t* w16, bool split, cudaStream_t stream);

}  // namespace b200mvs
