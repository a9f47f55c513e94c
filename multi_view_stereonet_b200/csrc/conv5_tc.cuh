// 5x5 stride-2 convolutions of FeatureNetwork on tcgen05 tensor cores (conv5_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace b200mvs {

// Reference weights (32, 32, 5, 5) / (32, 3, 5, 5) -> split hi/lo fp16 UMMA blocks in kernel issue order.
void pack_conv5_c32_weights(const float* w_oihw, std::vector<uint8_t>* out);
void pack_conv5_c3_weights(const float* w_oihw, std::vector<uint8_t>* out);

// in: (n, Hi, Wi, 32) channels-last fp32 -> out: (n, ceil(Hi/2), ceil(Wi/2), 32); padding 2, no bias.
int launch_conv5x5s2_c32_tc(const float* in, const uint8_t* w16, int n, int Hi, int Wi, float* out, cudaStream_t stream);
// in: (n, 3, Hi, Wi) planar fp32 -> out channels-last as above.
int launch_conv5x5s2_c3_tc(const float* in, const uint8_t* w16, int n, int Hi, int Wi, float* out, cudaStream_t stream);

}  // namespace b200mvs
