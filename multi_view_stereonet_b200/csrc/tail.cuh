// 32 -> 1 tail convolutions of IDepthmapRefiner and CostVolumeFilter (tail.cu).
#pragma once
#include "common.cuh"

namespace b200mvs {

// Reference weight layouts: conv_final.weight (1, 32, 3, 3) -> w[c * 9 + tap]; volume_filter4.conv4.weight
// (1, 32, 3, 3, 3) -> w[c * 27 + tap].  Passed to the kernels by value (constant bank).
struct RefineFinalW {
  float w[32 * 9];
  float bias;
};
// refiner0.conv0.weight (32, 4, 3, 3) in the reference's order w[(o * 4 + c) * 9 + tap], and its bias.
struct RefineHeadW {
  float w[32 * 4 * 9];
  float bias[32];
};
// The idepth channel of an IDepthmapRefiner.conv0 (the last input channel): w[tap * 32 + o].
struct RefineHeadIdW {
  float w[9 * 32];
};
struct CvfFinalW {
  float w[32 * 27];
  float bias;
};

// out = relu(prior * fx + conv3x3(lrelu(GN(y)) + resid) + b) / fx     (multi_view_stereonet.py:479-483, 607-611)
int launch_refine_final(const void* y, const void* resid, bool half_io, const double* stats, const float* gamma,
                        const float* beta, double inv_count, const RefineFinalW& w, const float* prior,
                        const float* fx, int fx_div, int fx_stride, int n, int H, int W, float* out,
                        cudaStream_t stream);

// IDepthmapRefiner.conv0 at level 0 (4 planar input channels: image (3) and idepth * fx; :469-470, 679-681):
// y = conv3x3(cat[image, prior * fx]) + b as fp16 or fp32 channels-last, plus the GroupNorm statistics of y.
int launch_refine_head_l0(const float* image, const float* prior, const float* fx, int fx_div, int fx_stride,
                          const RefineHeadW& w, int n, int H, int W, void* out, bool out_half, double* out_stats,
                          cudaStream_t stream);

// conv0 split into its guide part and its idepth part (tail.cu: refine_head_pre_kernel).
//   launch_refine_head_image_l0: pre = conv3x3(image) + b at level 0, fp32 channels-last (no statistics)
//   launch_refine_head_pre:      y = pre[img / pre_div] + conv3x3(prior * fx; idepth-channel weights), statistics of y
int launch_refine_head_image_l0(const float* image, const RefineHeadW& w, int n, int H, int W, float* pre,
                                cudaStream_t stream);
//                                coarse != nullptr: the prior is the bilinear upsampling of coarse (n, ch, cw),
//                                computed on the fly and also written to prior_out (n, H, W); else `prior` is read
int launch_refine_head_pre(const float* pre, int pre_div, const float* prior, const float* coarse, int ch, int cw,
                           float* prior_out, const float* fx, int fx_div, int fx_stride, const RefineHeadIdW& w, int n,
                           int H, int W, void* out, bool out_half, double* out_stats, cudaStream_t stream);

// cost1 = conv3d(lrelu(GN(y)), 32 -> 1) + b ; raw = soft-argmin over D (:350-352, 486-492).
// `part` is scratch of n * D * 27 * h * w floats.
bool cvf_final_supported(int D);
int launch_cvf_final(const float* y, const double* stats, const float* gamma, const float* beta, double inv_count,
                     const CvfFinalW& w, const float* samples, int n, int D, int h, int wd, float* part, float* cost1,
                     float* raw, cudaStream_t stream);

}  // namespace b200mvs
