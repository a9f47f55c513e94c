// Non-convolution kernels of the hot path: geometry, warps, cost build, soft-argmin, view
// reduction, upsampling.  Implementations in geom.cu / misc.cu.
#pragma once
#include "common.cuh"

namespace b200mvs {

constexpr int kMaxViews = 16;

// V per-view base pointers to (B, ...) tensors; image n = b * V + v reads p[v] + b * stride.
struct ViewPtrs {
  const float* p[kMaxViews];
  int views;
};

struct GeomOut {
  float* baseline;  // [n]
  float* samples;   // [n][D]
  float* H0;        // [n][9]     level-0 homography of hypothesis 0
  float* H;         // [n][D][9]  level-4 plane-sweep homographies
  float* Hinc;      // [n][D][9]  H_{d-1}^-1 H_d   (entry 0 unused = identity)
};

// Per (b, v): baseline normalisation, idepth samples, all homographies.
// multi_view_stereonet.py:566-576, 131-194, 280-282; stereo/image_predictor.py:120-209, 446-459.
int launch_geometry(const ViewPtrs& T, const float* K0, const float* K4, int batch, int D, int rows4, int cols4,
                    const GeomOut& out, cudaStream_t stream);

// HomographyImagePredictor.forward on planar images.  Image n reads src.p[n % views] + (n / views) * C*rows*cols.
int launch_warp_planar(const float* H, int h_stride, const ViewPtrs& src, int n, int channels, int rows, int cols,
                       bool zero_invalid, float* pred, uint8_t* mask, cudaStream_t stream);

// One step of the depth-sweep recurrence: warps the previous hypothesis' features by H_inc and the
// 1/16-scale right image by H_d (multi_view_stereonet.py:275, 285).
int launch_step_warp(const float* vol, const GeomOut& geo, const ViewPtrs& right_l4, int n, int D, int step,
                     int rows, int cols, float* wf, float* wimg, cudaStream_t stream);

// Image half of FeatureRefiner.conv0 (+ bias) for hypotheses 1..D-1 (see misc.cu): out [n][D][rows*cols][32], or
// [n][D][4 octets][rows*cols][8] with oct_major (what the wide sweep reads).
int launch_image_conv(const float* H, const ViewPtrs& right_l4, const float* w_tap8x32, const float* bias, int n,
                      int D, int rows, int cols, float* out, cudaStream_t stream, bool oct_major = false);

// cost = |L - R| where valid, 0 elsewhere; also emits the validity mask volume
// (multi_view_stereonet.py:293-298, 586-592).  `cost` may alias `vol`.
int launch_cost(const float* left_feat4, const float* vol, const float* H, int n, int views, int D, int rows,
                int cols, float* cost, uint8_t* mask, cudaStream_t stream);

// torch.norm(cost, dim=channels)  (multi_view_stereonet.py:598)
int launch_cost_norm(const float* cost, long long voxels, float* out, cudaStream_t stream);

// extract_idepthmap (multi_view_stereonet.py:486-492)
int launch_softargmin(const float* cost, const float* samples, int n, int D, int pixels, float* raw,
                      cudaStream_t stream);

// Per-view baseline un-normalisation and the mean over views (multi_view_stereonet.py:616-627).
int launch_mask_vote(const float* H, int batch, int views, int D, int rows, int cols, uint8_t* mask4,
                     cudaStream_t stream);
int launch_view_reduce(const float* raw_views, const float* refined_views, const uint8_t* mask_views,
                       const float* baseline, int batch, int views, int D, int pixels, bool refined_is_alias,
                       float* raw4, float* idepth4, uint8_t* mask4, cudaStream_t stream);

// Input preparation (multi_view_unpack_batch, multi_view_stereonet_utils.py:541-604): one pyramid level by area
// averaging, and the per-level intrinsics / inverse poses / baseline normalisation of a batch of image groups.
// level_sizes_dev: DEVICE int[2 * levels] = (rows_l, cols_l).
int launch_area_downsample(const float* in, int planes, int rows, int cols, float* out, cudaStream_t stream);
int launch_prepare_cameras(const float* K, const ViewPtrs& T, int batch, int levels, const int* level_sizes_dev,
                           float* K_pyr, float* T_norm, float* Tinv_norm, float* baseline, cudaStream_t stream);

// The reprojection layers of stereo/image_predictor.py (reproject.cu).  kind: 0 = `map` is an inverse depthmap,
// 1 = a general (non-rectified) disparity map, 2 = a rectified disparity map.  Every output pointer is optional.
int launch_reproject(const float* K, const float* T, const float* map, int kind, const float* right, int n,
                     int channels, int rows, int cols, float* pred, uint8_t* mask, float* right_pixels,
                     float* right_idepths, float* idepth_out, float* disparity_out, cudaStream_t stream);

// F.interpolate(bilinear, align_corners=False) of a (N, planes, h, w) float map, and the
// float->bilinear->(>0.5) mask variant (multi_view_stereonet.py:372-396).
int launch_upsample_f32(const float* in, int n_planes, int h, int w, int H, int W, float* out, cudaStream_t stream);
// packed: `out` is the bit volume (n_planes, H, ceil(W / 8)), numpy.packbits(axis=-1) layout; with h == H and
// w == W it only packs.
int launch_upsample_mask(const uint8_t* in, long long n_planes, int h, int w, int H, int W, uint8_t* out,
                         cudaStream_t stream, bool packed = false);

}  // namespace b200mvs
