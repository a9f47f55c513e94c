/*
 * b200mvs.h -- C ABI of the B200-native MultiViewStereoNet depth-inference hot path.
 *
 * The reference (robustrobotics/multi_view_stereonet) is pure Python over torch.nn and has no
 * FFI of its own; the drop-in boundary is the call `stereo_network(left_image_pyr, K_pyr,
 * T_right_in_left, right_image_pyr, num_idepth_samples, cost_volume_filter, refiners)` made at
 * multi_view_stereonet/multi_view_stereonet_utils.py:647-654 (signature at
 * multi_view_stereonet/multi_view_stereonet.py:538-545).  The entry points below are what a
 * binding for that call needs: plain pointers and sizes, no torch types.  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all tensors are contiguous float32 unless stated; images are NCHW (planar) exactly as the
 *     reference holds them; masks are one byte per element (0 / 1), True(1) = invalid, as in the
 *     reference (image_predictor.py:513-516).
 *   - every function returns 0 on success or a negative B200MVS_E* code; b200mvs_last_error()
 *     returns a thread-local message for the last failure on the calling thread.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised, unless
 *     the function name ends in _host.
 *   - a handle is bound to one CUDA device; one handle per device per process (one process per
 *     GPU).  A handle must not be used from two threads at once.
 */
#ifndef B200MVS_H_
#define B200MVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200MVS_API __attribute__((visibility("default")))
#else
#define B200MVS_API
#endif

#define B200MVS_OK 0
#define B200MVS_EINVAL (-1)   /* bad argument (the reference's `assert`s, e.g. multi_view_stereonet.py:548-549) */
#define B200MVS_ECUDA (-2)    /* CUDA runtime error */
#define B200MVS_EWEIGHTS (-3) /* state dict is missing a tensor or a tensor has the wrong size */
#define B200MVS_ENOMEM (-4)

#define B200MVS_NUM_LEVELS 5  /* multi_view_stereonet.py:504 */

typedef struct b200mvs_net b200mvs_net;

B200MVS_API const char* b200mvs_last_error(void);
B200MVS_API const char* b200mvs_version(void);

/* Builds a network on `device` from the reference's state dict (the 226 tensors named as in
 * multi_view_stereonet.py:506-532, e.g. "left_feature_extractor.conv0.weight"); replaces
 * `MultiViewStereoNet().load_state_dict(...)` + `.to(device)` (test.py:311-313).  `data[i]` are HOST
 * pointers in the reference's (O, I, k...) layout; the library repacks and uploads them. */
B200MVS_API int b200mvs_create(int device, int num_tensors, const char* const* names, const float* const* data,
                   const int64_t* numels, b200mvs_net** out);
B200MVS_API void b200mvs_destroy(b200mvs_net* net);

/* Shape of one call.  rows/cols are the level-0 image size; level l+1 is ((h+1)/2, (w+1)/2)
 * (utils/image_utils.py:111-128). */
typedef struct b200mvs_shape {
  int32_t batch;  /* B image groups */
  int32_t views;  /* V comparison views per group, len(T_right_in_lefts) */
  int32_t rows, cols;
  int32_t num_idepth_samples; /* D */
  int32_t do_cost_volume_filter;
  int32_t do_refiners[B200MVS_NUM_LEVELS];
} b200mvs_shape;

/* MultiViewStereoNet.forward (multi_view_stereonet.py:538-695) on DEVICE pointers.
 *   left_image_pyr[l]      (B,3,H_l,W_l)
 *   K_pyr[l]               (B,4,4)
 *   T_right_in_lefts[v]    (B,4,4)
 *   right_image_l0[v]      (B,3,H_0,W_0)   = right_image_pyrs[v][0]
 *   right_image_l4[v]      (B,3,H_4,W_4)   = right_image_pyrs[v][4]   (the only two levels the
 *                                            reference reads, multi_view_stereonet.py:255-257,275)
 * outputs, each level 0..4 (any pointer may be NULL to skip that output):
 *   out_idepth[l]          (B,1,H_l,W_l)   "left_idepthmap_pyr"
 *   out_idepth_raw[l]      (B,1,H_l,W_l)   "left_idepthmap_raw_pyr"
 *   out_mask[l]            (B,D,H_l,W_l) uint8  "left_idepthmap_mask_pyr"
 * Inputs are not modified. */
B200MVS_API int b200mvs_forward(b200mvs_net* net, const b200mvs_shape* shape, const float* const* left_image_pyr,
                    const float* const* K_pyr, const float* const* T_right_in_lefts,
                    const float* const* right_image_l0, const float* const* right_image_l4,
                    float* const* out_idepth, float* const* out_idepth_raw, uint8_t* const* out_mask,
                    void* stream);

/* Same call on HOST pointers (pinned or pageable): uploads the inputs, runs the path, downloads the
 * requested outputs and synchronises.  This is the end-to-end entry bench.py's `e2e` figure times.
 * `h2d_bytes` / `d2h_bytes` (may be NULL) receive the bytes copied in each direction. */
B200MVS_API int b200mvs_forward_host(b200mvs_net* net, const b200mvs_shape* shape, const float* const* left_image_pyr,
                         const float* const* K_pyr, const float* const* T_right_in_lefts,
                         const float* const* right_image_l0, const float* const* right_image_l4,
                         float* const* out_idepth, float* const* out_idepth_raw, uint8_t* const* out_mask,
                         int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Per-kernel-class device timing of the forwards issued since the last reset: when a class is
 * selected, every launch of that class is bracketed by CUDA events on the launch stream.  Classes:
 *   "refine_conv32_l0"  the 3x3 32->32 (dilated) convolutions of refiner0 at level 0 (6 per forward)
 *   "cvf_conv32"        the 3x3x3 32->32 convolutions of the cost-volume filter (4 per forward)
 *   "recurrence"        the persistent depth-sweep recurrence kernel (1 per forward)
 *   "none"              disable
 * b200mvs_probe_read synchronises the recorded events, returns their summed duration in ms and the
 * number of launches, and resets the accumulation. */
B200MVS_API int b200mvs_probe_select(b200mvs_net* net, const char* kernel_class);
B200MVS_API int b200mvs_probe_read(b200mvs_net* net, double* total_ms, int64_t* launches);

/* Number of kernels the last b200mvs_forward on this handle launched. */
B200MVS_API int64_t b200mvs_last_launch_count(const b200mvs_net* net);

/* With option "stage_profile" = 1 every forward records CUDA events at the stage boundaries of its main stream and
 * synchronises on the last one before returning (measurement mode: it ends the overlap of adjacent kernels at
 * those boundaries).  Returns "stage=us;...;total=us" of the last such forward (the reference's stages:
 * multi_view_stereonet.py:552-583 feature networks + warp, :279-290 depth sweep, :586-602 cost / filter /
 * soft-argmin, :605-627 level-4 refiner + view mean, :629-682 levels 3..0); the string lives until the next
 * forward on this handle. */
B200MVS_API const char* b200mvs_last_stage_profile(const b200mvs_net* net);

/* Stage access for parity tests: copies an internal buffer of the LAST forward into `dst` (DEVICE
 * pointer, `capacity` bytes) on `stream`; `*nbytes` receives its size.  Names:
 *   "idepth_samples" (B*V,D)  "baseline" (B*V)  "H0" (B*V,3,3)  "H" (B*V,D,3,3)  "H_inc" (B*V,D,3,3)
 *   "right_image0_warped" (B*V,3,H0,W0)   "l0_mask" (B*V,H0,W0) u8   "l4_mask" (B*V,D,h4,w4) u8
 *   "left_feature1".."left_feature4" (B,H_l,W_l,32 channels-last)
 *   "right_feature_volume" (B*V,D,h4,w4,32 channels-last, unmasked; only valid if kept, see
 *    b200mvs_set_debug)      "cost_filtered" (B*V,D,h4,w4)    "idepth4_raw_views" (B*V,h4,w4)
 *   "recurrence_flags" (B*V,17) i32: [16] != 0 when some tap of the sweep's incremental warp lies outside the shared-
 *    memory window of the CTA that gathers it (those taps are read from global memory behind the progress flags [0..15])
 * Image index n = b*V + v. */
B200MVS_API int b200mvs_get_stage(b200mvs_net* net, const char* name, void* dst, int64_t capacity, int64_t* nbytes,
                      void* stream);
/* keep_stages != 0 makes the forward preserve buffers it would otherwise overwrite in place. */
B200MVS_API int b200mvs_set_debug(b200mvs_net* net, int keep_stages);

/* Options: "tensor_cores" (default 1) -- run the 3x3 32->32 refiner convolutions of levels 0-2 on the
 * tcgen05 tensor cores with fp16 operands / fp32 accumulation; 0 keeps every layer on the fp32 path.
 * A/B switches of the schedule (all default 1, results unchanged up to float32 rounding): "warp_specialized",
 * "half_activations", "pdl", "overlap", "conv0_precompute", "left_late", "early_d2h", "l4_chain" (the level-4 tail of
 * the feature network as one cluster kernel per image); "lanes" (0 = automatic, the default: a call is split into two concurrent sub-batches when the last
 * round of the depth sweep's clusters would be nearly empty; 1 = never; 2 = whenever the sweep needs more than one round),
 * "sweep" (which kernel runs the depth sweep, multi_view_stereonet.py:279-290: 0 = automatic, the default -- the cluster
 * kernel while the chains of a call need at most two rounds of clusters, the wide kernel (one co-resident cooperative
 * grid for all chains, sweep_wide.cu) beyond that and for 1/16-scale images larger than one cluster, e.g. 1024x1280;
 * 1 = the wide kernel wherever it is supported; 2 = step by step, one launch per layer.  The wide kernel's CTAs wait
 * for each other: do not run two forwards that use it concurrently on one GPU),
 * "precise_refiners", "prio_main" (default 0: single-lane forwards on a high-priority stream of the library's own);
 * debugging: "recurrence_debug", "recurrence_profile" (phase totals and a per-warp timeline of one step, read back
 * with b200mvs_get_stage "recurrence_profile" / "recurrence_trace"), "stage_profile" (see
 * b200mvs_last_stage_profile). */
B200MVS_API int b200mvs_set_option(b200mvs_net* net, const char* name, int value);

/* Stage entry for kernel parity tests: y = conv3x3(x, dilation, padding = dilation) + bias on a
 * channels-last (n, rows, cols, 32) DEVICE tensor; `w_oihw_host` (32,32,3,3) and `bias_host` (32 or
 * NULL) are HOST pointers.  use_tensor_cores: 0 = fp32 FFMA kernel, 1 = tcgen05 with fp16 operands,
 * 2 = tcgen05 with split hi/lo fp16 operands (fp32-class accuracy).
 * Synchronises `stream` before returning. */
B200MVS_API int b200mvs_conv3x3_c32(const float* x, const float* w_oihw_host, const float* bias_host, int32_t n,
                                    int32_t rows, int32_t cols, int32_t dilation, int32_t use_tensor_cores, float* y,
                                    void* stream);

/* HomographyImagePredictor.forward (stereo/image_predictor.py:470-523) on DEVICE pointers:
 * H (N,3,3), image (N,C,rows,cols) -> pred (N,C,rows,cols), mask (N,rows,cols) uint8.
 * Needs no handle; `zero_invalid` != 0 additionally zeroes masked pixels as PlaneSweepWarper does
 * (multi_view_stereonet.py:230-233). */
B200MVS_API int b200mvs_homography_warp(const float* H, const float* image, int32_t n, int32_t channels, int32_t rows,
                            int32_t cols, int32_t zero_invalid, float* pred, uint8_t* mask, void* stream);

/* MaskUpsampler.forward (multi_view_stereonet.py:389-396) on DEVICE pointers, no handle: mask (planes, rows, cols)
 * uint8 0/1 -> float -> bilinear (align_corners=False) to (out_rows, out_cols) -> > 0.5.
 *   packed == 0: out (planes, out_rows, out_cols) uint8 0/1, as the reference returns it;
 *   packed != 0: out (planes, out_rows, ceil(out_cols / 8)) bits, bit 7 of a byte = its first pixel
 *                (numpy.packbits(axis=-1)); with out_rows == rows and out_cols == cols the call only packs.
 * The forward needs none of the mask volumes below level 4: a caller that passes NULL for
 * out_mask[0..3] to b200mvs_forward gets the level-4 volume only and produces the finer levels on demand with this
 * entry, one level at a time (each level is the upsampled *thresholded* coarser level, :630-673). */
B200MVS_API int b200mvs_upsample_mask(const uint8_t* mask, int64_t planes, int32_t rows, int32_t cols, int32_t out_rows,
                                      int32_t out_cols, int32_t packed, uint8_t* out, void* stream);

/* The reprojection layers of stereo/image_predictor.py next to the hot path, one fused per-pixel
 * kernel on DEVICE pointers; needs no handle.  K (N,4,4), T_right_in_left (N,4,4), map (N,1,rows,cols):
 *   map_kind 0: inverse depthmap   -> IDepthmapProjector.forward (:525-576) / IDepthImagePredictor.forward (:347-398)
 *   map_kind 1: disparity map      -> DisparityToIDepth.forward (:120-218) then the above = ImagePredictor.forward (:578-601)
 *   map_kind 2: rectified disparity-> RectifiedImagePredictor.forward (:275-345)
 * Outputs (any may be NULL): pred (N,C,rows,cols) sampled from right_image (N,C,rows,cols) with bilinear /
 * border / align_corners=False; mask (N,1,rows,cols) uint8, 1 = projects outside the image; right_pixels
 * (N,rows,cols,2) normalised grid coordinates; right_idepths (N,1,rows,cols); idepth_out (N,1,rows,cols) the
 * converted inverse depths (map_kind 1); disparity_out (N,1,rows,cols) = IDepthToDisparity.forward (:220-273) of the
 * inverse depths (map_kind 0 or 1). */
B200MVS_API int b200mvs_reproject(const float* K, const float* T_right_in_left, const float* map, int32_t map_kind,
                                  const float* right_image, int32_t n, int32_t channels, int32_t rows, int32_t cols,
                                  float* pred, uint8_t* mask, float* right_pixels, float* right_idepths,
                                  float* idepth_out, float* disparity_out, void* stream);

/* Input preparation before the hot path (multi_view_unpack_batch, multi_view_stereonet_utils.py:541-604), DEVICE
 * pointers, no handle.
 * b200mvs_area_downsample: one level of build_image_pyramid (utils/image_utils.py:111-128): `planes` images of
 * (rows, cols) -> ((rows+1)/2, (cols+1)/2) by area averaging (F.interpolate(mode="area")).
 * b200mvs_prepare_cameras: K (B,4,4); T_right_in_lefts[v] (B,4,4), v < views; level_sizes = HOST int32[2*levels]
 * (rows_l, cols_l) of the pyramid.  Writes K_pyr (levels,B,4,4) (:555-582), T_right_in_left / T_left_in_right
 * (views,B,4,4) with translations divided by the baseline to the first comparison camera (:585-604) and
 * baseline (B).  The caller checks baseline > 0 as the reference asserts (:598-599). */
B200MVS_API int b200mvs_area_downsample(const float* in, int32_t planes, int32_t rows, int32_t cols, float* out,
                                        void* stream);
B200MVS_API int b200mvs_prepare_cameras(const float* K, const float* const* T_right_in_lefts, int32_t batch,
                                        int32_t views, int32_t levels, const int32_t* level_sizes, float* K_pyr,
                                        float* T_right_in_left_out, float* T_left_in_right_out, float* baseline,
                                        void* stream);

/* Post-processing after the hot path, DEVICE pointers, no handle: what test.py does with
 * `left_idepthmap_pyr[0]` once the forward has returned, as one pass over the maps.
 *   est (batch, pixels)        est_is_depth == 0: inverse depth maps in baseline-normalised units; the function
 *                              divides by `baseline` and inverts the positive values (test.py:211-214).
 *                              est_is_depth != 0: depth maps, used as they are.
 *   baseline (batch)           inputs["baseline"] (multi_view_stereonet_utils.py:596-604); NULL = 1
 *   depth_true (batch, pixels) baseline-normalised ground-truth depth as multi_view_unpack_batch holds it; it is
 *                              multiplied back by the baseline as get_groundtruth_depthmap does (test.py:167-186).
 *                              NULL = conversion only.
 *   min_depth, max_depth       validity range, exclusive on both sides, applied to ground truth AND estimate
 *                              (test.py:222, 236): gta_sfm 0 / 1e3, demon 0.5 / 10.
 * outputs (any may be NULL):
 *   idepth_out, depth_out (batch, pixels)   batch_left_idepthmap_est / batch_left_depthmap_est (test.py:211-213)
 *   metrics (batch, 8) float64              abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 (get_depth_prediction_metrics,
 *                                           test.py:41-71) over the valid pixels, then the valid-pixel count; the
 *                                           seven metrics are NaN where the count is 0 (the reference skips such
 *                                           images, test.py:224-226). */
B200MVS_API int b200mvs_depth_metrics(const float* est, const float* baseline, const float* depth_true,
                                      int32_t est_is_depth, float min_depth, float max_depth, int32_t batch,
                                      int64_t pixels, float* idepth_out, float* depth_out, double* metrics,
                                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MVS_H_ */
