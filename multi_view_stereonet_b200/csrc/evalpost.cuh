// Post-processing after the hot path (evalpost.cu): inverse depth -> depth, ground-truth range mask, depth metrics.
#pragma once
#include "common.cuh"

namespace b200mvs {

// est (batch, pixels): inverse depth maps in baseline-normalised units (est_is_depth = false; test.py:211-214) or
// depth maps used as they are (est_is_depth = true); baseline (batch) or null (= 1); depth_true (batch, pixels)
// baseline-normalised ground-truth depth (test.py:167-186) or null.  Optional outputs: idepth_out / depth_out
// (batch, pixels); metrics (batch, 8) float64 = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, valid-pixel count
// (test.py:41-71 over the pixels where ground truth and estimate lie strictly inside (min_depth, max_depth)).
int launch_depth_metrics(const float* est, const float* baseline, const float* depth_true, bool est_is_depth,
                         float min_depth, float max_depth, int batch, long long pixels, float* idepth_out,
                         float* depth_out, double* metrics, cudaStream_t stream);

}  // namespace b200mvs
