// tcgen05 Conv3d 3x3x3 32->32 for the cost-volume filter (cvf_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace b200mvs {

struct CvfArgs {
  const float* in = nullptr;      // [n][D][h][w][32] raw output of the previous layer (or the cost volume)
  int mode = 1;                   // FEAT_RAW or FEAT_GN
  const double* stats = nullptr;  // [n][4][2] statistics of `in` (FEAT_GN)
  const float* gamma = nullptr;
  const float* beta = nullptr;
  double inv_count = 0.0;
  const uint8_t* w16 = nullptr;   // pack_cvf_tc_weights
  const float* bias = nullptr;
  float* out = nullptr;           // [n][D][h][w][32]
  double* out_stats = nullptr;    // [n][4][2]
  int n = 0, D = 0, h = 0, w = 0;
  int tag = 0;
};

void pack_cvf_tc_weights(const float* w_oidhw, std::vector<uint8_t>* out);
bool cvf_tc_supported(int h, int w);
int launch_cvf_tc(const CvfArgs& a, cudaStream_t stream);

}  // namespace b200mvs
