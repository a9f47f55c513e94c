// Persistent cluster kernel for the depth-sweep feature recurrence (recurrence.cu).
#pragma once
#include <vector>

#include "kernels.cuh"

namespace b200mvs {

struct RecurrenceArgs {
  float* vol;            // [n][D][rows*cols][32], hypothesis 0 filled; hypotheses 1..D-1 are written
  GeomOut geo;
  ViewPtrs right_l4;
  const uint8_t* w16;    // pack_recurrence_weights
  const float *bias0, *bias1, *bias2;
  const float *gamma0, *beta0, *gamma1, *beta1;
  const float* imgconv;  // [n][D][rows*cols][32] image half of conv0 + bias0 (launch_image_conv)
  const float* plan;     // [n][D][recurrence_plan_stride][4] gather plan of every step (launch_gather_plan)
  int* flags;            // [n][17] progress flags + "taps outside the window" flag (launch_gather_plan zeroes / sets)
  int n, D, rows, cols;
  int debug = 0;              // timing ablations, see recurrence.cu
  long long* prof = nullptr;  // optional [16 ranks][12 phases] cycle totals (debug builds of the tests)
};

// The level-4 tail of FeatureNetwork (six residual blocks + conv_final, multi_view_stereonet.py:119-127) as one
// cluster kernel per image (recurrence.cu: l4_tail_kernel); weights in the split-fp16 block format of conv_tc.cu.
struct L4TailArgs {
  const float* x0;        // [n][rows*cols][32]: conv3 output
  float* out;             // image i at out + i * out_stride
  long long out_stride;
  const uint8_t* w[7];    // res0..res5, conv_final
  const float* bias[7];
  const float* gamma[6];
  const float* beta[6];
  int n, rows, cols;
};
bool l4_tail_supported(int rows, int cols);
int launch_l4_tail(const L4TailArgs& a, cudaStream_t stream);

void pack_recurrence_weights(const float* w0_oihw35, const float* w1_oihw32, const float* w2_oihw32,
                             std::vector<uint8_t>* out);
// Gather plan of all steps (needs only the incremental homographies): floats per (image, hypothesis) = 4 * stride.
int recurrence_plan_stride(int rows, int cols);
int launch_gather_plan(const float* Hinc, int n, int D, int rows, int cols, float* plan, int* flags,
                       cudaStream_t stream);
bool recurrence_supported(int rows, int cols, int* n_tiles, size_t* smem_bytes);
int launch_recurrence(const RecurrenceArgs& a, cudaStream_t stream);
// The same recurrence for images beyond one cluster (sweep_wide.cu): one cooperative launch, up to five M-tiles per CTA,
// the chain's CTAs exchange statistics and boundary rows through L2 behind a counter barrier.  Uses the same
// weights, image half and gather plan as launch_recurrence.
struct WideScratch {
  float* wf = nullptr;       // sizes from sweep_wide_scratch
  float* y = nullptr;
  float* part = nullptr;     // float2 pairs
  unsigned* ctr = nullptr;   // chain counters + the watchdog's abort flag
  unsigned* abort_host = nullptr;   // pinned host word that receives the abort flag behind the sweep (optional)
};
bool sweep_wide_supported(int rows, int cols);
void sweep_wide_scratch(int rows, int cols, int n, size_t* wf_floats, size_t* y_floats, size_t* part_float2,
                        size_t* counters);
int launch_sweep_wide(const RecurrenceArgs& a, const WideScratch& s, cudaStream_t stream);
// How many of the kernel's clusters can be resident at once on the current device (cudaOccupancyMaxActiveClusters;
// B200: 7 clusters of 11 CTAs -- one GPC cannot take an 11-CTA cluster); 0 if the shape is not supported.
int recurrence_max_clusters(int rows, int cols);

}  // namespace b200mvs
