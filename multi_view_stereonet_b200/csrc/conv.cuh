// Direct convolution kernels (fp32 FFMA) with fused GroupNorm prologue / statistics epilogue.
//
// Every convolution of the hot path is followed by GroupNorm(4, 32) whose statistics span the whole
// image (or whole D*h*w volume), so a layer cannot normalise its own output.  Each conv therefore
//   - applies the PREVIOUS layer's GroupNorm + LeakyReLU (+ residual) while staging its input tile
//     in shared memory (prologue), and
//   - accumulates sum / sum-of-squares of its own raw output per (image, group) (epilogue),
// so each layer reads its input once and writes its output once.
#pragma once
#include "common.cuh"

namespace b200mvs {

enum FeatMode : int {
  FEAT_NONE = 0,    // no 32-channel source
  FEAT_RAW = 1,     // x = src
  FEAT_GN = 2,      // x = lrelu(gn(src))
  FEAT_GN_RES = 3,  // x = lrelu(gn(src)) + resid            (utils/resnet.py:93-109)
};

// 32-channel channels-last source [n][D][H][W][32].
struct FeatSrc {
  const float* ptr = nullptr;
  int mode = FEAT_NONE;
  int img_div = 1;               // source image = img / img_div (a guide shared by V views)
  const double* stats = nullptr; // [img][4][2] sum, sum of squares of `ptr` (GN modes)
  const float* gamma = nullptr;
  const float* beta = nullptr;
  double inv_count = 0.0;        // 1 / (8 * D * H * W)
  const float* resid = nullptr;  // [img][D][H][W][32]
  float* x_out = nullptr;        // if set, the transformed input is written back (tile interior)
  // fp16 activation storage (refiner levels 0-2 on the tensor-core path): ptr, resid and x_out then point to
  // __half data of the same [n][H][W][32] shape.  Statistics always come from fp32 accumulators.
  int half_io = 0;
};

// Up to 4 planar single-channel sources (image planes, an idepth plane).
struct ExtraSrc {
  int n = 0;
  const float* ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  long long img_stride[4] = {0, 0, 0, 0};
  int img_div[4] = {1, 1, 1, 1};
  const float* scale[4] = {nullptr, nullptr, nullptr, nullptr};  // value *= scale[(img / scale_div) * scale_stride]
  int scale_div[4] = {1, 1, 1, 1};
  int scale_stride[4] = {1, 1, 1, 1};
};

struct ConvParams {
  FeatSrc feat;
  ExtraSrc extra;
  const float* w = nullptr;     // packed [chunk][tap][8][COUT]; chunks = feat (4) then extra (1)
  const float* bias = nullptr;  // [COUT] or null
  int n_img = 0;
  int Di = 1, Hi = 0, Wi = 0;
  int Do = 1, Ho = 0, Wo = 0;
  int dil = 1;
  float* out = nullptr;          // COUT=32: [img][Do][Ho][Wo][32]   COUT=1: [img][Do][Ho][Wo]
  long long out_img_stride = 0;  // elements between output images; 0 = dense (Do*Ho*Wo*COUT)
  int out_half = 0;              // COUT=32 on the tensor-core path: store the raw output as __half
  double* out_stats = nullptr;   // [img][4][2], COUT=32 only
  const float* add_src = nullptr;  // COUT=32: out += add_src (same layout)
  // COUT=1 epilogue: 0 -> acc + bias ; 1 -> relu(prior * fx + acc + bias) / fx
  int epi1_mode = 0;
  const float* prior = nullptr;  // [img][Ho][Wo]
  const float* fx = nullptr;     // fx[(img / fx_div) * fx_stride]
  int fx_div = 1, fx_stride = 1;
  int tag = 0;                   // kernel class for b200mvs_probe_select (host side only)
};

enum ConvTag : int { TAG_NONE = 0, TAG_REFINE_CONV32_L0 = 1, TAG_CVF_CONV32 = 2, TAG_RECURRENCE = 3 };

// Host hooks called immediately before / after a tagged launch (api.cu).
void probe_before(int tag, cudaStream_t stream);
void probe_after(int tag, cudaStream_t stream);

enum ConvKind : int { CONV_3x3 = 0, CONV_5x5_S2 = 1, CONV_3x3x3 = 2 };

// Launches the FFMA convolution.  cout is 32 or 1.
int launch_conv(ConvKind kind, int cout, const ConvParams& p, cudaStream_t stream);
int conv_init();  // raises the dynamic shared memory limits once per process

}  // namespace b200mvs
