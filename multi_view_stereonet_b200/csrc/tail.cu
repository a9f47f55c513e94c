// The 32 -> 1 convolutions that close IDepthmapRefiner (multi_view_stereonet.py:479-483) and CostVolumeFilter
// (:350-352), restructured as  pointwise 32 -> taps linear map  +  shifted sum over the taps:
//     out(p) = b + sum_t sum_c x_c(p + o_t) w[c][t] = b + sum_t P_t(p + o_t),      P_t(q) = sum_c x_c(q) w[c][t]
// Every input pixel is read and normalised once, its 9 (27) partial products are kept in shared memory (L2 for the
// 3-D filter), and an output is 9 (27) adds.  The weights travel as kernel parameters, so every FFMA takes its
// weight operand straight from the constant bank.  Both are pure bandwidth kernels: one read of the activation.
#include <cuda_fp16.h>

#include <cstdlib>

#include "conv.cuh"
#include "tail.cuh"

namespace b200mvs {
namespace {

// ------------------------------------------------------------------------------------------------------------
// IDepthmapRefiner tail:  x = lrelu(GN(y)) + resid ; delta = conv3x3(x, 32 -> 1) + b ;
//                         out = relu(prior * fx + delta) / fx          (:482 and the caller's scaling :607-611)
// ------------------------------------------------------------------------------------------------------------
constexpr int F_TH = 14, F_TW = 30;            // output tile; the halo-extended tile is 16 x 32 = 2 pixels per thread
constexpr int F_HH = F_TH + 2, F_HW = F_TW + 2;
constexpr int F_NT = 256;

struct FinalParams {
  const void* y;        // raw output of the last residual conv, [n][H][W][32] fp32 or fp16
  const void* resid;    // the residual stream, same type
  const double* stats;  // [n][4][2]
  const float* gamma;
  const float* beta;
  double inv_count;
  const float* prior;   // [n][H][W]
  const float* fx;
  int fx_div, fx_stride;
  float* out;           // [n][H][W]
  int H, W;
};

__device__ __forceinline__ void load32(const float* p, float* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
  }
}
__device__ __forceinline__ void load32(const __half* p, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p) + q);
    const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      v[8 * q + 2 * k] = f.x;
      v[8 * q + 2 * k + 1] = f.y;
    }
  }
}

template <class T>
__global__ void __launch_bounds__(F_NT) refine_final_kernel(const FinalParams p, const RefineFinalW W) {
  __shared__ float s_part[9][F_HH * F_HW];
  __shared__ float s_a[kC], s_b[kC];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, img = blockIdx.y;
  const int tiles_x = cdiv(p.W, F_TW);
  const int tx0 = (blockIdx.x % tiles_x) * F_TW, ty0 = (blockIdx.x / tiles_x) * F_TH;
  if (tid < kC) {
    const int grp = tid >> 3;
    const double sum = p.stats[(img * kGroups + grp) * 2 + 0];
    const double sq = p.stats[(img * kGroups + grp) * 2 + 1];
    const double mean = sum * p.inv_count;
    double var = sq * p.inv_count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = gn_rstd(var);
    s_a[tid] = (float)((double)p.gamma[tid] * rstd);
    s_b[tid] = (float)((double)p.beta[tid] - mean * (double)p.gamma[tid] * rstd);
  }
  __syncthreads();
  const size_t ibase = (size_t)img * p.H * p.W;
#pragma unroll
  for (int r = 0; r < F_HH * F_HW / F_NT; ++r) {
    const int hp = tid + r * F_NT;
    const int gy = ty0 - 1 + hp / F_HW, gx = tx0 - 1 + hp % F_HW;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
      const size_t off = (ibase + (size_t)gy * p.W + gx) * kC;
      float v[32], rs[32];
      load32(reinterpret_cast<const T*>(p.y) + off, v);
      load32(reinterpret_cast<const T*>(p.resid) + off, rs);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = lrelu(fmaf(v[c], s_a[c], s_b[c])) + rs[c];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = fmaf(x, W.w[c * 9 + t], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_part[t][hp] = acc[t];
  }
  __syncthreads();
  const float f = __ldg(p.fx + (size_t)(img / p.fx_div) * p.fx_stride);
  for (int o = tid; o < F_TH * F_TW; o += F_NT) {
    const int ly = o / F_TW, lx = o % F_TW;
    const int oy = ty0 + ly, ox = tx0 + lx;
    if (oy >= p.H || ox >= p.W) continue;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc += s_part[t][(ly + t / 3) * F_HW + lx + t % 3];
    const size_t opix = ibase + (size_t)oy * p.W + ox;
    const float v = acc + W.bias;
    const float scaled = __fmul_rn(__ldg(p.prior + opix), f);
    p.out[opix] = __fdiv_rn(fmaxf(__fadd_rn(scaled, v), 0.f), f);
  }
}

// ------------------------------------------------------------------------------------------------------------
// IDepthmapRefiner.conv0 at level 0: 4 planar inputs -> 32 channels, fp32 FFMA with constant-bank weights.
// CTA tile 8 x 32 pixels.
// ------------------------------------------------------------------------------------------------------------
constexpr int H_TH = 8, H_TW = 32;
struct HeadParams {
  const float* image;   // [n][3][H][W]
  const float* prior;   // [n][H][W]
  const float* fx;
  int fx_div, fx_stride;
  void* out;            // [n][H][W][32] fp16 or fp32
  double* out_stats;    // [n][4][2]
  int H, W;
};

// Thread = (4 vertically adjacent output pixels, one channel octet): the octet's 8 weights of a (channel, tap) are
// two shared-memory vector loads reused by 32 FFMAs; the four lanes of a quad cover a pixel's 32 channels, so stores
// are full 64-byte (fp16) or 128-byte (fp32) pixel records.
// NC = 4: image + prior * fx (the whole conv0).  NC = 3: image channels only (+ bias) -- the part of conv0 that does
// not depend on the coarser level's result, computed ahead of time (see refine_head_pre_kernel).
template <bool HALF, int NC>
__global__ void __launch_bounds__(H_TH * H_TW, 4) refine_head_l0_kernel(const HeadParams p,
                                                                      const __grid_constant__ RefineHeadW W) {
  __shared__ float s_in[NC][H_TH + 2][H_TW + 2];
  __shared__ __align__(16) float s_w[4 * 9][32];
  __shared__ float s_bias[32];
  __shared__ double s_stats[2 * kGroups];
  const int tid = threadIdx.x, img = blockIdx.y;
  for (int i = tid; i < 4 * 9 * 32; i += H_TH * H_TW) {
    const int o = i & 31, ct = i >> 5;        // ct = c * 9 + tap
    s_w[ct][o] = W.w[o * 36 + ct];
  }
  if (tid < 32) s_bias[tid] = W.bias[tid];
  if (tid < 2 * kGroups) s_stats[tid] = 0.0;
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_x = cdiv(p.W, H_TW);
  const int tx0 = (blockIdx.x % tiles_x) * H_TW, ty0 = (blockIdx.x / tiles_x) * H_TH;
  const size_t plane = (size_t)p.H * p.W;
  const float f = NC == 4 ? __ldg(p.fx + (size_t)(img / p.fx_div) * p.fx_stride) : 1.0f;
  {
    // all loads of the tile are issued before any is stored: one memory round trip per CTA
    constexpr int N_IN = NC * (H_TH + 2) * (H_TW + 2);
    constexpr int ITERS = (N_IN + H_TH * H_TW - 1) / (H_TH * H_TW);
    float v[ITERS];
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
      const int i = tid + k * H_TH * H_TW;
      const int x = i % (H_TW + 2), y = (i / (H_TW + 2)) % (H_TH + 2), c = i / ((H_TW + 2) * (H_TH + 2));
      const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
      v[k] = 0.f;
      if (i < N_IN && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
        v[k] = __ldg((c < 3 ? p.image + ((size_t)img * 3 + c) * plane : p.prior + (size_t)img * plane) + (size_t)gy * p.W + gx);
    }
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
      const int i = tid + k * H_TH * H_TW;
      const int x = i % (H_TW + 2), y = (i / (H_TW + 2)) % (H_TH + 2), c = i / ((H_TW + 2) * (H_TH + 2));
      if (i < N_IN) s_in[c][y][x] = c < 3 ? v[k] : __fmul_rn(v[k], f);
    }
  }
  __syncthreads();
  const int o8 = tid & 3, g = tid >> 2;
  const int lx = g % H_TW, ly0 = (g / H_TW) * 4;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = s_bias[8 * o8 + e];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float in[6][3];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int x = 0; x < 3; ++x) in[r][x] = s_in[c][ly0 + r][lx + x];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c * 9 + t][8 * o8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c * 9 + t][8 * o8 + 4]);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = in[k + t / 3][t % 3];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[k][e] = fmaf(x, w[e], acc[k][e]);
      }
    }
  }
  const int ox = tx0 + lx;
  float gs = 0.f, gq = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int oy = ty0 + ly0 + k;
    if (ox < p.W && oy < p.H) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gs += acc[k][e];
        gq = fmaf(acc[k][e], acc[k][e], gq);
      }
      const size_t opix = ((size_t)img * p.H + oy) * p.W + ox;
      if (HALF) {
        uint32_t h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2 hh = __floats2half2_rn(acc[k][2 * q], acc[k][2 * q + 1]);
          h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + opix * kC + 8 * o8) = make_uint4(h[0], h[1], h[2], h[3]);
      } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * kC + 8 * o8);
        if (NC == 3) {
          // the precomputed guide part (42 MB at 512x640) is read ~1.5 ms later: streaming stores, so that it does not
          // push the depth sweep's working set (feature volume, image half of conv0) out of L2 while the sweep runs
          __stcs(o, make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]));
          __stcs(o + 1, make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]));
        } else {
          o[0] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
          o[1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
        }
      }
    }
  }
  if (p.out_stats != nullptr) {
    // lanes with equal (lane & 3) hold the same GroupNorm group
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
      gq += __shfl_xor_sync(0xffffffffu, gq, o);
    }
    if ((tid & 31) < 4) {
      atomicAdd(&s_stats[2 * o8 + 0], (double)gs);
      atomicAdd(&s_stats[2 * o8 + 1], (double)gq);
    }
    __syncthreads();
    if (tid < 2 * kGroups) atomicAdd(p.out_stats + (size_t)img * 2 * kGroups + tid, s_stats[tid]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// IDepthmapRefiner.conv0 with the guide part precomputed.  conv0 is linear in its input channels:
//   conv0(cat[guide, idepth * fx]) = [conv(guide) + b] + conv(idepth * fx)
// The bracket (35 of the 36 input channels at levels 1-4, 3 of 4 at level 0) depends only on the reference image and
// its features, so it is computed on the side stream while the depth sweep runs (`pre`, fp32 channels-last).  What
// is left on the critical path when the coarser level's idepth arrives is this kernel: 9 taps of one planar
// channel per output, one read of `pre`, the GroupNorm statistics of the sum.  Same thread mapping as above.
// ------------------------------------------------------------------------------------------------------------
struct HeadPreParams {
  const float* pre;     // [n / pre_div][H][W][32] fp32
  int pre_div;
  const float* prior;   // [n][H][W], read when coarse == nullptr
  const float* coarse;  // [n][ch][cw]: the coarser level's idepth; the prior is its bilinear upsampling
  int ch, cw;           // (Upsampler.forward, multi_view_stereonet.py:372-380, 630), computed on the fly and
  float* prior_out;     // written to prior_out [n][H][W] for the tile interior ("left_idepthmap_raw_pyr")
  const float* fx;
  int fx_div, fx_stride;
  void* out;            // [n][H][W][32] fp16 or fp32
  double* out_stats;    // [n][4][2]
  int H, W;
};

template <bool HALF>
__global__ void __launch_bounds__(H_TH * H_TW, 4) refine_head_pre_kernel(const HeadPreParams p,
                                                                       const __grid_constant__ RefineHeadIdW W) {
  __shared__ float s_in[H_TH + 2][H_TW + 2];
  __shared__ __align__(16) float s_w[9][32];
  __shared__ double s_stats[2 * kGroups];
  const int tid = threadIdx.x, img = blockIdx.y;
  for (int i = tid; i < 9 * 32; i += H_TH * H_TW) s_w[i >> 5][i & 31] = W.w[i];
  if (tid < 2 * kGroups) s_stats[tid] = 0.0;
  pdl_launch_dependents();
  pdl_wait();
  const int tiles_x = cdiv(p.W, H_TW);
  const int tx0 = (blockIdx.x % tiles_x) * H_TW, ty0 = (blockIdx.x / tiles_x) * H_TH;
  const size_t plane = (size_t)p.H * p.W;
  const float f = __ldg(p.fx + (size_t)(img / p.fx_div) * p.fx_stride);
  const int o8 = tid & 3, g = tid >> 2;
  const int lx = g % H_TW, ly0 = (g / H_TW) * 4;
  const int ox = tx0 + lx;
  // every global load of the thread is in flight before the first use
  constexpr int N_IN = (H_TH + 2) * (H_TW + 2);
  constexpr int ITERS = (N_IN + H_TH * H_TW - 1) / (H_TH * H_TW);
  float v[ITERS];
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const int i = tid + k * H_TH * H_TW;
    const int x = i % (H_TW + 2), y = i / (H_TW + 2);
    const int gy = ty0 - 1 + y, gx = tx0 - 1 + x;
    v[k] = 0.f;
    if (i < N_IN && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
      if (p.coarse != nullptr) {
        v[k] = upsample_bilinear_at(p.coarse + (size_t)img * p.ch * p.cw, p.ch, p.cw, p.H, p.W, gy, gx);
        if (y >= 1 && y <= H_TH && x >= 1 && x <= H_TW) p.prior_out[(size_t)img * plane + (size_t)gy * p.W + gx] = v[k];
      } else {
        v[k] = __ldg(p.prior + (size_t)img * plane + (size_t)gy * p.W + gx);
      }
    }
  }
  float acc[4][8];
  const float* pre = p.pre + (size_t)(img / p.pre_div) * plane * kC + 8 * o8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int oy = ty0 + ly0 + k;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ox < p.W && oy < p.H) {
      const float4* src = reinterpret_cast<const float4*>(pre + ((size_t)oy * p.W + ox) * kC);
      a = __ldg(src);
      b = __ldg(src + 1);
    }
    acc[k][0] = a.x; acc[k][1] = a.y; acc[k][2] = a.z; acc[k][3] = a.w;
    acc[k][4] = b.x; acc[k][5] = b.y; acc[k][6] = b.z; acc[k][7] = b.w;
  }
#pragma unroll
  for (int k = 0; k < ITERS; ++k) {
    const int i = tid + k * H_TH * H_TW;
    if (i < N_IN) s_in[i / (H_TW + 2)][i % (H_TW + 2)] = __fmul_rn(v[k], f);
  }
  __syncthreads();
  {
    float in[6][3];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int x = 0; x < 3; ++x) in[r][x] = s_in[ly0 + r][lx + x];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[t][8 * o8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[t][8 * o8 + 4]);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = in[k + t / 3][t % 3];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[k][e] = fmaf(x, w[e], acc[k][e]);
      }
    }
  }
  float gs = 0.f, gq = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int oy = ty0 + ly0 + k;
    if (ox < p.W && oy < p.H) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gs += acc[k][e];
        gq = fmaf(acc[k][e], acc[k][e], gq);
      }
      const size_t opix = ((size_t)img * p.H + oy) * p.W + ox;
      if (HALF) {
        uint32_t h[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2 hh = __floats2half2_rn(acc[k][2 * q], acc[k][2 * q + 1]);
          h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + opix * kC + 8 * o8) = make_uint4(h[0], h[1], h[2], h[3]);
      } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * kC + 8 * o8);
        o[0] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
        o[1] = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
      }
    }
  }
  if (p.out_stats != nullptr) {
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
      gq += __shfl_xor_sync(0xffffffffu, gq, o);
    }
    if ((tid & 31) < 4) {
      atomicAdd(&s_stats[2 * o8 + 0], (double)gs);
      atomicAdd(&s_stats[2 * o8 + 1], (double)gq);
    }
    __syncthreads();
    if (tid < 2 * kGroups) atomicAdd(p.out_stats + (size_t)img * 2 * kGroups + tid, s_stats[tid]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// CostVolumeFilter tail: conv3d 3x3x3 32 -> 1 on lrelu(GN(y3)), then the soft-argmin over the hypothesis axis
// (extract_idepthmap, :486-492).
// ------------------------------------------------------------------------------------------------------------
struct CvfPartialParams {
  const float* y;       // [n][D][P][32]
  const double* stats;  // [n][4][2]
  const float* gamma;
  const float* beta;
  double inv_count;
  float* part;          // [n][D][27][P]
  int D, P;
};

__global__ void __launch_bounds__(128) cvf_final_partial_kernel(const CvfPartialParams p, const CvfFinalW W) {
  __shared__ float s_a[kC], s_b[kC];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, n = blockIdx.y;
  if (tid < kC) {
    const int grp = tid >> 3;
    const double sum = p.stats[(n * kGroups + grp) * 2 + 0];
    const double sq = p.stats[(n * kGroups + grp) * 2 + 1];
    const double mean = sum * p.inv_count;
    double var = sq * p.inv_count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = gn_rstd(var);
    s_a[tid] = (float)((double)p.gamma[tid] * rstd);
    s_b[tid] = (float)((double)p.beta[tid] - mean * (double)p.gamma[tid] * rstd);
  }
  __syncthreads();
  const int vox = blockIdx.x * blockDim.x + tid;   // d * P + pixel
  if (vox >= p.D * p.P) return;
  const int d = vox / p.P, pix = vox % p.P;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  const float* src = p.y + ((size_t)n * p.D * p.P + vox) * kC;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
    const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * q + e;
      const float x = lrelu(fmaf(v[e], s_a[c], s_b[c]));
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[t] = fmaf(x, W.w[c * 27 + t], acc[t]);
    }
  }
  float* dst = p.part + ((size_t)(n * p.D + d) * 27) * p.P + pix;
#pragma unroll
  for (int t = 0; t < 27; ++t) dst[(size_t)t * p.P] = acc[t];
}

// 16 consecutive pixels per CTA, 1024 threads: thread (pixel, slot) sums the 27 shifted partials of hypotheses
// slot, slot + 64, ... (all 27 loads of a hypothesis in flight at once), then 8 lanes per pixel run the soft-argmin
// exactly as softargmin_kernel.
constexpr int S_PX = 16, S_SLOTS = 64;
__global__ void __launch_bounds__(S_PX * S_SLOTS) cvf_final_sum_softargmin_kernel(const float* __restrict__ part,
                                                                                  const float* __restrict__ samples,
                                                                                  float bias, int D, int h, int w,
                                                                                  float* __restrict__ cost1,
                                                                                  float* __restrict__ raw) {
  extern __shared__ float s_cost[];   // [D][S_PX]
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int n = blockIdx.y, P = h * w;
  {
    const int lp = tid % S_PX, slot = tid / S_PX;
    const int pix = blockIdx.x * S_PX + lp;
    const bool okp = pix < P;
    const int y = okp ? pix / w : 0, x = okp ? pix % w : 0;
    for (int d = slot; d < D; d += S_SLOTS) {
      float t[27];
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        const int dd = d + kd - 1;
        const bool okd = okp && dd >= 0 && dd < D;
        const float* pl = part + ((size_t)(n * D + (okd ? dd : 0)) * 27 + kd * 9) * P;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int yy = y + ky - 1;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xx = x + kx - 1;
            const bool ok = okd && yy >= 0 && yy < h && xx >= 0 && xx < w;
            t[kd * 9 + ky * 3 + kx] = ok ? __ldg(pl + (size_t)(ky * 3 + kx) * P + yy * w + xx) : 0.f;
          }
        }
      }
      float acc = bias;
#pragma unroll
      for (int k = 0; k < 27; ++k) acc += t[k];
      if (okp) cost1[((size_t)n * D + d) * P + pix] = acc;
      s_cost[d * S_PX + lp] = acc;
    }
  }
  __syncthreads();
  if (tid >= S_PX * 8) return;
  const int sub = tid & 7, pl = tid >> 3;
  const int p2 = blockIdx.x * S_PX + pl;
  float m = -INFINITY;
  for (int d = sub; d < D; d += 8) m = fmaxf(m, -s_cost[d * S_PX + pl]);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float den = 0.f, num = 0.f;
  for (int d = sub; d < D; d += 8) {
    const float e = expf(-s_cost[d * S_PX + pl] - m);
    den += e;
    num += e * __ldg(samples + (size_t)n * D + d);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    den += __shfl_xor_sync(0xffffffffu, den, o);
    num += __shfl_xor_sync(0xffffffffu, num, o);
  }
  if (p2 < P && sub == 0) raw[(size_t)n * P + p2] = num / den;
}

}  // namespace

int launch_refine_final(const void* y, const void* resid, bool half_io, const double* stats, const float* gamma,
                        const float* beta, double inv_count, const RefineFinalW& w, const float* prior,
                        const float* fx, int fx_div, int fx_stride, int n, int H, int W, float* out,
                        cudaStream_t stream) {
  if (n <= 0) return 0;
  FinalParams p;
  p.y = y;
  p.resid = resid;
  p.stats = stats;
  p.gamma = gamma;
  p.beta = beta;
  p.inv_count = inv_count;
  p.prior = prior;
  p.fx = fx;
  p.fx_div = fx_div;
  p.fx_stride = fx_stride;
  p.out = out;
  p.H = H;
  p.W = W;
  dim3 grid(cdiv(W, F_TW) * cdiv(H, F_TH), n);
  if (half_io)
    launch_pdl(refine_final_kernel<__half>, grid, dim3(F_NT), (size_t)0, stream, p, w);
  else
    launch_pdl(refine_final_kernel<float>, grid, dim3(F_NT), (size_t)0, stream, p, w);
  B200MVS_LAUNCH_OK("refine_final_kernel");
  return 0;
}

int launch_refine_head_l0(const float* image, const float* prior, const float* fx, int fx_div, int fx_stride,
                          const RefineHeadW& w, int n, int H, int W, void* out, bool out_half, double* out_stats,
                          cudaStream_t stream) {
  if (n <= 0) return 0;
  HeadParams p;
  p.image = image;
  p.prior = prior;
  p.fx = fx;
  p.fx_div = fx_div;
  p.fx_stride = fx_stride;
  p.out = out;
  static const bool nostats = getenv("B200MVS_NOSTATS") != nullptr;   // timing experiment (wrong results)
  p.out_stats = nostats ? nullptr : out_stats;
  p.H = H;
  p.W = W;
  dim3 grid(cdiv(W, H_TW) * cdiv(H, H_TH), n);
  if (out_half)
    launch_pdl(refine_head_l0_kernel<true, 4>, grid, dim3(H_TH * H_TW), (size_t)0, stream, p, w);
  else
    launch_pdl(refine_head_l0_kernel<false, 4>, grid, dim3(H_TH * H_TW), (size_t)0, stream, p, w);
  B200MVS_LAUNCH_OK("refine_head_l0_kernel");
  return 0;
}

int launch_refine_head_image_l0(const float* image, const RefineHeadW& w, int n, int H, int W, float* pre,
                                cudaStream_t stream) {
  if (n <= 0) return 0;
  HeadParams p;
  p.image = image;
  p.prior = nullptr;
  p.fx = nullptr;
  p.fx_div = 1;
  p.fx_stride = 0;
  p.out = pre;
  p.out_stats = nullptr;
  p.H = H;
  p.W = W;
  dim3 grid(cdiv(W, H_TW) * cdiv(H, H_TH), n);
  launch_pdl(refine_head_l0_kernel<false, 3>, grid, dim3(H_TH * H_TW), (size_t)0, stream, p, w);
  B200MVS_LAUNCH_OK("refine_head_l0_kernel<image>");
  return 0;
}

int launch_refine_head_pre(const float* pre, int pre_div, const float* prior, const float* coarse, int ch, int cw,
                           float* prior_out, const float* fx, int fx_div, int fx_stride, const RefineHeadIdW& w, int n,
                           int H, int W, void* out, bool out_half, double* out_stats, cudaStream_t stream) {
  if (n <= 0) return 0;
  HeadPreParams p;
  p.pre = pre;
  p.pre_div = pre_div;
  p.prior = prior;
  p.coarse = coarse;
  p.ch = ch;
  p.cw = cw;
  p.prior_out = prior_out;
  p.fx = fx;
  p.fx_div = fx_div;
  p.fx_stride = fx_stride;
  p.out = out;
  p.out_stats = out_stats;
  p.H = H;
  p.W = W;
  dim3 grid(cdiv(W, H_TW) * cdiv(H, H_TH), n);
  if (out_half)
    launch_pdl(refine_head_pre_kernel<true>, grid, dim3(H_TH * H_TW), (size_t)0, stream, p, w);
  else
    launch_pdl(refine_head_pre_kernel<false>, grid, dim3(H_TH * H_TW), (size_t)0, stream, p, w);
  B200MVS_LAUNCH_OK("refine_head_pre_kernel");
  return 0;
}

bool cvf_final_supported(int D) { return (size_t)D * S_PX * sizeof(float) <= 48 * 1024; }

int launch_cvf_final(const float* y, const double* stats, const float* gamma, const float* beta, double inv_count,
                     const CvfFinalW& w, const float* samples, int n, int D, int h, int wd, float* part, float* cost1,
                     float* raw, cudaStream_t stream) {
  if (n <= 0) return 0;
  CvfPartialParams p;
  p.y = y;
  p.stats = stats;
  p.gamma = gamma;
  p.beta = beta;
  p.inv_count = inv_count;
  p.part = part;
  p.D = D;
  p.P = h * wd;
  launch_pdl(cvf_final_partial_kernel, dim3(cdiv(D * h * wd, 128), n), dim3(128), (size_t)0, stream, p, w);
  B200MVS_LAUNCH_OK("cvf_final_partial_kernel");
  launch_pdl(cvf_final_sum_softargmin_kernel, dim3(cdiv(h * wd, S_PX), n), dim3(S_PX * S_SLOTS),
             (size_t)D * S_PX * sizeof(float),
             stream, (const float*)part, samples, w.bias, D, h, wd, cost1, raw);
  B200MVS_LAUNCH_OK("cvf_final_sum_softargmin_kernel");
  return 0;
}

}  // namespace b200mvs
