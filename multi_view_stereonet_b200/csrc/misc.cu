// Warps, cost build, soft-argmin, view reduction and upsampling kernels.
#include "kernels.cuh"

namespace b200mvs {
namespace {

// ---------------------------------------------------------------------------------------------
// HomographyImagePredictor.forward on planar (NCHW) images; stereo/image_predictor.py:470-523.
// One thread per output pixel, looping over channels; writes are coalesced along x.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_planar_kernel(const float* __restrict__ H, int h_stride, ViewPtrs src,
                                                          int channels, int rows, int cols, int zero_invalid,
                                                          float* __restrict__ pred, uint8_t* __restrict__ mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= rows * cols) return;
  const int x = p % cols, y = p / cols;
  const float* Hn = H + (size_t)n * h_stride;
  const WarpCoord c = homography_coord(Hn, (float)x, (float)y, rows, cols);
  const Bilinear bl = bilinear_setup(c, rows, cols);
  const size_t plane = (size_t)rows * cols;
  const float* img = src.p[n % src.views] + (size_t)(n / src.views) * channels * plane;
  if (mask != nullptr) mask[(size_t)n * plane + p] = c.invalid ? 1 : 0;
  const bool zero = zero_invalid && c.invalid;
  for (int ch = 0; ch < channels; ++ch) {
    const float* pl = img + ch * plane;
    float v = 0.f;
    if (!zero) {
      v = __ldg(pl + (size_t)bl.y0 * cols + bl.x0) * bl.w00 + __ldg(pl + (size_t)bl.y0 * cols + bl.x1) * bl.w01 +
          __ldg(pl + (size_t)bl.y1 * cols + bl.x0) * bl.w10 + __ldg(pl + (size_t)bl.y1 * cols + bl.x1) * bl.w11;
    }
    pred[((size_t)n * channels + ch) * plane + p] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Image part of FeatureRefiner.conv0 for every hypothesis at once.  conv0 acts on cat([warped 1/16 image (3),
// warped features (32)]) (multi_view_stereonet.py:425-426); convolution is linear in its input channels, and the
// image half depends only on H_d (multi_view_stereonet.py:270-275), not on the recurrence.  This kernel computes
//   out[n][d] = conv3x3(warp(right_l4, H_d) zeroed outside, W0[:, 0:3]) + bias0        for d = 1..D-1
// so that the persistent recurrence kernel only runs the 32-channel half on the tensor core.  fp32 FFMA.
// ---------------------------------------------------------------------------------------------
constexpr int kIcRows = 8;
__global__ void __launch_bounds__(256) image_conv_kernel(const float* __restrict__ H, ViewPtrs right_l4,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         int D, int rows, int cols, int oct_major,
                                                         float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_ic[];
  const int PWc = cols + 2;
  float* s_img = s_ic;                              // [3][kIcRows + 2][PWc]
  float* s_w = s_ic + 3 * (kIcRows + 2) * PWc;      // [tap][3][32]
  float* s_b = s_w + 27 * 32;
  const int tid = threadIdx.x;
  const int d = blockIdx.y + 1, n = blockIdx.z, y0 = blockIdx.x * kIcRows;
  const int pixels = rows * cols;
  for (int i = tid; i < 27 * 32; i += 256) {
    const int tap = i / 96, c = (i / 32) % 3, o = i % 32;
    s_w[i] = __ldg(w + (tap * 8 + c) * 32 + o);     // packed [tap][8][32], image channels are k = 0..2
  }
  if (tid < 32) s_b[tid] = __ldg(bias + tid);
  const float* Hd = H + ((size_t)n * D + d) * 9;
  const float* img = right_l4.p[n % right_l4.views] + (size_t)(n / right_l4.views) * 3 * pixels;
  for (int i = tid; i < (kIcRows + 2) * PWc; i += 256) {
    const int iy = i / PWc, ix = i % PWc;
    const int gy = y0 - 1 + iy, gx = ix - 1;
    float v[3] = {0.f, 0.f, 0.f};
    if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
      const WarpCoord c = homography_coord(Hd, (float)gx, (float)gy, rows, cols);
      if (!c.invalid) {
        const Bilinear b = bilinear_setup(c, rows, cols);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float* pl = img + (size_t)ch * pixels;
          v[ch] = __ldg(pl + b.y0 * cols + b.x0) * b.w00 + __ldg(pl + b.y0 * cols + b.x1) * b.w01 +
                  __ldg(pl + b.y1 * cols + b.x0) * b.w10 + __ldg(pl + b.y1 * cols + b.x1) * b.w11;
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) s_img[(ch * (kIcRows + 2) + iy) * PWc + ix] = v[ch];
  }
  __syncthreads();
  for (int t = tid; t < kIcRows * cols * 4; t += 256) {
    const int oct = t & 3, pix = t >> 2;
    const int y = pix / cols, x = pix % cols;
    if (y0 + y >= rows) continue;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = s_b[oct * 8 + k];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = s_img[(c * (kIcRows + 2) + y + tap / 3) * PWc + x + tap % 3];
        const float* wr = s_w + (tap * 3 + c) * 32 + oct * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(a, wr[k], acc[k]);
      }
    }
    // oct_major: [n][D][octet][pixel][8] -- the wide sweep's epilogue warps (lane = position, warp = octet) then read
    // 1 KB contiguous per 32 positions instead of 32 bytes from each of 32 lines
    const size_t opix = (size_t)(y0 + y) * cols + x;
    float4* o = reinterpret_cast<float4*>(
        out + (oct_major ? ((((size_t)n * D + d) * 4 + oct) * pixels + opix) * 8 : (((size_t)n * D + d) * pixels + opix) * kC + oct * 8));
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---------------------------------------------------------------------------------------------
// Input preparation: what multi_view_unpack_batch does before the hot path
// (multi_view_stereonet_utils.py:541-604).
// ---------------------------------------------------------------------------------------------
// One level of build_image_pyramid (utils/image_utils.py:111-128): F.interpolate(mode="area") to
// ((h+1)/2, (w+1)/2) = adaptive average pooling with windows [floor(i*in/out), ceil((i+1)*in/out)).
__global__ void __launch_bounds__(256) area_downsample_kernel(const float* __restrict__ in, int rows, int cols,
                                                              int orows, int ocols, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= orows * ocols) return;
  const int ox = o % ocols, oy = o / ocols;
  const int y0 = (int)(((long long)oy * rows) / orows), y1 = (int)((((long long)oy + 1) * rows + orows - 1) / orows);
  const int x0 = (int)(((long long)ox * cols) / ocols), x1 = (int)((((long long)ox + 1) * cols + ocols - 1) / ocols);
  const float* src = in + (size_t)plane * rows * cols;
  float sum = 0.f;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) sum = __fadd_rn(sum, __ldg(src + (size_t)y * cols + x));
  out[(size_t)plane * orows * ocols + o] = __fdiv_rn(sum, (float)((y1 - y0) * (x1 - x0)));
}

// Per image group: the per-level intrinsics (:555-582), T_left_in_right = inverse(T_right_in_left) (:590) and the
// normalisation of all translations by the baseline to the first comparison camera (:596-604).
__global__ void __launch_bounds__(32) prepare_cameras_kernel(const float* __restrict__ K, ViewPtrs T, int batch,
                                                             int levels, const int* __restrict__ level_sizes,
                                                             float* __restrict__ K_pyr, float* __restrict__ T_norm,
                                                             float* __restrict__ Tinv_norm,
                                                             float* __restrict__ baseline) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float* Kb = K + (size_t)b * 16;
  for (int l = 0; l < levels; ++l) {
    float* Ko = K_pyr + ((size_t)l * batch + b) * 16;
    for (int i = 0; i < 16; ++i) Ko[i] = Kb[i];
    if (l > 0) {
      const float sx = (float)level_sizes[2 * l + 1] / (float)level_sizes[1];
      const float sy = (float)level_sizes[2 * l] / (float)level_sizes[0];
      Ko[0] = __fmul_rn(Kb[0], sx);
      Ko[5] = __fmul_rn(Kb[5], sy);
      Ko[2] = __fsub_rn(__fmul_rn(sx, __fadd_rn(Kb[2], 0.5f)), 0.5f);
      Ko[6] = __fsub_rn(__fmul_rn(sy, __fadd_rn(Kb[6], 0.5f)), 0.5f);
    }
  }
  const float* T0 = T.p[0] + (size_t)b * 16;
  const float bl = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(T0[3], T0[3]), __fmul_rn(T0[7], T0[7])), __fmul_rn(T0[11], T0[11])));
  baseline[b] = bl;
  for (int v = 0; v < T.views; ++v) {
    const float* Tv = T.p[v] + (size_t)b * 16;
    // general 4x4 inverse (the reference calls torch.inverse), Gauss-Jordan with partial pivoting in float64
    double m[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        m[i][j] = (double)Tv[i * 4 + j];
        m[i][4 + j] = (i == j) ? 1.0 : 0.0;
      }
    for (int c = 0; c < 4; ++c) {
      int piv = c;
      double best = fabs(m[c][c]);
      for (int r = c + 1; r < 4; ++r)
        if (fabs(m[r][c]) > best) {
          best = fabs(m[r][c]);
          piv = r;
        }
      if (piv != c)
        for (int j = 0; j < 8; ++j) {
          const double tmp = m[c][j];
          m[c][j] = m[piv][j];
          m[piv][j] = tmp;
        }
      const double dd = 1.0 / m[c][c];
      for (int j = 0; j < 8; ++j) m[c][j] *= dd;
      for (int r = 0; r < 4; ++r)
        if (r != c) {
          const double f = m[r][c];
          for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    float* To = T_norm + ((size_t)v * batch + b) * 16;
    float* Io = Tinv_norm + ((size_t)v * batch + b) * 16;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        float t = Tv[i * 4 + j], ti = (float)m[i][4 + j];
        if (j == 3 && i < 3) {
          t = __fdiv_rn(t, bl);
          ti = __fdiv_rn(ti, bl);
        }
        To[i * 4 + j] = t;
        Io[i * 4 + j] = ti;
      }
  }
}

// ---------------------------------------------------------------------------------------------
// One recurrence step's warps (multi_view_stereonet.py:275, 285).  8 lanes per pixel, one float4
// of the 32 feature channels each; lanes 0..2 additionally warp one plane of the 1/16 right image.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) step_warp_kernel(const float* __restrict__ vol, GeomOut geo, ViewPtrs right_l4,
                                                        int D, int step, int rows, int cols,
                                                        float* __restrict__ wf, float* __restrict__ wimg) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pixels = rows * cols;
  const int p = i >> 3, q = i & 7;
  if (p >= pixels) return;
  const int x = p % cols, y = p / cols;
  {
    const float* Hinc = geo.Hinc + ((size_t)n * D + step) * 9;
    const WarpCoord c = homography_coord(Hinc, (float)x, (float)y, rows, cols);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!c.invalid) {
      const Bilinear bl = bilinear_setup(c, rows, cols);
      const float* prev = vol + ((size_t)n * D + (step - 1)) * pixels * kC + 4 * q;
      const float4 a = __ldg(reinterpret_cast<const float4*>(prev + ((size_t)bl.y0 * cols + bl.x0) * kC));
      const float4 b = __ldg(reinterpret_cast<const float4*>(prev + ((size_t)bl.y0 * cols + bl.x1) * kC));
      const float4 cc = __ldg(reinterpret_cast<const float4*>(prev + ((size_t)bl.y1 * cols + bl.x0) * kC));
      const float4 d = __ldg(reinterpret_cast<const float4*>(prev + ((size_t)bl.y1 * cols + bl.x1) * kC));
      v.x = a.x * bl.w00 + b.x * bl.w01 + cc.x * bl.w10 + d.x * bl.w11;
      v.y = a.y * bl.w00 + b.y * bl.w01 + cc.y * bl.w10 + d.y * bl.w11;
      v.z = a.z * bl.w00 + b.z * bl.w01 + cc.z * bl.w10 + d.z * bl.w11;
      v.w = a.w * bl.w00 + b.w * bl.w01 + cc.w * bl.w10 + d.w * bl.w11;
    }
    *reinterpret_cast<float4*>(wf + ((size_t)n * pixels + p) * kC + 4 * q) = v;
  }
  if (q < 3) {
    const float* Hd = geo.H + ((size_t)n * D + step) * 9;
    const WarpCoord c = homography_coord(Hd, (float)x, (float)y, rows, cols);
    float v = 0.f;
    if (!c.invalid) {
      const Bilinear bl = bilinear_setup(c, rows, cols);
      const float* pl = right_l4.p[n % right_l4.views] + ((size_t)(n / right_l4.views) * 3 + q) * pixels;
      v = __ldg(pl + bl.y0 * cols + bl.x0) * bl.w00 + __ldg(pl + bl.y0 * cols + bl.x1) * bl.w01 +
          __ldg(pl + bl.y1 * cols + bl.x0) * bl.w10 + __ldg(pl + bl.y1 * cols + bl.x1) * bl.w11;
    }
    wimg[((size_t)n * 3 + q) * pixels + p] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// cost = valid ? |L - R| : 0   and the mask volume  (multi_view_stereonet.py:293-298, 586-592).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cost_kernel(const float* __restrict__ left, const float* vol,
                                                   const float* __restrict__ H, int views, int D, int rows, int cols,
                                                   float* cost, uint8_t* __restrict__ mask) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.z, d = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pixels = rows * cols;
  const int p = i >> 3, q = i & 7;
  if (p >= pixels) return;
  const WarpCoord c = homography_coord(H + ((size_t)n * D + d) * 9, (float)(p % cols), (float)(p / cols), rows, cols);
  const size_t o = (((size_t)n * D + d) * pixels + p) * kC + 4 * q;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!c.invalid) {
    const float4 l = __ldg(reinterpret_cast<const float4*>(left + ((size_t)(n / views) * pixels + p) * kC + 4 * q));
    const float4 v = *reinterpret_cast<const float4*>(vol + o);
    r.x = fabsf(l.x - v.x);
    r.y = fabsf(l.y - v.y);
    r.z = fabsf(l.z - v.z);
    r.w = fabsf(l.w - v.w);
  }
  *reinterpret_cast<float4*>(cost + o) = r;
  if (q == 0) mask[((size_t)n * D + d) * pixels + p] = c.invalid ? 1 : 0;
}

__global__ void __launch_bounds__(256) cost_norm_kernel(const float* __restrict__ cost, long long voxels,
                                                        float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= voxels) return;
  const float4* c = reinterpret_cast<const float4*>(cost + i * kC);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kC / 4; ++k) {
    const float4 v = __ldg(c + k);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  out[i] = sqrtf(s);
}

// ---------------------------------------------------------------------------------------------
// Soft-argmin over the hypothesis axis (multi_view_stereonet.py:486-492).  One thread per pixel;
// consecutive threads read consecutive pixels of each hypothesis plane.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) softargmin_kernel(const float* __restrict__ cost,
                                                         const float* __restrict__ samples, int D, int pixels,
                                                         float* __restrict__ raw) {
  pdl_launch_dependents();
  pdl_wait();
  // 8 lanes per pixel share the hypothesis axis (lane s handles d = s, s + 8, ...); 16 pixels per CTA.
  const int n = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const int p = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const bool ok = p < pixels;
  const float* c = cost + (size_t)n * D * pixels + (ok ? p : 0);
  float m = -INFINITY;
  for (int d = sub; d < D; d += 8) m = fmaxf(m, -__ldg(c + (size_t)d * pixels));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float den = 0.f, num = 0.f;
  for (int d = sub; d < D; d += 8) {
    const float e = expf(-__ldg(c + (size_t)d * pixels) - m);
    den += e;
    num += e * __ldg(samples + (size_t)n * D + d);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    den += __shfl_xor_sync(0xffffffffu, den, o);
    num += __shfl_xor_sync(0xffffffffu, num, o);
  }
  if (ok && sub == 0) raw[(size_t)n * pixels + p] = num / den;
}

// ---------------------------------------------------------------------------------------------
// Baseline un-normalisation per view, mean over views, mask vote (multi_view_stereonet.py:616-627).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) view_reduce_kernel(const float* __restrict__ raw_views,
                                                          const float* __restrict__ refined_views,
                                                          const uint8_t* __restrict__ mask_views,
                                                          const float* __restrict__ baseline, int views, int D,
                                                          int pixels, int alias, float* __restrict__ raw4,
                                                          float* __restrict__ idepth4, uint8_t* __restrict__ mask4) {
  pdl_launch_dependents();
  pdl_wait();
  // blockIdx.z = 0: the two idepth maps; blockIdx.z = 1 + d: mask plane d
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  const float nv = (float)views;
  if (blockIdx.z == 0) {
    float rs = 0.f, is = 0.f;
    for (int v = 0; v < views; ++v) {
      const int n = b * views + v;
      const float bl = __ldg(baseline + n);
      float r = __ldg(raw_views + (size_t)n * pixels + p);
      float f;
      if (alias) {
        // do_refiners[4] == False: the reference's two in-place divisions hit the same tensor
        // (multi_view_stereonet.py:613-619).
        r = __fdiv_rn(__fdiv_rn(r, bl), bl);
        f = r;
      } else {
        r = __fdiv_rn(r, bl);
        f = __fdiv_rn(__ldg(refined_views + (size_t)n * pixels + p), bl);
      }
      rs += r;
      is += f;
    }
    if (raw4 != nullptr) raw4[(size_t)b * pixels + p] = __fdiv_rn(rs, nv);
    idepth4[(size_t)b * pixels + p] = __fdiv_rn(is, nv);
  } else {
    const int d = blockIdx.z - 1;
    float ms = 0.f;
    for (int v = 0; v < views; ++v) ms += (float)__ldg(mask_views + ((size_t)(b * views + v) * D + d) * pixels + p);
    mask4[((size_t)b * D + d) * pixels + p] = (__fdiv_rn(ms, nv) > 0.5f) ? 1 : 0;
  }
}

// The same vote straight from the plane-sweep homographies (the mask of a view is where its homography leaves the
// image, multi_view_stereonet.py:293-298): the mask volumes depend on the cameras only, so their whole chain can run
// next to the depth sweep instead of next to the refiners.
__global__ void __launch_bounds__(128) mask_vote_kernel(const float* __restrict__ H, int views, int D, int rows, int cols,
                                                        uint8_t* __restrict__ mask4) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, d = blockIdx.z;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int pixels = rows * cols;
  if (p >= pixels) return;
  float ms = 0.f;
  for (int v = 0; v < views; ++v) {
    const WarpCoord c = homography_coord(H + ((size_t)(b * views + v) * D + d) * 9, (float)(p % cols), (float)(p / cols), rows, cols);
    ms += c.invalid ? 1.f : 0.f;
  }
  mask4[((size_t)b * D + d) * pixels + p] = (__fdiv_rn(ms, (float)views) > 0.5f) ? 1 : 0;
}

__global__ void __launch_bounds__(256) upsample_f32_kernel(const float* __restrict__ in, int h, int w, int H, int W,
                                                           float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int x = p % W, y = p / W;
  out[(size_t)plane * H * W + p] = upsample_bilinear_at(in + (size_t)plane * h * w, h, w, H, W, y, x);
}

// Mask volume upsampling: float(mask) -> bilinear -> > 0.5 (multi_view_stereonet.py:389-396).
// Each thread produces VEC horizontally adjacent outputs (one 32- or 128-bit store) when W % VEC == 0.
template <int VEC>
__global__ void __launch_bounds__(256) upsample_mask_kernel(const uint8_t* __restrict__ in, int h, int w, int H, int W,
                                                            uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long plane = blockIdx.y;
  const int wv = W / VEC;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * wv) return;
  const int xv = i % wv, y = i / wv;
  const Lerp ly = lerp_setup(y, (float)h / (float)H, h);
  const uint8_t* r0 = in + (size_t)plane * h * w + (size_t)ly.i0 * w;
  const uint8_t* r1 = in + (size_t)plane * h * w + (size_t)ly.i1 * w;
  uint32_t packed[(VEC + 3) / 4];
#pragma unroll
  for (int k = 0; k < (VEC + 3) / 4; ++k) packed[k] = 0;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const Lerp lx = lerp_setup(xv * VEC + k, (float)w / (float)W, w);
    const float v00 = (float)__ldg(r0 + lx.i0), v01 = (float)__ldg(r0 + lx.i1);
    const float v10 = (float)__ldg(r1 + lx.i0), v11 = (float)__ldg(r1 + lx.i1);
    const float v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
    packed[k >> 2] |= (v > 0.5f ? 1u : 0u) << (8 * (k & 3));
  }
  uint8_t* dst = out + (size_t)plane * H * W + (size_t)y * W + xv * VEC;
  if (VEC == 16) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  } else if (VEC == 4) {
    *reinterpret_cast<uint32_t*>(dst) = packed[0];
  } else {
    dst[0] = (uint8_t)packed[0];
  }
}

// Exact 2x upsampling (H = 2h, W = 2w, every level of an even-sized pyramid).  With scale 1/2 the source coordinate of
// output x is x/2 - 1/4: the four taps carry the weights {3/4, 1/4} x {3/4, 1/4} (all exact in float32), so for 0/1
// inputs the interpolated value exceeds 1/2 exactly when the NEAREST input pixel is set (9/16 alone, at most 7/16
// without it; at the border the clamped coordinate puts weight 1 on it).  The bilinear-then-threshold of
// multi_view_stereonet.py:389-396 is therefore a 2x2 replication -- bit-identical to upsample_mask_kernel (tested) --
// and the kernel is a byte expander at HBM speed: one thread reads 8 input bytes and writes 16 bytes to each of the
// two output rows (the general kernel issues four byte loads per output pixel: 0.6 TB/s).
__global__ void __launch_bounds__(256) upsample_mask2x_kernel(const uint8_t* __restrict__ in, int h, int w,
                                                              uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long plane = blockIdx.y;
  const int w8 = w >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w8) return;
  const int xv = i % w8, y = i / w8;
  const uint2 s = __ldg(reinterpret_cast<const uint2*>(in + (size_t)plane * h * w + (size_t)y * w) + xv);
  uint4 o;
  o.x = __byte_perm(s.x, 0, 0x1100);
  o.y = __byte_perm(s.x, 0, 0x3322);
  o.z = __byte_perm(s.y, 0, 0x1100);
  o.w = __byte_perm(s.y, 0, 0x3322);
  const int W = 2 * w;
  uint8_t* dst = out + (size_t)plane * 4 * h * w + (size_t)(2 * y) * W + (size_t)xv * 16;
  __stcs(reinterpret_cast<uint4*>(dst), o);        // written once, read by nobody on the device at the finest level
  __stcs(reinterpret_cast<uint4*>(dst + W), o);
}

// The same upsampling written as a bit volume: out (planes, H, ceil(W / 8)), bit 7 of a byte = its first pixel
// (numpy.packbits(mask, axis=-1)).  One thread = one output byte.
__global__ void __launch_bounds__(256) upsample_mask_packed_kernel(const uint8_t* __restrict__ in, int h, int w, int H,
                                                                   int W, uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long plane = blockIdx.y;
  const int wb = (W + 7) / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * wb) return;
  const int xb = i % wb, y = i / wb;
  const Lerp ly = lerp_setup(y, (float)h / (float)H, h);
  const uint8_t* r0 = in + (size_t)plane * h * w + (size_t)ly.i0 * w;
  const uint8_t* r1 = in + (size_t)plane * h * w + (size_t)ly.i1 * w;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = xb * 8 + k;
    if (x < W) {
      const Lerp lx = lerp_setup(x, (float)w / (float)W, w);
      const float v00 = (float)__ldg(r0 + lx.i0), v01 = (float)__ldg(r0 + lx.i1);
      const float v10 = (float)__ldg(r1 + lx.i0), v11 = (float)__ldg(r1 + lx.i1);
      const float v = ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
      bits |= (v > 0.5f ? 1u : 0u) << (7 - k);
    }
  }
  out[(size_t)plane * H * wb + (size_t)y * wb + xb] = (uint8_t)bits;
}

// Dense (planes, H, W) bytes -> bits, same layout as above (for the level whose dense volume already exists).
__global__ void __launch_bounds__(256) pack_mask_kernel(const uint8_t* __restrict__ in, int H, int W,
                                                        uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long plane = blockIdx.y;
  const int wb = (W + 7) / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * wb) return;
  const int xb = i % wb, y = i / wb;
  const uint8_t* r = in + (size_t)plane * H * W + (size_t)y * W;
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = xb * 8 + k;
    if (x < W && __ldg(r + x) != 0) bits |= 1u << (7 - k);
  }
  out[(size_t)plane * H * wb + (size_t)y * wb + xb] = (uint8_t)bits;
}

}  // namespace

int launch_warp_planar(const float* H, int h_stride, const ViewPtrs& src, int n, int channels, int rows, int cols,
                       bool zero_invalid, float* pred, uint8_t* mask, cudaStream_t stream) {
  if (n <= 0) return 0;
  dim3 grid(cdiv(rows * cols, 256), n);
  launch_pdl(warp_planar_kernel, grid, dim3(256), (size_t)0, stream, H, h_stride, src, channels, rows, cols, zero_invalid ? 1 : 0, pred,
                                               mask);
  B200MVS_LAUNCH_OK("warp_planar_kernel");
  return 0;
}

int launch_image_conv(const float* H, const ViewPtrs& right_l4, const float* w_tap8x32, const float* bias, int n,
                      int D, int rows, int cols, float* out, cudaStream_t stream, bool oct_major) {
  if (D < 2) return 0;
  dim3 grid(cdiv(rows, kIcRows), D - 1, n);
  const size_t smem = ((size_t)3 * (kIcRows + 2) * (cols + 2) + 27 * 32 + 32) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("launch_image_conv: image too wide");
    return -1;
  }
  launch_pdl(image_conv_kernel, grid, dim3(256), smem, stream, H, right_l4, w_tap8x32, bias, D, rows, cols,
             oct_major ? 1 : 0, out);
  B200MVS_LAUNCH_OK("image_conv_kernel");
  return 0;
}

int launch_area_downsample(const float* in, int planes, int rows, int cols, float* out, cudaStream_t stream) {
  if (planes <= 0) return 0;
  const int orows = (rows + 1) / 2, ocols = (cols + 1) / 2;
  for (int p0 = 0; p0 < planes; p0 += 65535) {
    const int np = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid(cdiv(orows * ocols, 256), np);
    launch_pdl(area_downsample_kernel, grid, dim3(256), (size_t)0, stream, in + (size_t)p0 * rows * cols, rows, cols, orows,
               ocols, out + (size_t)p0 * orows * ocols);
    B200MVS_LAUNCH_OK("area_downsample_kernel");
  }
  return 0;
}

int launch_prepare_cameras(const float* K, const ViewPtrs& T, int batch, int levels, const int* level_sizes_dev,
                           float* K_pyr, float* T_norm, float* Tinv_norm, float* baseline, cudaStream_t stream) {
  if (batch <= 0) return 0;
  launch_pdl(prepare_cameras_kernel, dim3(cdiv(batch, 32)), dim3(32), (size_t)0, stream, K, T, batch, levels,
             level_sizes_dev, K_pyr, T_norm, Tinv_norm, baseline);
  B200MVS_LAUNCH_OK("prepare_cameras_kernel");
  return 0;
}

int launch_step_warp(const float* vol, const GeomOut& geo, const ViewPtrs& right_l4, int n, int D, int step,
                     int rows, int cols, float* wf, float* wimg, cudaStream_t stream) {
  dim3 grid(cdiv(rows * cols * 8, 256), n);
  launch_pdl(step_warp_kernel, grid, dim3(256), (size_t)0, stream, vol, geo, right_l4, D, step, rows, cols, wf, wimg);
  B200MVS_LAUNCH_OK("step_warp_kernel");
  return 0;
}

int launch_cost(const float* left_feat4, const float* vol, const float* H, int n, int views, int D, int rows,
                int cols, float* cost, uint8_t* mask, cudaStream_t stream) {
  dim3 grid(cdiv(rows * cols * 8, 256), D, n);
  launch_pdl(cost_kernel, grid, dim3(256), (size_t)0, stream, left_feat4, vol, H, views, D, rows, cols, cost, mask);
  B200MVS_LAUNCH_OK("cost_kernel");
  return 0;
}

int launch_cost_norm(const float* cost, long long voxels, float* out, cudaStream_t stream) {
  launch_pdl(cost_norm_kernel, dim3((unsigned)((voxels + 255) / 256)), dim3(256), (size_t)0, stream, cost, voxels, out);
  B200MVS_LAUNCH_OK("cost_norm_kernel");
  return 0;
}

int launch_softargmin(const float* cost, const float* samples, int n, int D, int pixels, float* raw,
                      cudaStream_t stream) {
  dim3 grid(cdiv(pixels, 16), n);
  launch_pdl(softargmin_kernel, grid, dim3(128), (size_t)0, stream, cost, samples, D, pixels, raw);
  B200MVS_LAUNCH_OK("softargmin_kernel");
  return 0;
}

int launch_mask_vote(const float* H, int batch, int views, int D, int rows, int cols, uint8_t* mask4,
                     cudaStream_t stream) {
  launch_pdl(mask_vote_kernel, dim3(cdiv(rows * cols, 128), batch, D), dim3(128), (size_t)0, stream, H, views, D, rows, cols,
             mask4);
  B200MVS_LAUNCH_OK("mask_vote_kernel");
  return 0;
}

int launch_view_reduce(const float* raw_views, const float* refined_views, const uint8_t* mask_views,
                       const float* baseline, int batch, int views, int D, int pixels, bool refined_is_alias,
                       float* raw4, float* idepth4, uint8_t* mask4, cudaStream_t stream) {
  dim3 grid(cdiv(pixels, 128), batch, mask4 != nullptr ? 1 + D : 1);   // (mask4 == nullptr: the vote ran elsewhere)
  launch_pdl(view_reduce_kernel, grid, dim3(128), (size_t)0, stream, raw_views, refined_views, mask_views, baseline, views, D, pixels,
                                               refined_is_alias ? 1 : 0, raw4, idepth4, mask4);
  B200MVS_LAUNCH_OK("view_reduce_kernel");
  return 0;
}

int launch_upsample_f32(const float* in, int n_planes, int h, int w, int H, int W, float* out, cudaStream_t stream) {
  dim3 grid(cdiv(H * W, 256), n_planes);
  launch_pdl(upsample_f32_kernel, grid, dim3(256), (size_t)0, stream, in, h, w, H, W, out);
  B200MVS_LAUNCH_OK("upsample_f32_kernel");
  return 0;
}

int launch_upsample_mask(const uint8_t* in, long long n_planes, int h, int w, int H, int W, uint8_t* out,
                         cudaStream_t stream, bool packed) {
  // gridDim.y is limited to 65535 planes per launch.
  if (packed) {
    const int wb = (W + 7) / 8;
    for (long long p0 = 0; p0 < n_planes; p0 += 65535) {
      const int np = (int)((n_planes - p0) < 65535 ? (n_planes - p0) : 65535);
      dim3 grid(cdiv(H * wb, 256), np);
      if (h == H && w == W)
        launch_pdl(pack_mask_kernel, grid, dim3(256), (size_t)0, stream, in + (size_t)p0 * h * w, H, W,
                   out + (size_t)p0 * H * wb);
      else
        launch_pdl(upsample_mask_packed_kernel, grid, dim3(256), (size_t)0, stream, in + (size_t)p0 * h * w, h, w, H, W,
                   out + (size_t)p0 * H * wb);
      B200MVS_LAUNCH_OK("upsample_mask_packed_kernel");
    }
    return 0;
  }
  const bool exact2x = H == 2 * h && W == 2 * w && w % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(in) & 7) == 0;
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0) && (((long long)H * W) % 4 == 0);
  const bool vec16 = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (long long p0 = 0; p0 < n_planes; p0 += 65535) {
    const int np = (int)((n_planes - p0) < 65535 ? (n_planes - p0) : 65535);
    const uint8_t* src = in + (size_t)p0 * h * w;
    uint8_t* dst = out + (size_t)p0 * H * W;
    if (exact2x) {
      dim3 grid(cdiv(h * (w / 8), 256), np);
      launch_pdl(upsample_mask2x_kernel, grid, dim3(256), (size_t)0, stream, src, h, w, dst);
    } else if (vec16) {
      dim3 grid(cdiv(H * (W / 16), 256), np);
      launch_pdl(upsample_mask_kernel<16>, grid, dim3(256), (size_t)0, stream, src, h, w, H, W, dst);
    } else if (vec) {
      dim3 grid(cdiv(H * (W / 4), 256), np);
      launch_pdl(upsample_mask_kernel<4>, grid, dim3(256), (size_t)0, stream, src, h, w, H, W, dst);
    } else {
      dim3 grid(cdiv(H * W, 256), np);
      launch_pdl(upsample_mask_kernel<1>, grid, dim3(256), (size_t)0, stream, src, h, w, H, W, dst);
    }
    B200MVS_LAUNCH_OK("upsample_mask_kernel");
  }
  return 0;
}

}  // namespace b200mvs
