// fp32 direct convolution for sm_100a: 3x3 (dilated), 5x5 stride-2 and 3x3x3, 32 or 1 output
// channels, channels-last activations.  See conv.cuh for the fusion contract.
//
// Work decomposition: one CTA = one TD x TH x TW tile of output pixels of one image, all output
// channels.  Input channels are consumed in chunks of 8: the (halo-extended) input tile of the chunk
// is staged in shared memory channel-planar (so lanes that walk along x hit distinct banks) with the
// previous layer's GroupNorm/LeakyReLU/residual applied on the fly, next to the chunk's weights
// [tap][8][COUT] which every lane reads as a broadcast.  Each thread owns PXT pixels x COUT/CSPLIT
// channels in registers.  The epilogue transposes through shared memory so that global stores are
// full 128-byte channel vectors, and reduces the GroupNorm statistics of what it stores.
#include <cuda_fp16.h>

#include "conv.cuh"

namespace b200mvs {
namespace {

constexpr int CK = 8;  // input channels per chunk

template <int KD, int KH, int KW, int S, int COUT, int TD, int TH, int TW, int PXT, int CSPLIT>
struct Cfg {
  static constexpr int TILE = TD * TH * TW;
  static constexpr int PIX_THREADS = TILE / PXT;
  static constexpr int NT = PIX_THREADS * CSPLIT;
  static constexpr int TWT = TW / PXT;
  static constexpr int TAPS = KD * KH * KW;
  static constexpr int CPT = COUT / CSPLIT;  // output channels per thread
  static_assert(TW % PXT == 0 && COUT % CSPLIT == 0, "bad tiling");
  static_assert(PIX_THREADS % 32 == 0, "channel split must be warp-uniform");
  static_assert(NT % 8 == 0, "epilogue mapping needs NT % 8 == 0");
};

struct TileGeom {
  int IZ, IY, IX, IXP, plane;
};

template <class C, int KD, int KH, int KW, int S, int TD, int TH, int TW>
__host__ __device__ inline TileGeom tile_geom(int dil) {
  TileGeom g;
  g.IZ = (TD - 1) + (KD - 1) + 1;
  g.IY = (TH - 1) * S + (KH - 1) * dil + 1;
  g.IX = (TW - 1) * S + (KW - 1) * dil + 1;
  int ixp = g.IX;
  if (S == 1) {
    // A warp covers 32/TWT tile rows of TWT consecutive floats: rows must start TWT banks apart.
    if (C::TWT < 32) {
      while (ixp % (2 * C::TWT) != C::TWT) ++ixp;
    }
  } else {
    if (ixp % 2 == 0) ++ixp;  // stride-2 lanes use every other bank; odd pitch interleaves rows
  }
  g.IXP = ixp;
  int plane = g.IZ * g.IY * g.IXP;
  while (plane % 8 != 4) ++plane;  // loader lanes (pixel, quad) write planes 4 apart
  g.plane = plane;
  return g;
}

template <class C>
__host__ __device__ inline size_t smem_floats(const TileGeom& g, int cout) {
  size_t a = (size_t)CK * g.plane + (size_t)C::TAPS * CK * cout;
  size_t b = (cout == 32) ? (size_t)C::TILE * 33 : 0;
  return a > b ? a : b;
}

template <int KD, int KH, int KW, int S, int COUT, int TD, int TH, int TW, int PXT, int CSPLIT>
__global__ void __launch_bounds__(Cfg<KD, KH, KW, S, COUT, TD, TH, TW, PXT, CSPLIT>::NT)
conv_kernel(const ConvParams p) {
  using C = Cfg<KD, KH, KW, S, COUT, TD, TH, TW, PXT, CSPLIT>;
  constexpr int NT = C::NT;
  constexpr int TAPS = C::TAPS;
  constexpr int CPT = C::CPT;
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_a[kC], s_b[kC];
  __shared__ double s_stats[2 * kGroups];

  const int tid = threadIdx.x;
  const int img = blockIdx.y;
  const int dil = p.dil;
  const TileGeom g = tile_geom<C, KD, KH, KW, S, TD, TH, TW>(dil);
  float* s_in = smem;                  // [CK][plane]
  float* s_w = smem + CK * g.plane;    // [TAPS][CK][COUT]

  // Tile origin.
  const int tiles_x = cdiv(p.Wo, TW), tiles_y = cdiv(p.Ho, TH);
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * TW;
  t /= tiles_x;
  const int ty0 = (t % tiles_y) * TH;
  const int tz0 = (t / tiles_y) * TD;
  const int ix0 = tx0 * S - (KW / 2) * dil;
  const int iy0 = ty0 * S - (KH / 2) * dil;
  const int iz0 = tz0 - (KD / 2);

  pdl_launch_dependents();
  pdl_wait();
  // Previous layer's GroupNorm folded into a per-channel scale/shift.
  if (p.feat.mode >= FEAT_GN) {
    if (tid < kC) {
      const int grp = tid >> 3;
      const double sum = p.feat.stats[(img * kGroups + grp) * 2 + 0];
      const double sq = p.feat.stats[(img * kGroups + grp) * 2 + 1];
      const double mean = sum * p.feat.inv_count;
      double var = sq * p.feat.inv_count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const double rstd = gn_rstd(var);
      const float a = (float)((double)p.feat.gamma[tid] * rstd);
      s_a[tid] = a;
      s_b[tid] = (float)((double)p.feat.beta[tid] - mean * (double)p.feat.gamma[tid] * rstd);
    }
  }
  if (tid < 2 * kGroups) s_stats[tid] = 0.0;
  __syncthreads();

  // This thread's output pixels.
  const int ptid = tid % C::PIX_THREADS;
  const int cgrp = tid / C::PIX_THREADS;  // warp-uniform
  const int lx = ptid % C::TWT;
  const int ly = (ptid / C::TWT) % TH;
  const int lz = ptid / (C::TWT * TH);
  int base[PXT];
#pragma unroll
  for (int j = 0; j < PXT; ++j) base[j] = (lz * g.IY + ly * S) * g.IXP + (lx + j * C::TWT) * S;

  float acc[PXT][CPT];
#pragma unroll
  for (int j = 0; j < PXT; ++j)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[j][c] = 0.0f;

  const int n_feat_chunks = (p.feat.mode != FEAT_NONE) ? kC / CK : 0;
  const int n_chunks = n_feat_chunks + (p.extra.n > 0 ? 1 : 0);
  const int in_px = g.IZ * g.IY * g.IX;
  const int fimg = img / p.feat.img_div;

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    // ---- stage the input tile of this chunk ----
    if (chunk < n_feat_chunks) {
      const int c0 = chunk * CK;
      for (int i = tid; i < in_px * 2; i += NT) {
        const int q = i & 1;
        int pp = i >> 1;
        const int x = pp % g.IX;
        pp /= g.IX;
        const int y = pp % g.IY;
        const int z = pp / g.IY;
        const int gx = ix0 + x, gy = iy0 + y, gz = iz0 + z;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < p.Wi && gy >= 0 && gy < p.Hi && gz >= 0 && gz < p.Di) {
          const size_t pix = ((size_t)gz * p.Hi + gy) * p.Wi + gx;
          const size_t vol = (size_t)p.Di * p.Hi * p.Wi;
          const int c = c0 + 4 * q;
          if (p.feat.half_io) {
            const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.feat.ptr) +
                                                                 ((size_t)fimg * vol + pix) * kC + c));
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            v = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            v = __ldg(reinterpret_cast<const float4*>(p.feat.ptr + ((size_t)fimg * vol + pix) * kC + c));
          }
          if (p.feat.mode >= FEAT_GN) {
            v.x = lrelu(fmaf(v.x, s_a[c + 0], s_b[c + 0]));
            v.y = lrelu(fmaf(v.y, s_a[c + 1], s_b[c + 1]));
            v.z = lrelu(fmaf(v.z, s_a[c + 2], s_b[c + 2]));
            v.w = lrelu(fmaf(v.w, s_a[c + 3], s_b[c + 3]));
            if (p.feat.mode == FEAT_GN_RES) {
              float4 r;
              if (p.feat.half_io) {
                const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.feat.resid) +
                                                                     ((size_t)img * vol + pix) * kC + c));
                const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
                const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
                r = make_float4(lo.x, lo.y, hi.x, hi.y);
              } else {
                r = __ldg(reinterpret_cast<const float4*>(p.feat.resid + ((size_t)img * vol + pix) * kC + c));
              }
              v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            // Only stride-1 same-size layers carry GroupNorm inputs: output pixel == input pixel.
            if (p.feat.x_out != nullptr && gx >= tx0 && gx < tx0 + TW && gy >= ty0 && gy < ty0 + TH &&
                gz >= tz0 && gz < tz0 + TD) {
              *reinterpret_cast<float4*>(p.feat.x_out + ((size_t)img * vol + pix) * kC + c) = v;
            }
          }
        }
        float* dst = s_in + (4 * q) * g.plane + (z * g.IY + y) * g.IXP + x;
        dst[0] = v.x;
        dst[g.plane] = v.y;
        dst[2 * g.plane] = v.z;
        dst[3 * g.plane] = v.w;
      }
    } else {
      for (int i = tid; i < in_px * CK; i += NT) {
        const int e = i / in_px;
        int pp = i - e * in_px;
        const int x = pp % g.IX;
        pp /= g.IX;
        const int y = pp % g.IY;
        const int z = pp / g.IY;
        const int gx = ix0 + x, gy = iy0 + y, gz = iz0 + z;
        float v = 0.f;
        if (e < p.extra.n && gx >= 0 && gx < p.Wi && gy >= 0 && gy < p.Hi && gz >= 0 && gz < p.Di) {
          const size_t pix = ((size_t)gz * p.Hi + gy) * p.Wi + gx;
          v = __ldg(p.extra.ptr[e] + (size_t)(img / p.extra.img_div[e]) * p.extra.img_stride[e] + pix);
          if (p.extra.scale[e] != nullptr)
            v = __fmul_rn(v, __ldg(p.extra.scale[e] + (size_t)(img / p.extra.scale_div[e]) * p.extra.scale_stride[e]));
        }
        s_in[e * g.plane + (z * g.IY + y) * g.IXP + x] = v;
      }
    }
    // ---- stage the chunk's weights ----
    {
      const float* wsrc = p.w + (size_t)chunk * TAPS * CK * COUT;
      if (COUT % 4 == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(wsrc);
        float4* d4 = reinterpret_cast<float4*>(s_w);
        for (int i = tid; i < TAPS * CK * COUT / 4; i += NT) d4[i] = __ldg(w4 + i);
      } else {
        for (int i = tid; i < TAPS * CK * COUT; i += NT) s_w[i] = __ldg(wsrc + i);
      }
    }
    __syncthreads();

    // ---- accumulate ----
#pragma unroll 1
    for (int kz = 0; kz < KD; ++kz) {
#pragma unroll 1
      for (int ky = 0; ky < KH; ++ky) {
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          const int tap = (kz * KH + ky) * KW + kx;
          const int off = (kz * g.IY + ky * dil) * g.IXP + kx * dil;
#pragma unroll
          for (int c = 0; c < CK; ++c) {
            float v[PXT];
#pragma unroll
            for (int j = 0; j < PXT; ++j) v[j] = s_in[c * g.plane + base[j] + off];
            if (COUT == 32) {
              const float4* w4 = reinterpret_cast<const float4*>(s_w + (tap * CK + c) * COUT + cgrp * CPT);
#pragma unroll
              for (int q = 0; q < CPT / 4; ++q) {
                const float4 w = w4[q];
#pragma unroll
                for (int j = 0; j < PXT; ++j) {
                  acc[j][4 * q + 0] = fmaf(v[j], w.x, acc[j][4 * q + 0]);
                  acc[j][4 * q + 1] = fmaf(v[j], w.y, acc[j][4 * q + 1]);
                  acc[j][4 * q + 2] = fmaf(v[j], w.z, acc[j][4 * q + 2]);
                  acc[j][4 * q + 3] = fmaf(v[j], w.w, acc[j][4 * q + 3]);
                }
              }
            } else {
              const float w = s_w[tap * CK + c];
#pragma unroll
              for (int j = 0; j < PXT; ++j) acc[j][0] = fmaf(v[j], w, acc[j][0]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const size_t ovol = (size_t)p.Do * p.Ho * p.Wo;
  const size_t ostride = p.out_img_stride != 0 ? (size_t)p.out_img_stride : ovol * COUT;
  if (COUT == 32) {
    // ---- transpose through shared memory, store full channel vectors, reduce GN statistics ----
    float* s_out = smem;  // [TILE][33]
#pragma unroll
    for (int j = 0; j < PXT; ++j) {
      const int pix = (lz * TH + ly) * TW + lx + j * C::TWT;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int ch = cgrp * CPT + c;
        s_out[pix * 33 + ch] = acc[j][c] + (p.bias != nullptr ? __ldg(p.bias + ch) : 0.f);
      }
    }
    __syncthreads();
    float sum = 0.f, sq = 0.f;
    for (int i = tid; i < C::TILE * 8; i += NT) {
      const int q = i & 7;
      int pix = i >> 3;
      const int x = pix % TW;
      const int y = (pix / TW) % TH;
      const int z = pix / (TW * TH);
      const int ox = tx0 + x, oy = ty0 + y, oz = tz0 + z;
      if (ox < p.Wo && oy < p.Ho && oz < p.Do) {
        const size_t opix = ((size_t)oz * p.Ho + oy) * p.Wo + ox;
        const size_t o = (size_t)img * ostride + opix * kC + 4 * q;
        float4 v;
        v.x = s_out[pix * 33 + 4 * q + 0];
        v.y = s_out[pix * 33 + 4 * q + 1];
        v.z = s_out[pix * 33 + 4 * q + 2];
        v.w = s_out[pix * 33 + 4 * q + 3];
        if (p.add_src != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p.add_src + ((size_t)img * ovol + opix) * kC + 4 * q));
          v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        *reinterpret_cast<float4*>(p.out + o) = v;
        sum += (v.x + v.y) + (v.z + v.w);
        sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    }
    if (p.out_stats != nullptr) {
      // lane bits: [0] low bit of the channel quad, [2:1] group, [4:3] pixel.
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sq += __shfl_xor_sync(0xffffffffu, sq, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      sq += __shfl_xor_sync(0xffffffffu, sq, 16);
      const int lane = tid & 31;
      if (lane < 8 && (lane & 1) == 0) {
        atomicAdd(&s_stats[(lane >> 1) * 2 + 0], (double)sum);
        atomicAdd(&s_stats[(lane >> 1) * 2 + 1], (double)sq);
      }
      __syncthreads();
      if (tid < 2 * kGroups) atomicAdd(p.out_stats + (size_t)img * 2 * kGroups + tid, s_stats[tid]);
    }
  } else {
    const float b = p.bias != nullptr ? __ldg(p.bias) : 0.f;
#pragma unroll
    for (int j = 0; j < PXT; ++j) {
      const int ox = tx0 + lx + j * C::TWT, oy = ty0 + ly, oz = tz0 + lz;
      if (ox < p.Wo && oy < p.Ho && oz < p.Do) {
        const size_t opix = ((size_t)oz * p.Ho + oy) * p.Wo + ox;
        const size_t o = (size_t)img * ostride + opix;
        float v = acc[j][0] + b;
        if (p.epi1_mode == 1) {
          // IDepthmapRefiner tail with the caller's scaling (multi_view_stereonet.py:482, 607-611):
          // relu(idepth * fx + delta) / fx
          const float f = __ldg(p.fx + (size_t)(img / p.fx_div) * p.fx_stride);
          const float scaled = __fmul_rn(__ldg(p.prior + (size_t)img * ovol + opix), f);
          v = __fdiv_rn(fmaxf(__fadd_rn(scaled, v), 0.f), f);
        }
        p.out[o] = v;
      }
    }
  }
}

template <int KD, int KH, int KW, int S, int COUT, int TD, int TH, int TW, int PXT, int CSPLIT>
int launch_cfg(const ConvParams& p, cudaStream_t stream) {
  using C = Cfg<KD, KH, KW, S, COUT, TD, TH, TW, PXT, CSPLIT>;
  const TileGeom g = tile_geom<C, KD, KH, KW, S, TD, TH, TW>(p.dil);
  const size_t smem = smem_floats<C>(g, COUT) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("conv tile does not fit in shared memory");
    return -1;
  }
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&conv_kernel<KD, KH, KW, S, COUT, TD, TH, TW, PXT, CSPLIT>),
                                200 * 1024))
    return rc;
  dim3 grid(cdiv(p.Wo, TW) * cdiv(p.Ho, TH) * cdiv(p.Do, TD), p.n_img);
  if (p.tag != TAG_NONE) probe_before(p.tag, stream);
  launch_pdl(conv_kernel<KD, KH, KW, S, COUT, TD, TH, TW, PXT, CSPLIT>, grid, dim3(C::NT), smem, stream, p);
  if (p.tag != TAG_NONE) probe_after(p.tag, stream);
  B200MVS_LAUNCH_OK("conv_kernel");
  return 0;
}

}  // namespace

int conv_init() { return 0; }

int launch_conv(ConvKind kind, int cout, const ConvParams& p, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  if (cout != 32 && cout != 1) {
    set_error("launch_conv: cout must be 32 or 1");
    return -1;
  }
  // Small images (the 1/16 and 1/8 levels) get small tiles with the output channels split over
  // four thread groups: few pixels, so latency per layer matters more than reuse.
  const bool small = (long long)p.Ho * p.Wo <= 96 * 128;
  switch (kind) {
    case CONV_3x3:
      if (cout == 32)
        return small ? launch_cfg<1, 3, 3, 1, 32, 1, 8, 8, 1, 4>(p, stream)
                     : launch_cfg<1, 3, 3, 1, 32, 1, 8, 32, 2, 1>(p, stream);
      return small ? launch_cfg<1, 3, 3, 1, 1, 1, 8, 8, 1, 1>(p, stream)
                   : launch_cfg<1, 3, 3, 1, 1, 1, 8, 32, 2, 1>(p, stream);
    case CONV_5x5_S2:
      if (cout != 32) break;
      return small ? launch_cfg<1, 5, 5, 2, 32, 1, 8, 8, 1, 4>(p, stream)
                   : launch_cfg<1, 5, 5, 2, 32, 1, 8, 32, 2, 1>(p, stream);
    case CONV_3x3x3:
      if (cout == 32) return launch_cfg<3, 3, 3, 1, 32, 4, 8, 8, 2, 1>(p, stream);
      return launch_cfg<3, 3, 3, 1, 1, 4, 8, 8, 2, 1>(p, stream);
  }
  set_error("launch_conv: unsupported configuration");
  return -1;
}

}  // namespace b200mvs
