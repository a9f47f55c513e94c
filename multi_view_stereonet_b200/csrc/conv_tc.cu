// 3x3 (dilated) 32->32 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Implicit GEMM without an im2col copy.  One CTA owns a TH x TW tile of output pixels of one image:
//   * the halo-extended input tile ((TH+2) rows of PW = 64 positions of one vertical polyphase component, see
//     tc_npos; TW = PW - 2d of the positions of a row are valid outputs) is staged ONCE in shared memory, previous layer's GroupNorm + LeakyReLU (+ residual)
//     applied on the way, converted to fp16 and laid out as four planes of [position][8 channels]
//     (16 bytes per position).  With the UMMA "no swizzle, K-major" canonical layout (core matrix =
//     8 rows x 16 bytes, contiguous) a plane IS a valid A operand whose row m is position m, and
//     the A operand of filter tap (ky, kx) is the same plane started (ky*PW + kx*d) positions
//     later -- the nine taps are nine descriptors into one buffer;
//   * per M-tile of 128 consecutive positions (two output rows) 9 taps x 2 k-steps of
//     tcgen05.mma.kind::f16 (M=128, N=32, K=16) accumulate in fp32 into 32 TMEM columns;
//   * the epilogue reads TMEM with tcgen05.ld (thread = output pixel, 32 channels), adds the bias,
//     reduces the GroupNorm statistics of the raw output and stores channels-last fp32.
// Positions whose x falls in the 2d padding columns compute garbage that is never stored.
//
// Operand precision: fp16 (10-bit mantissa, as TF32) with round-to-nearest on conversion, fp32
// accumulation.  SURVEY.md 7.3: the refiner stack tolerates this inside the 1e-3 parity bar; the
// 1/16-scale stages do not and stay on the fp32 path (conv.cu).
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv.cuh"
#include "conv_tc.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int PW = 64;        // positions per tile row (valid outputs: PW - 2*dil)
constexpr int NT = 256;       // threads per CTA

// Dilated layers run as `dil` vertical polyphase components (as in conv_ws.cu): a tile is TH rows of the sub-image made
// of image rows py, py + dil, py + 2 dil, ..., on which the vertical taps are +-1 tile row -- a halo of 2 rows instead
// of 2 dil (at dilation 8 a two-row tile staged 18 rows).  Horizontally the dilation stays in the tap offsets.
__host__ __device__ inline int tc_npos(int TH, int dil) {
  int n = (TH + 2) * PW + 2 * dil;
  return ((n + 5) & ~7) + 2;   // 2 (mod 8): the four octet planes a lane quad writes start in distinct banks
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& h, float* v) {
  const uint32_t w[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
    v[2 * k] = f.x;
    v[2 * k + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
}

// Stores 8 channels of one position: fp16 (single) or hi/lo split into two plane sets.
template <bool SPLIT>
__device__ __forceinline__ void store8(uint8_t* plane_hi, uint32_t lo_offset, int L, const float* v) {
  if (SPLIT) {
    uint4 hi, lo;
    tc::split8(v, &hi, &lo);
    *reinterpret_cast<uint4*>(plane_hi + (size_t)L * 16) = hi;
    *reinterpret_cast<uint4*>(plane_hi + lo_offset + (size_t)L * 16) = lo;
  } else {
    uint4 h;
    h.x = pack_half2(v[0], v[1]);
    h.y = pack_half2(v[2], v[3]);
    h.z = pack_half2(v[4], v[5]);
    h.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(plane_hi + (size_t)L * 16) = h;
  }
}

// Optional timeline of CTA (0,0) in globaltimer ns (B200MVS_TC_PROFILE=1): kernel entry, prologue done (before
// griddepcontrol.wait), dependencies met, coefficients ready, operand staged, MMAs complete, stores issued, exit.
__device__ long long g_tc_prof[8];
__device__ int g_tc_prof_on;
__device__ __forceinline__ long long tc_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(k)                                                                             \
  do {                                                                                          \
    if (g_tc_prof_on && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_tc_prof[k] = tc_gtime(); \
  } while (0)

// Plane sets in shared memory: [feat 0..3 (if a 32-channel source)][extra][zero (if planar extras)], and the same
// again for the lo halves when SPLIT.  K-steps: two over the feature planes, one over (extra, zero).
template <int TH, bool SPLIT, int MINB>
__global__ void __launch_bounds__(NT, MINB) conv3x3_tc_kernel(const ConvParams p, const uint8_t* __restrict__ w16) {
  constexpr int MT = TH * PW / 128;             // M-tiles (two output rows each)
  constexpr int ACC_COLS = SPLIT ? 64 : 32;     // TMEM columns per M-tile
  constexpr int TMEM_COLS = (MT * ACC_COLS) < 32 ? 32 : MT * ACC_COLS;
  constexpr uint32_t WBLOCK = SPLIT ? 2048u : 1024u;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_a[kC], s_b[kC], s_bias[kC];
  __shared__ __align__(8) uint64_t s_bar, s_wbar;
  __shared__ uint32_t s_tmem;

  TC_STAMP(0);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int img = blockIdx.y;
  const int d = p.dil;
  const int TW = PW - 2 * d;
  const int tiles_x = cdiv(p.Wo, TW);
  const int sub_tiles = cdiv(cdiv(p.Ho, d), TH);   // row tiles per vertical phase
  const int tx0 = (blockIdx.x % tiles_x) * TW;
  const int py = (blockIdx.x / tiles_x) / sub_tiles;            // vertical phase: image rows py, py + d, ...
  const int sy0 = ((blockIdx.x / tiles_x) % sub_tiles) * TH;    // first sub-image row of the tile
  const int npos = tc_npos(TH, d);
  const uint32_t plane_bytes = (uint32_t)npos * 16u;
  const int mode = p.feat.mode;
  const bool has_feat = mode != FEAT_NONE;
  const bool has_x = p.extra.n > 0;
  const int ks_feat = has_feat ? 2 : 0;
  const int KS = ks_feat + (has_x ? 1 : 0);
  const int fplanes = has_feat ? 4 : 0;
  const int set_planes = fplanes + (has_x ? 2 : 0);
  const uint32_t lo_offset = (uint32_t)set_planes * plane_bytes;
  const uint32_t w_bytes = 9u * (uint32_t)KS * WBLOCK;
  uint8_t* s_w = smem;
  uint8_t* s_in = smem + w_bytes;

  // ---- one-time setup (nothing here depends on earlier kernels: runs under the previous kernel's tail) ----
  if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)TMEM_COLS);
  if (tid == 32) {
    tc::mbar_init(&s_bar, 1);
    tc::mbar_init(&s_wbar, 1);
    tc::mbar_init_fence();
    // weights (already in the canonical fp16 layout): constant data, fetched by the copy engine; the MMA lane waits
    tc::bulk_load_weights(s_w, w16, w_bytes, &s_wbar);
  }
  if (tid < kC) s_bias[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
  __syncthreads();
  pdl_launch_dependents();   // after the TMEM allocation (see common.cuh)
  TC_STAMP(1);
  pdl_wait();
  TC_STAMP(2);
  const size_t vol = (size_t)p.Hi * p.Wi;
  const int rows_in = TH + 2;
  // ---- stage the transformed 32-channel source ----
  // The raw loads of a batch do not depend on the GroupNorm coefficients: the first batch is issued before the
  // coefficients (two statistics loads + float64 math by 32 threads, then a block barrier) are awaited, so the two
  // memory round trips overlap instead of following each other.
  const bool hio = p.feat.half_io != 0;
  const size_t esz = hio ? 2 : 4;   // bytes per stored activation element
  const uint8_t* fbase = has_feat ? reinterpret_cast<const uint8_t*>(p.feat.ptr) + (size_t)(img / p.feat.img_div) * vol * kC * esz
                                  : nullptr;
  const uint8_t* rbase = p.feat.resid != nullptr
                             ? reinterpret_cast<const uint8_t*>(p.feat.resid) + (size_t)img * vol * kC * esz : nullptr;
  uint8_t* xbase = p.feat.x_out != nullptr ? reinterpret_cast<uint8_t*>(p.feat.x_out) + (size_t)img * vol * kC * esz
                                           : nullptr;
  constexpr int BATCH = MINB >= 3 ? 2 : 4;   // tasks whose global loads are all issued before any is consumed
  float4 ya[BATCH], yb[BATCH], ra[BATCH], rb[BATCH];   // fp32: two float4; fp16: ya / ra hold 8 halves
  size_t off[BATCH];
  bool inb[BATCH];
  auto issue_loads = [&](int i0) {
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int i = i0 + k * NT;
      const int c8 = i & 3;
      const int L = i >> 2;
      const int iy = L / PW, ix = L % PW;
      const int sy = sy0 - 1 + iy;
      const int gy = py + d * sy, gx = tx0 - d + ix;
      inb[k] = i < npos * 4 && iy < rows_in && sy >= 0 && gy < p.Hi && gx >= 0 && gx < p.Wi;
      off[k] = inb[k] ? (((size_t)gy * p.Wi + gx) * kC + 8 * c8) * esz : 0;
      if (inb[k]) {
        ya[k] = __ldg(reinterpret_cast<const float4*>(fbase + off[k]));
        if (!hio) yb[k] = __ldg(reinterpret_cast<const float4*>(fbase + off[k] + 16));
        if (mode == FEAT_GN_RES) {
          ra[k] = __ldg(reinterpret_cast<const float4*>(rbase + off[k]));
          if (!hio) rb[k] = __ldg(reinterpret_cast<const float4*>(rbase + off[k] + 16));
        }
      }
    }
  };
  if (has_feat) issue_loads(tid);
  if (tid < kC && mode >= FEAT_GN) {
    const int grp = tid >> 3;
    const double sum = p.feat.stats[(img * kGroups + grp) * 2 + 0];
    const double sq = p.feat.stats[(img * kGroups + grp) * 2 + 1];
    const double mean = sum * p.feat.inv_count;
    double var = sq * p.feat.inv_count - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = gn_rstd(var);
    s_a[tid] = (float)((double)p.feat.gamma[tid] * rstd);
    s_b[tid] = (float)((double)p.feat.beta[tid] - mean * (double)p.feat.gamma[tid] * rstd);
  }
  __syncthreads();
  TC_STAMP(3);
  if (has_feat) {
    for (int i0 = tid; i0 < npos * 4; i0 += NT * BATCH) {
      if (i0 != tid) issue_loads(i0);
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * NT;
        if (i >= npos * 4) continue;
        const int c8 = i & 3;
        const int L = i >> 2;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (inb[k]) {
          if (hio) {
            unpack8(*reinterpret_cast<const uint4*>(&ya[k]), v);
          } else {
            v[0] = ya[k].x; v[1] = ya[k].y; v[2] = ya[k].z; v[3] = ya[k].w;
            v[4] = yb[k].x; v[5] = yb[k].y; v[6] = yb[k].z; v[7] = yb[k].w;
          }
          if (mode >= FEAT_GN) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(v[e], s_a[8 * c8 + e], s_b[8 * c8 + e]));
            if (mode == FEAT_GN_RES) {
              float r[8];
              if (hio) {
                unpack8(*reinterpret_cast<const uint4*>(&ra[k]), r);
              } else {
                r[0] = ra[k].x; r[1] = ra[k].y; r[2] = ra[k].z; r[3] = ra[k].w;
                r[4] = rb[k].x; r[5] = rb[k].y; r[6] = rb[k].z; r[7] = rb[k].w;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += r[e];
            }
            const int iy = L / PW, ix = L % PW;
            if (xbase != nullptr && iy >= 1 && iy < 1 + TH && ix >= d && ix < d + TW) {
              if (hio) {
                *reinterpret_cast<uint4*>(xbase + off[k]) = pack8(v);
              } else {
                *reinterpret_cast<float4*>(xbase + off[k]) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(xbase + off[k] + 16) = make_float4(v[4], v[5], v[6], v[7]);
              }
            }
          }
        }
        store8<SPLIT>(s_in + (size_t)c8 * plane_bytes, lo_offset, L, v);
      }
    }
  }
  // ---- stage the planar extra channels (image planes, scaled idepth) and the zero plane ----
  if (has_x) {
    uint8_t* xplane = s_in + (size_t)fplanes * plane_bytes;
    for (int L = tid; L < npos; L += NT) {
      const int iy = L / PW, ix = L % PW;
      const int sy = sy0 - 1 + iy;
      const int gy = py + d * sy, gx = tx0 - d + ix;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (iy < rows_in && sy >= 0 && gy < p.Hi && gx >= 0 && gx < p.Wi) {
        const size_t pix = (size_t)gy * p.Wi + gx;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (e < p.extra.n) {
            float x = __ldg(p.extra.ptr[e] + (size_t)(img / p.extra.img_div[e]) * p.extra.img_stride[e] + pix);
            if (p.extra.scale[e] != nullptr)
              x = __fmul_rn(x, __ldg(p.extra.scale[e] + (size_t)(img / p.extra.scale_div[e]) * p.extra.scale_stride[e]));
            v[e] = x;
          }
        }
      }
      store8<SPLIT>(xplane, lo_offset, L, v);
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(xplane + plane_bytes + (size_t)L * 16) = z;
      if (SPLIT) *reinterpret_cast<uint4*>(xplane + plane_bytes + lo_offset + (size_t)L * 16) = z;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = s_tmem;
  TC_STAMP(4);

  // ---- one elected lane of warp 0 issues every MMA of the tile, then commits to the mbarrier ----
  if (warp == 0) {
   if (tc::elect_one()) {
    const uint32_t plane_u16 = plane_bytes >> 4;
    const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_in), plane_bytes, 128u);
    const uint64_t da_lo0 = da0 + (uint64_t)(set_planes * plane_u16);
    const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), SPLIT ? 1024u : 512u, 128u);
    tc::mbar_wait(&s_wbar, 0u);   // the bulk-copied weights have landed
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t dcol = tmem_base + (uint32_t)(mt * ACC_COLS);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t pos = (uint32_t)(mt * 128 + (tap / 3) * PW + (tap % 3) * d);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          if (ks < KS) {
            const uint32_t pl = (ks < ks_feat) ? 2u * ks : (uint32_t)fplanes;
            const uint64_t a_off = (uint64_t)(pl * plane_u16 + pos);
            const uint64_t b = db0 + (uint64_t)((tap * KS + ks) * (WBLOCK >> 4));
            if (SPLIT) {
              tc::mma_f16(dcol, da0 + a_off, b, tc::idesc_f16(64), (tap | ks) != 0 ? 1u : 0u);
              tc::mma_f16(dcol, da_lo0 + a_off, b, tc::idesc_f16(32), 1u);
            } else {
              tc::mma_f16(dcol, da0 + a_off, b, tc::idesc_f16(32), (tap | ks) != 0 ? 1u : 0u);
            }
          }
        }
      }
    }
    tc::mma_commit(&s_bar);
    tc::mbar_wait(&s_bar, 0u);   // only the issuing thread polls; the rest park at the hardware barrier
    tc::fence_before_sync();
   }
   __syncwarp();
  }
  __syncthreads();
  tc::fence_after_sync();

  TC_STAMP(5);
  // ---- epilogue: TMEM -> registers -> bias, statistics, channels-last store ----
  const int wq = warp & 3;  // TMEM lane quarter this warp may read
  float gsum[kGroups], gsq[kGroups];
#pragma unroll
  for (int g = 0; g < kGroups; ++g) gsum[g] = gsq[g] = 0.f;
  const size_t ovol = (size_t)p.Ho * p.Wo;
  const size_t ostride = p.out_img_stride != 0 ? (size_t)p.out_img_stride : ovol * kC;
  for (int mt = warp >> 2; mt < MT; mt += NT / 128) {
    const int j = mt * 128 + wq * 32 + lane;
    const int oy = py + d * (sy0 + j / PW), ox_t = j % PW;
    const int ox = tx0 + ox_t;
    const bool valid = ox_t < TW && ox < p.Wo && oy < p.Ho;
    const size_t opix = valid ? (size_t)oy * p.Wo + ox : 0;
    float* o = p.out + (size_t)img * ostride + opix * kC;
    const float* add = p.add_src != nullptr ? p.add_src + ((size_t)img * ovol + opix) * kC : nullptr;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[16];
      const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * ACC_COLS + half * 16);
      tc::tmem_ld16(ta, v);
      if (SPLIT) {
        float c[16];
        tc::tmem_ld16(ta + 32u, c);
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += c[k];
      }
      if (valid) {
        float r16[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = half * 16 + 4 * q;
          float4 r;
          r.x = v[4 * q + 0] + s_bias[ch + 0];
          r.y = v[4 * q + 1] + s_bias[ch + 1];
          r.z = v[4 * q + 2] + s_bias[ch + 2];
          r.w = v[4 * q + 3] + s_bias[ch + 3];
          if (add != nullptr) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(add + ch));
            r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
          }
          r16[4 * q + 0] = r.x; r16[4 * q + 1] = r.y; r16[4 * q + 2] = r.z; r16[4 * q + 3] = r.w;
          gsum[ch >> 3] += (r.x + r.y) + (r.z + r.w);
          gsq[ch >> 3] += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
        }
        if (p.out_half) {
          __half* oh = reinterpret_cast<__half*>(p.out) + (size_t)img * ostride + opix * kC + half * 16;
          *reinterpret_cast<uint4*>(oh) = pack8(r16);
          *reinterpret_cast<uint4*>(oh + 8) = pack8(r16 + 8);
        } else {
          st8(o + half * 16, r16);       // 256-bit stores
          st8(o + half * 16 + 8, r16 + 8);
        }
      }
    }
  }
  TC_STAMP(6);
  if (p.out_stats != nullptr) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gsum[g] += __shfl_xor_sync(0xffffffffu, gsum[g], o);
        gsq[g] += __shfl_xor_sync(0xffffffffu, gsq[g], o);
      }
    }
    // every warp adds its totals straight to the layer's statistics (8 reductions in flight per warp, nothing waits
    // for them but the end of the kernel): lane 2g adds the sum of group g, lane 2g + 1 its sum of squares
    if (lane < 2 * kGroups) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < kGroups; ++g)
        if ((lane >> 1) == g) v = (lane & 1) ? gsq[g] : gsum[g];
      atomicAdd(p.out_stats + (size_t)img * 2 * kGroups + lane, (double)v);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  TC_STAMP(7);
  if (warp == 0) tc::tmem_dealloc(tmem_base, (uint32_t)TMEM_COLS);
}

size_t tc_smem_bytes(int TH, bool split, const ConvParams& p) {
  const int ks = (p.feat.mode != FEAT_NONE ? 2 : 0) + (p.extra.n > 0 ? 1 : 0);
  const int planes = ((p.feat.mode != FEAT_NONE ? 4 : 0) + (p.extra.n > 0 ? 2 : 0)) * (split ? 2 : 1);
  return (size_t)9 * ks * (split ? 2048 : 1024) + (size_t)planes * tc_npos(TH, p.dil) * 16;
}

template <int TH, bool SPLIT, int MINB = 1>
int launch_th(const ConvParams& p, const uint8_t* w16, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(TH, SPLIT, p);
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&conv3x3_tc_kernel<TH, SPLIT, MINB>), 220 * 1024)) return rc;
  const int TW = PW - 2 * p.dil;
  dim3 grid(cdiv(p.Wo, TW) * p.dil * cdiv(cdiv(p.Ho, p.dil), TH), p.n_img);
  if (p.tag != TAG_NONE) probe_before(p.tag, stream);
  static const bool prof = getenv("B200MVS_TC_PROFILE") != nullptr;
  if (prof) {
    const int one = 1;
    cudaMemcpyToSymbol(g_tc_prof_on, &one, sizeof(int));
  }
  launch_pdl(conv3x3_tc_kernel<TH, SPLIT, MINB>, grid, dim3(NT), smem, stream, p, w16);
  if (prof) {
    long long h[8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
    fprintf(stderr, "tc TH=%d split=%d grid=%dx%d dil=%d: prologue %lld | dep wait %lld | coeffs %lld | stage %lld | mma %lld | epilogue %lld | stats+exit %lld (ns)\n",
            TH, (int)SPLIT, grid.x, grid.y, p.dil, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6]);
  }
  if (p.tag != TAG_NONE) probe_after(p.tag, stream);
  B200MVS_LAUNCH_OK("conv3x3_tc_kernel");
  return 0;
}

}  // namespace

void pack_conv3x3_tc_weights(const float* w_oihw, int cin, bool has_feat, int feat_off,
                             const std::vector<int>& extra_idx, bool split, std::vector<uint8_t>* out) {
  const int ks_feat = has_feat ? 2 : 0;
  const int KS = ks_feat + (extra_idx.empty() ? 0 : 1);
  const size_t block_halves = split ? 1024 : 512;
  out->assign((size_t)9 * KS * block_halves * 2, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  auto put = [&](int block, int k, int n, float w) {
    if (split) {
      tc::put_split_weight(h, block, k, n, w);
    } else {
      h[(size_t)block * 512 + (size_t)(k / 8) * 256 + (size_t)n * 8 + (size_t)(k % 8)] = __float2half_rn(w);
    }
  };
  for (int tap = 0; tap < 9; ++tap)
    for (int n = 0; n < 32; ++n) {
      if (has_feat)
        for (int c = 0; c < 32; ++c) put(tap * KS + c / 16, c % 16, n, w_oihw[((size_t)n * cin + feat_off + c) * 9 + tap]);
      for (size_t e = 0; e < extra_idx.size(); ++e)
        put(tap * KS + ks_feat, (int)e, n, w_oihw[((size_t)n * cin + extra_idx[e]) * 9 + tap]);
    }
}

bool conv3x3_tc_supported(const ConvParams& p) {
  return (p.feat.mode != FEAT_NONE || p.extra.n > 0) && p.extra.n <= 4 && p.Di == 1 && p.Do == 1 && p.Hi == p.Ho &&
         p.Wi == p.Wo && p.dil >= 1 && p.dil <= 8;
}

int launch_conv3x3_tc(const ConvParams& p, const uint8_t* w16, bool split, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  if (!conv3x3_tc_supported(p)) {
    set_error("launch_conv3x3_tc: unsupported configuration");
    return -1;
  }
  // Tile height: the tallest tile that still gives every SM work and fits two CTAs per SM; small images get
  // short tiles so that a layer is not a handful of long-running CTAs.
  const int TW = PW - 2 * p.dil;
  const int ths[3] = {16, 8, 4};
  int pick = 4;
  for (int k = 0; k < 3; ++k) {
    const int th = ths[k];
    if (split && th == 16) continue;  // 8 M-tiles x 64 columns would exceed TMEM
    const long long tiles = (long long)cdiv(p.Wo, TW) * p.dil * cdiv(cdiv(p.Ho, p.dil), th) * p.n_img;
    const size_t smem = tc_smem_bytes(th, split, p);
    if (smem > 210 * 1024) continue;
    if (tiles >= 296 && (smem <= 110 * 1024 || p.dil >= 8)) { pick = th; break; }  // big halo: amortise it
    if (tiles >= 148 && th != 16) { pick = th; break; }
    if (th == 4) pick = 4;
  }
  if (tc_smem_bytes(pick, split, p) > 210 * 1024) {
    set_error("launch_conv3x3_tc: tile does not fit in shared memory");
    return -1;
  }
  // Small images (a few dozen tiles at most): the layer is one dependent latency chain, so halve the chain of every
  // CTA -- two-row tiles = one M-tile each -- rather than amortise the halo.
  static const int small_th = getenv("B200MVS_TC_SMALL_TH") ? atoi(getenv("B200MVS_TC_SMALL_TH")) : 2;
  const bool tiny = pick == 4 && (long long)cdiv(p.Wo, TW) * p.dil * cdiv(cdiv(p.Ho, p.dil), 4) * p.n_img < 74 && small_th == 2;
  if (split) {
    if (pick == 8) return launch_th<8, true>(p, w16, stream);
    if (tiny) return launch_th<2, true>(p, w16, stream);
    return launch_th<4, true>(p, w16, stream);
  }
  if (tiny) return launch_th<2, false>(p, w16, stream);
  // Large images, small halo: 8-row tiles at three CTAs per SM overlap one CTA's loads with another's MMAs and
  // stores better than 16-row tiles at two per SM (measured ~50 vs ~72 us per level-0 layer).
  static const int variant = getenv("B200MVS_TC_VARIANT") ? atoi(getenv("B200MVS_TC_VARIANT")) : 1;
  if (variant == 1 && pick == 16 && p.dil <= 2) return launch_th<8, false, 3>(p, w16, stream);
  if (pick == 16) return launch_th<16, false>(p, w16, stream);
  if (pick == 8) return launch_th<8, false>(p, w16, stream);
  return launch_th<4, false>(p, w16, stream);
}

}  // namespace b200mvs
