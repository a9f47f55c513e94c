// The four 5x5 stride-2 convolutions of FeatureNetwork (multi_view_stereonet.py:90-94, 113-116; no bias, no
// activation between them) on the tcgen05 tensor cores with split-fp16 operands (x = hi + lo, three products,
// fp32 accumulation in TMEM: fp32-class accuracy, which the 1/16-scale stages need -- SURVEY.md 7.3).
//
// 32 -> 32 channels (conv1..conv3).  A stride-2 convolution is four stride-1 convolutions over the parity phases
// of the input: with ky = 2a + py, kx = 2b + px,
//     in(2 oy + ky - 2, 2 ox + kx - 2) = P[py][px](oy + a - 1, ox + b - 1),      P[py][px](r, c) = in(2r + py, 2c + px)
// so each phase image is staged once as [position][8 channels] fp16 planes (conv_tc.cu) and tap (a, b) of a phase is
// the same plane started a * PW + b positions later: 9 + 6 + 6 + 4 = 25 shifted-window descriptors, no im2col copy.
//
// 3 -> 32 channels (conv0, planar full-resolution input).  K = 75 is too thin for the window trick; the CTA stages
// its input tile in shared memory and builds the im2col operand [position][80] explicitly (5 k-steps of 16).
#include <cuda_fp16.h>

#include <type_traits>
#include <vector>

#include "conv5_tc.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int NT = 256;

// ------------------------------------------------------------------------------------------------------------
// 32 -> 32
// ------------------------------------------------------------------------------------------------------------
constexpr int A_PW = 32;                 // positions per phase-tile row; valid output columns: A_PW - 2
constexpr int A_TH = 4;                  // output rows per tile -> one 128-position M-tile
constexpr int A_TW = A_PW - 2;
constexpr int A_NPOS = 201;              // (A_TH + 2) * A_PW + 2 = 194, padded so that the plane stride is an odd
                                         // multiple of 16 B (the 8 planes a warp writes at once hit distinct banks)
constexpr uint32_t A_PLANE = A_NPOS * 16u;
constexpr int A_TAPS = 25;
constexpr uint32_t A_WBYTES = A_TAPS * 2 * 2048u;
constexpr uint32_t A_SMEM = A_WBYTES + 32u * A_PLANE;
constexpr int A_IN_ROWS = 2 * (A_TH + 2), A_IN_COLS = 2 * A_PW;

__global__ void __launch_bounds__(NT, 1) conv5x5s2_c32_tc_kernel(const float* __restrict__ in,
                                                                  const uint8_t* __restrict__ w16, int Hi, int Wi,
                                                                  int Ho, int Wo, float* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar, s_wbar;
  __shared__ uint32_t s_tmem;
  uint8_t* s_w = smem;
  uint8_t* s_pl = smem + A_WBYTES;   // planes [hi | lo][phase (py * 2 + px)][octet]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int img = blockIdx.y;
  const int tiles_x = cdiv(Wo, A_TW);
  const int ox0 = (blockIdx.x % tiles_x) * A_TW, oy0 = (blockIdx.x / tiles_x) * A_TH;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 64u);
  if (tid == 32) {
    tc::mbar_init(&s_bar, 1);
    tc::mbar_init(&s_wbar, 1);
    tc::mbar_init_fence();
    tc::bulk_load_weights(s_w, w16, A_WBYTES, &s_wbar);   // constant data: before griddepcontrol.wait
  }
  __syncthreads();
  pdl_launch_dependents();
  {
    // positions past the last staged row are read by the garbage columns only, but must not hold NaN patterns
    for (int i = tid; i < 32 * (A_NPOS - A_IN_ROWS / 2 * A_PW); i += NT) {
      const int pl = i / (A_NPOS - A_IN_ROWS / 2 * A_PW), r = i % (A_NPOS - A_IN_ROWS / 2 * A_PW);
      *reinterpret_cast<uint4*>(s_pl + (size_t)pl * A_PLANE + (size_t)(A_IN_ROWS / 2 * A_PW + r) * 16) =
          make_uint4(0, 0, 0, 0);
    }
  }
  pdl_wait();

  // ---- stage the input tile: 12 x 64 pixels x 32 channels, de-interleaved into the four parity phases ----
  {
    const float* base = in + (size_t)img * Hi * Wi * kC;
    const int iy0 = 2 * (oy0 - 1), ix0 = 2 * (ox0 - 1);
    constexpr int TASKS = A_IN_ROWS * A_IN_COLS * 4;   // (pixel, channel octet)
    constexpr int BATCH = 4;
    for (int i0 = tid; i0 < TASKS; i0 += NT * BATCH) {
      float4 a[BATCH], b[BATCH];
      bool inb[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * NT;
        const int oct = i & 3, px = i >> 2;
        const int lx = px % A_IN_COLS, ly = px / A_IN_COLS;
        const int gy = iy0 + ly, gx = ix0 + lx;
        inb[k] = i < TASKS && gy >= 0 && gy < Hi && gx >= 0 && gx < Wi;
        if (inb[k]) {
          const float4* p = reinterpret_cast<const float4*>(base + ((size_t)gy * Wi + gx) * kC + 8 * oct);
          a[k] = __ldg(p);
          b[k] = __ldg(p + 1);
        }
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * NT;
        if (i >= TASKS) continue;
        const int oct = i & 3, px = i >> 2;
        const int lx = px % A_IN_COLS, ly = px / A_IN_COLS;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (inb[k]) {
          v[0] = a[k].x; v[1] = a[k].y; v[2] = a[k].z; v[3] = a[k].w;
          v[4] = b[k].x; v[5] = b[k].y; v[6] = b[k].z; v[7] = b[k].w;
        }
        const int phase = (ly & 1) * 2 + (lx & 1);
        const int pos = (ly >> 1) * A_PW + (lx >> 1);
        uint4 hi, lo;
        tc::split8(v, &hi, &lo);
        uint8_t* dst = s_pl + (size_t)(phase * 4 + oct) * A_PLANE + (size_t)pos * 16;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + 16 * A_PLANE) = lo;
      }
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_wait(&s_wbar, 0u);   // the bulk-copied weights have landed
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_pl), A_PLANE, 128u);
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 1024u, 128u);
      constexpr uint32_t plane_u16 = A_PLANE >> 4;
      int blk = 0;
#pragma unroll
      for (int phase = 0; phase < 4; ++phase) {
        const int na = (phase >> 1) == 0 ? 3 : 2, nb = (phase & 1) == 0 ? 3 : 2;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            if (a < na && b < nb) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t a_hi = da0 + (uint64_t)((phase * 4 + 2 * ks) * plane_u16 + a * A_PW + b);
                const uint64_t a_lo = a_hi + (uint64_t)(16 * plane_u16);
                const uint64_t bd = db0 + (uint64_t)(blk * 128);
                tc::mma_f16(tmem_base, a_hi, bd, tc::idesc_f16(64), blk != 0 ? 1u : 0u);
                tc::mma_f16(tmem_base, a_lo, bd, tc::idesc_f16(32), 1u);
                ++blk;
              }
            }
          }
        }
      }
      tc::mma_commit(&s_bar);
      tc::mbar_wait(&s_bar, 0u);
      tc::fence_before_sync();
    }
    __syncwarp();
  }
  __syncthreads();
  tc::fence_after_sync();

  // ---- epilogue: thread = (output position, 16 channels) ----
  {
    const int wq = warp & 3, half = warp >> 2;
    const int j = wq * 32 + lane;
    const int oy = oy0 + j / A_PW, oxl = j % A_PW, ox = ox0 + oxl;
    float v[16], c[16];
    const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * 16);
    tc::tmem_ld16(ta, v);
    tc::tmem_ld16(ta + 32u, c);
    if (oxl < A_TW && ox < Wo && oy < Ho) {
      float* o = out + (((size_t)img * Ho + oy) * Wo + ox) * kC + half * 16;
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] += c[k];
      st8(o, v);   // 256-bit stores: a thread's 64 bytes in two instructions (half the L1 wavefronts of four float4)
      st8(o + 8, v + 8);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 64u);
}

// ------------------------------------------------------------------------------------------------------------
// 3 -> 32 (planar input)
// ------------------------------------------------------------------------------------------------------------
constexpr int B_TW = 64;                           // output tile: TH x 64 positions = TH / 2 M-tiles
constexpr int B_IC = 2 * B_TW + 3;
constexpr int B_PITCH = B_IC + 2;                  // 133: odd pitch, rows start in different banks
constexpr int B_KSTEPS = 5;                        // K = 75 padded to 80
constexpr uint32_t B_WBYTES = B_KSTEPS * 2048u;
constexpr uint32_t B_CHUNK = 128u * 16u;           // one 8-wide K chunk of an M-tile: 128 rows x 16 B
constexpr uint32_t B_ATILE = 2u * B_KSTEPS * B_CHUNK;   // one M-tile of the operand (hi or lo)
template <int TH>
struct BCfg {
  static constexpr int MT = TH * B_TW / 128;
  static constexpr int IR = 2 * TH + 3;
  static constexpr int PLANE = IR * B_PITCH;
  static constexpr uint32_t SMEM = B_WBYTES + 2u * MT * B_ATILE + 3u * PLANE * 4u;
};

template <int TH, int MINB>
__global__ void __launch_bounds__(NT, MINB) conv5x5s2_c3_tc_kernel(const float* __restrict__ in,
                                                                 const uint8_t* __restrict__ w16, int Hi, int Wi,
                                                                 int Ho, int Wo, float* __restrict__ out) {
  constexpr int B_TH = TH, B_MT = BCfg<TH>::MT, B_IR = BCfg<TH>::IR, B_PLANE = BCfg<TH>::PLANE;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + B_WBYTES;                                   // [hi | lo][M-tile][chunk 10][128][8 fp16]
  float* s_t = reinterpret_cast<float*>(smem + B_WBYTES + 2u * B_MT * B_ATILE);   // [3][B_IR][B_PITCH]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int img = blockIdx.y;
  const int tiles_x = cdiv(Wo, B_TW);
  const int ox0 = (blockIdx.x % tiles_x) * B_TW, oy0 = (blockIdx.x / tiles_x) * B_TH;

  if (warp == 0) tc::tmem_alloc(&s_tmem, (uint32_t)(B_MT * 64));
  if (tid == 32) {
    tc::mbar_init(&s_bar, 1);
    tc::mbar_init_fence();
  }
  __syncthreads();
  pdl_launch_dependents();
  {
    const uint4* src = reinterpret_cast<const uint4*>(w16);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < (int)(B_WBYTES / 16); i += NT) dst[i] = __ldg(src + i);
  }
  pdl_wait();
  // ---- input tile, zero padded ----
  {
    const size_t plane = (size_t)Hi * Wi;
    const float* base = in + (size_t)img * 3 * plane;
    const int iy0 = 2 * oy0 - 2, ix0 = 2 * ox0 - 2;
    for (int row = warp; row < 3 * B_IR; row += NT / 32) {   // one warp per tile row: coalesced, no index divisions
      const int c = row / B_IR, y = row % B_IR;
      const int gy = iy0 + y;
      const bool rowok = gy >= 0 && gy < Hi;
      const float* rp = base + c * plane + (size_t)(rowok ? gy : 0) * Wi;
      float* dp = s_t + (c * B_IR + y) * B_PITCH;
#pragma unroll
      for (int x0 = 0; x0 < B_IC; x0 += 32) {
        const int x = x0 + lane, gx = ix0 + x;
        if (x < B_IC) dp[x] = (rowok && gx >= 0 && gx < Wi) ? __ldg(rp + gx) : 0.f;
      }
    }
  }
  __syncthreads();
  // ---- im2col: task = (output position, half of the K chunks), k = tap * 3 + channel; the half is warp-uniform ----
  auto build = [&](int pos, auto first_chunk) {
    constexpr int J0 = decltype(first_chunk)::value;
    const int oyl = pos / B_TW, oxl = pos % B_TW;
    const float* src = s_t + (2 * oyl) * B_PITCH + 2 * oxl;
    const int mt = pos >> 7, m = pos & 127;
    uint8_t* dst = s_a + (size_t)mt * B_ATILE + (size_t)(m >> 3) * 128 + (size_t)(m & 7) * 16;
#pragma unroll
    for (int j = J0; j < J0 + B_KSTEPS; ++j) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = 8 * j + e;
        if (k < 75) {
          const int tap = k / 3, c = k % 3;
          v[e] = src[c * B_PLANE + (tap / 5) * B_PITCH + (tap % 5)];
        } else {
          v[e] = 0.f;
        }
      }
      uint4 hi, lo;
      tc::split8(v, &hi, &lo);
      *reinterpret_cast<uint4*>(dst + (size_t)j * B_CHUNK) = hi;
      *reinterpret_cast<uint4*>(dst + (size_t)j * B_CHUNK + (size_t)B_MT * B_ATILE) = lo;
    }
  };
  for (int task = tid; task < 2 * B_TH * B_TW; task += NT) {
    const int pos = task % (B_TH * B_TW);
    if (task < B_TH * B_TW) build(pos, std::integral_constant<int, 0>{});
    else build(pos, std::integral_constant<int, B_KSTEPS>{});
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_a), B_CHUNK, 128u);
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 1024u, 128u);
#pragma unroll
      for (int mt = 0; mt < B_MT; ++mt) {
#pragma unroll
        for (int ks = 0; ks < B_KSTEPS; ++ks) {
          const uint64_t a_hi = da0 + (uint64_t)((mt * B_ATILE + 2 * ks * B_CHUNK) >> 4);
          const uint64_t a_lo = a_hi + (uint64_t)((B_MT * B_ATILE) >> 4);
          const uint64_t bd = db0 + (uint64_t)(ks * 128);
          tc::mma_f16(tmem_base + (uint32_t)(mt * 64), a_hi, bd, tc::idesc_f16(64), ks != 0 ? 1u : 0u);
          tc::mma_f16(tmem_base + (uint32_t)(mt * 64), a_lo, bd, tc::idesc_f16(32), 1u);
        }
      }
      tc::mma_commit(&s_bar);
      tc::mbar_wait(&s_bar, 0u);
      tc::fence_before_sync();
    }
    __syncwarp();
  }
  __syncthreads();
  tc::fence_after_sync();

  // ---- epilogue: thread = (output position, 16 channels), two M-tiles per warp ----
  {
    const int wq = warp & 3, half = warp >> 2;
#pragma unroll
    for (int mt = 0; mt < B_MT; ++mt) {
      const int pos = mt * 128 + wq * 32 + lane;
      const int oy = oy0 + pos / B_TW, ox = ox0 + pos % B_TW;
      float v[16], c[16];
      const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * 64 + half * 16);
      tc::tmem_ld16(ta, v);
      tc::tmem_ld16(ta + 32u, c);
      if (ox < Wo && oy < Ho) {
        float* o = out + (((size_t)img * Ho + oy) * Wo + ox) * kC + half * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] += c[k];
        st8(o, v);
        st8(o + 8, v + 8);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, (uint32_t)(B_MT * 64));
}

}  // namespace

// Block order = the issue order of the kernel: phase (py, px) major, then (a, b), then k-step.
void pack_conv5_c32_weights(const float* w_oihw, std::vector<uint8_t>* out) {
  out->assign(A_WBYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  int blk = 0;
  for (int phase = 0; phase < 4; ++phase) {
    const int py = phase >> 1, px = phase & 1;
    for (int a = 0; a < (py == 0 ? 3 : 2); ++a)
      for (int b = 0; b < (px == 0 ? 3 : 2); ++b) {
        const int ky = 2 * a + py, kx = 2 * b + px;
        for (int ks = 0; ks < 2; ++ks, ++blk)
          for (int k = 0; k < 16; ++k)
            for (int n = 0; n < 32; ++n)
              tc::put_split_weight(h, blk, k, n, w_oihw[(((size_t)n * 32 + ks * 16 + k) * 5 + ky) * 5 + kx]);
      }
  }
}

void pack_conv5_c3_weights(const float* w_oihw, std::vector<uint8_t>* out) {
  out->assign(B_WBYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  for (int kk = 0; kk < 75; ++kk) {
    const int tap = kk / 3, c = kk % 3;
    for (int n = 0; n < 32; ++n)
      tc::put_split_weight(h, kk / 16, kk % 16, n, w_oihw[((size_t)n * 3 + c) * 25 + tap]);
  }
}

int launch_conv5x5s2_c32_tc(const float* in, const uint8_t* w16, int n, int Hi, int Wi, float* out,
                            cudaStream_t stream) {
  if (n <= 0) return 0;
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&conv5x5s2_c32_tc_kernel), A_SMEM)) return rc;
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  dim3 grid(cdiv(Wo, A_TW) * cdiv(Ho, A_TH), n);
  launch_pdl(conv5x5s2_c32_tc_kernel, grid, dim3(NT), (size_t)A_SMEM, stream, in, w16, Hi, Wi, Ho, Wo, out);
  B200MVS_LAUNCH_OK("conv5x5s2_c32_tc_kernel");
  return 0;
}

template <int TH, int MINB>
int launch_c3(const float* in, const uint8_t* w16, int n, int Hi, int Wi, float* out, cudaStream_t stream) {
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&conv5x5s2_c3_tc_kernel<TH, MINB>), BCfg<TH>::SMEM)) return rc;
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  dim3 grid(cdiv(Wo, B_TW) * cdiv(Ho, TH), n);
  launch_pdl(conv5x5s2_c3_tc_kernel<TH, MINB>, grid, dim3(NT), (size_t)BCfg<TH>::SMEM, stream, in, w16, Hi, Wi, Ho, Wo, out);
  B200MVS_LAUNCH_OK("conv5x5s2_c3_tc_kernel");
  return 0;
}

int launch_conv5x5s2_c3_tc(const float* in, const uint8_t* w16, int n, int Hi, int Wi, float* out,
                           cudaStream_t stream) {
  if (n <= 0) return 0;
  const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
  // Few image groups: short tiles (three CTAs per SM, fine-grained tail); many: tall tiles amortise the weights.
  const long long tall = (long long)cdiv(Wo, B_TW) * cdiv(Ho, 8) * n;
  if (tall >= 296) return launch_c3<8, 1>(in, w16, n, Hi, Wi, out, stream);
  return launch_c3<2, 3>(in, w16, n, Hi, Wi, out, stream);
}

}  // namespace b200mvs
