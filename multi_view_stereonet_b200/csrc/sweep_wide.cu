// The depth-sweep feature recurrence (multi_view_stereonet.py:279-290) for 1/16-scale images that are too large for
// one thread-block cluster (recurrence.cu holds a whole image in the shared memory of <= 16 CTAs; BASELINE cfg5's
// 64 x 80 image is 41 M-tiles): ONE persistent launch whose CTAs are all co-resident (cooperative launch), every
// CTA owning up to five consecutive 128-position M-tiles of one (image group, view) chain for all D-1 dependent steps.
//
// Same arithmetic as recurrence.cu (pitch PW = cols + 2, split-fp16 operands, three products per k-step as an
// N=64 + an N=32 tcgen05.mma, fp32 accumulation in TMEM, image half of conv0 precomputed by image_conv_kernel,
// gather plan precomputed by gather_plan_kernel, weights resident in shared memory).  What differs is the exchange
// between the CTAs of a chain, which goes through L2 instead of distributed shared memory:
//   * the previous hypothesis is gathered straight from the feature volume in global memory (L1-cached loads: a
//     hypothesis is written once and only read behind the chain barrier that published it);
//   * per normalised layer every CTA writes its GroupNorm partial (sum, sumsq) per group into its own slot and the raw
//     outputs of the positions within PW + 1 of its range boundaries into a scratch row, then arrives at the chain's
//     counter (release); behind the counter (acquire) every CTA adds the slots in one fixed order (bit-identical
//     coefficients in all CTAs, no atomics on data) and normalises its own tiles (accumulators kept in registers)
//     and the neighbours' boundary rows;
//   * the warped features of the own positions (needed again by the last epilogue, 80 KB at five tiles: no room next
//     to 108 KB of weights and 103 KB of operand planes) take a round trip through an L2-resident scratch buffer.
// Three chain-wide barriers per step (~2 us each) against 3 x 180 MMAs per CTA: with 16 chains in flight (cfg5's
// per-GPU share) the kernel is throughput-, not latency-bound, which is what distinguishes it from recurrence.cu.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "conv.cuh"
#include "recurrence.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int NT = 512;          // worker threads: 16 warps = (4 lane quarters of an M-tile) x (4 channel octets)
constexpr int NT_ALL = NT + 32;  // + a warp that only issues the MMAs (tcgen05.mma issue blocks at the rate the tensor pipe
                                 //   drains: issued from worker warp 0, that warp started every epilogue ~9 k cycles late
                                 //   and the other fifteen waited for it at the next block barrier, ~3 k cycles per conv)
constexpr int MTILE = 128;
constexpr int MAX_MT = 5;        // M-tiles per CTA: 5 x 64 accumulator columns <= 512 TMEM columns, planes <= 117 KB
constexpr int W_BLOCKS = 9 * 2;  // taps x k-steps of one layer
constexpr uint32_t W_TOTAL_BYTES = 3u * W_BLOCKS * 2048u;   // three layers, [W_hi | W_lo] blocks (pack_recurrence_weights)
constexpr uint32_t kSmemBudget = 227u * 1024u - 4u * 1024u;  // dynamic part; the static arrays take the rest
constexpr int NUM_PLANES = 8;    // [hi f0..f3][lo f0..f3], each npl_pad x 16 B
constexpr int PLANE_HI = 0, PLANE_LO = 4;

struct WideLayout {
  int PW, halo, npl, npl_pad;
  uint32_t plane_bytes, off_w, off_planes, total;
};
__host__ __device__ inline WideLayout make_wide_layout(int cols, int mt) {
  WideLayout L;
  L.PW = cols + 2;
  L.halo = L.PW + 1;
  L.npl = mt * MTILE + 2 * L.halo;
  L.npl_pad = ((L.npl + 5) & ~7) + 2;   // 2 (mod 8): the four octet planes a quad writes start in distinct banks
  L.plane_bytes = (uint32_t)L.npl_pad * 16u;
  L.off_w = 0;
  L.off_planes = W_TOTAL_BYTES;
  L.total = L.off_planes + NUM_PLANES * L.plane_bytes;
  return L;
}

struct WideParams {
  const float* vol_in;   // feature volume [n][D][rows*cols][32]; hypothesis 0 filled
  float* vol;            // same buffer
  const uint8_t* w16;    // pack_recurrence_weights
  const float *bias1, *bias2, *gamma0, *beta0, *gamma1, *beta1;
  const float* imgconv;  // [n][D][4 octets][rows*cols][8]: image half of conv0 + bias0 (octet-major)
  const float4* plan;    // [n][D][plan_stride]
  int plan_stride;
  float* wfbuf;          // [n][4 octets][npos][8] warped features of the current step (octet-major: an epilogue warp
                         // reads 32 positions x 32 B contiguously)
  float* ybuf;           // [n][2][npos][32] raw layer outputs next to the CTA boundaries
  float2* part;          // [n][2][T_max][4] GroupNorm partials per CTA
  unsigned* ctr;         // [n] arrivals at the chain barrier (zeroed before the launch)
  unsigned* abort_flag;  // raised by a CTA whose chain barrier did not complete within kSpinLimit cycles
  int part_stride;       // T_max
  int npos;              // tiles * 128
  int chain0;            // first chain of this launch
  int D, rows, cols, tiles, T;
  long long* prof;       // optional [16] phase cycle totals of CTA (0, 0)
  int debug;             // 32: CTA 0 of chain 0 skips its first arrival (watchdog test: the chain's barrier times out)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// The chain barrier's wait (one thread per CTA).  A spin barrier is only safe while every CTA of the chain is resident;
// the cooperative launch guarantees that, but a wait that can never end would hang the GPU instead of reporting an
// error, so it carries a watchdog: after kSpinLimit cycles (~1 s) the CTA raises the launch's abort flag, every CTA
// that sees the flag stops waiting, the kernel ends (with garbage) and the host reports the failure (api.cu).
constexpr long long kSpinLimit = 2000000000ll;
__device__ __forceinline__ bool chain_wait(const unsigned* ctr, unsigned target, unsigned* abort_flag) {
  unsigned it = 0;
  long long t0 = 0;
  while (ld_acquire_u32(ctr) < target) {
    if ((++it & 1023u) == 0) {
      if (it == 1024u) t0 = clock64();
      if (*reinterpret_cast<volatile unsigned*>(abort_flag) != 0) return false;
      if (clock64() - t0 > kSpinLimit) {
        atomicExch(abort_flag, 1u);
        return false;
      }
    }
  }
  return true;
}
// arrival at a chain counter: one release reduction (everything this thread wrote, and what the block barrier in front
// of it made it observe, is ordered before the increment)
__device__ __forceinline__ void red_release_add(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

__device__ __forceinline__ void unsplit8(const uint4& hi, const uint4& lo, float* v) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[k]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l[k]));
    v[2 * k] = a.x + b.x;
    v[2 * k + 1] = a.y + b.y;
  }
}

// Two 8-column slices of this thread's TMEM lane (the hi*hi + lo*hi and the hi*lo accumulators).
__device__ __forceinline__ void tmem_ld8x2(uint32_t taddr0, uint32_t taddr1, float* v, float* c) {
  uint32_t r[8], q[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr0)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
               : "r"(taddr1)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = __uint_as_float(r[i]);
    c[i] = __uint_as_float(q[i]);
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// registers -> this thread's TMEM lane (the caller issues tcgen05.wait::st before it reads them back)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

// The 36 MMAs of one M-tile of a 3x3 conv (see recurrence.cu: issue_conv_mmas).
__device__ __forceinline__ void issue_tile_mmas(uint64_t da_hi, uint64_t da_lo, uint64_t db, uint32_t plane_u16,
                                                uint32_t PW, uint32_t d_tmem) {
  constexpr uint32_t kN64 = tc::idesc_f16(64), kN32 = tc::idesc_f16(32);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t pos = (uint32_t)(tap / 3) * PW + (uint32_t)(tap % 3);   // 16-byte units
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t a_off = (uint64_t)(2u * ks * plane_u16 + pos);
      const uint64_t b = db + (uint64_t)((tap * 2 + ks) * (2048 / 16));
      tc::mma_f16(d_tmem, da_hi + a_off, b, kN64, (tap | ks) != 0 ? 1u : 0u);
      tc::mma_f16(d_tmem, da_lo + a_off, b, kN32, 1u);
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(NT_ALL, 1) sweep_wide_kernel(const WideParams p) {
  constexpr int ITERS = MT + 2;   // gather tasks per thread: (MT * 128 + 2 halo) * 4 / 512, halo <= 128
  constexpr uint32_t TMEM_COLS = MT == 1 ? 64u : (MT == 2 ? 128u : (MT <= 4 ? 256u : 512u));
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar[MAX_MT];   // MMA completion per M-tile
  __shared__ __align__(8) uint64_t s_wbar;          // bulk copy of the weights
  __shared__ __align__(16) float2 s_loc[kGroups][4];
  __shared__ __align__(16) float s_ca[kC], s_cb[kC];
  __shared__ __align__(16) float s_bias[2][kC], s_gamma[2][kC], s_beta[2][kC];
  __shared__ long long s_prof[10];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int cta = blockIdx.x, chain = p.chain0 + blockIdx.y;
  const WideLayout L = make_wide_layout(p.cols, MT);
  const int PW = L.PW, halo = L.halo;
  const int pixels = p.rows * p.cols, npos_img = p.rows * PW;
  const int nmy = min(MT, p.tiles - cta * MT);   // >= 1: T = ceil(tiles / MT)
  const int pos0 = cta * MT * MTILE;             // first own output position = input position of local l = 0
  const int own_n = nmy * MTILE;
  const int npl_my = own_n + 2 * halo;

  uint8_t* s_w = smem + L.off_w;
  uint8_t* s_planes = smem + L.off_planes;

  if (warp == 0) tc::tmem_alloc(&s_tmem, TMEM_COLS);
  if (tid == 32) {
#pragma unroll
    for (int t = 0; t < MAX_MT; ++t) tc::mbar_init(&s_bar[t], 1);
    tc::mbar_init(&s_wbar, 1);
    tc::mbar_init_fence();
  }
  if (tid < kC) {
    s_bias[0][tid] = __ldg(p.bias1 + tid);
    s_bias[1][tid] = __ldg(p.bias2 + tid);
    s_gamma[0][tid] = __ldg(p.gamma0 + tid);
    s_beta[0][tid] = __ldg(p.beta0 + tid);
    s_gamma[1][tid] = __ldg(p.gamma1 + tid);
    s_beta[1][tid] = __ldg(p.beta1 + tid);
  }
  {
    uint4* pl = reinterpret_cast<uint4*>(s_planes);   // padding positions stay zero
    for (int i = tid; i < NUM_PLANES * L.npl_pad; i += NT_ALL) pl[i] = make_uint4(0, 0, 0, 0);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) tc::bulk_load_weights(s_w, p.w16, W_TOTAL_BYTES, &s_wbar);
  const uint32_t tmem_base = s_tmem;

  const float inv_count = 1.0f / (8.0f * (float)pixels);

  uint32_t conv_phase = 0;
  // ---- the issuing warp: sleeps at named barrier 2 until the 512 workers have staged an operand (they only ARRIVE
  //      there and go on to the tiles' completion barriers), issues every tile's MMAs from one elected lane, one
  //      commit per tile so that a tile's epilogue runs under the following tiles' MMAs
  if (warp == NT / 32) {
    tc::mbar_wait(&s_wbar, 0);   // the bulk-copied weights have landed
    for (int step = 1; step < p.D; ++step) {
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        asm volatile("bar.sync 2, %0;" ::"n"(NT_ALL) : "memory");
        if (tc::elect_one()) {
          tc::fence_after_sync();
          const uint32_t plane_u16 = L.plane_bytes >> 4;
          const uint64_t da_hi0 = tc::umma_desc(tc::smem_u32(s_planes) + (uint32_t)PLANE_HI * L.plane_bytes, L.plane_bytes, 128u);
          const uint64_t da_lo0 = tc::umma_desc(tc::smem_u32(s_planes) + (uint32_t)PLANE_LO * L.plane_bytes, L.plane_bytes, 128u);
          const uint64_t db = tc::umma_desc(tc::smem_u32(s_w), 1024u, 128u) + (uint64_t)(layer * W_BLOCKS * (2048 / 16));
#pragma unroll 1
          for (int t = 0; t < nmy; ++t) {
            issue_tile_mmas(da_hi0 + (uint64_t)(t * MTILE), da_lo0 + (uint64_t)(t * MTILE), db, plane_u16, (uint32_t)PW,
                            tmem_base + (uint32_t)(t * 64));
            tc::mma_commit(&s_bar[t]);
          }
        }
        __syncwarp();
      }
    }
    tc::fence_before_sync();
    __syncthreads();   // the kernel's last block barrier (TMEM is freed behind it)
    return;
  }
  // workers: the operand is staged (generic-proxy writes fenced here) -> wake the issuing warp, do not wait
  auto issue_conv = [&](int) {
    tc::fence_proxy_async();
    tc::fence_before_sync();
    asm volatile("bar.arrive 2, %0;" ::"n"(NT_ALL) : "memory");
  };
  auto sync_workers = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); };

  // ---- chain barrier: arrivals counted in global memory, one polling thread per CTA ----
  unsigned* const ctr = p.ctr + chain;
  unsigned sync_k = 0;
  const unsigned T = (unsigned)p.T;
  bool alive = true;   // (thread 0) no chain barrier has timed out

  // This thread's accumulator slices: tile t, lane quarter wq, ONE channel octet (= one GroupNorm group).
  const int wq = warp & 3, oct_e = warp >> 2;
  int own_pix[MT];    // pixel index of the own output position of tile t, -1: pad column / beyond the image / no tile
  bool own_bnd[MT];   // within `halo` positions of the CTA's range boundaries: the neighbours normalise it too
#pragma unroll
  for (int t = 0; t < MT; ++t) {
    const int jl = t * MTILE + wq * 32 + lane, jg = pos0 + jl;
    const int oy = jg / PW, ox = jg - oy * PW;
    const bool real = t < nmy && ox < p.cols && oy < p.rows;
    own_pix[t] = real ? oy * p.cols + ox : -1;
    own_bnd[t] = real && (jl < halo || jl >= own_n - halo);
  }
  const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(oct_e * 8);
  auto plane_ptr = [&](int plane, int l) -> uint4* {
    return reinterpret_cast<uint4*>(s_planes + (size_t)plane * L.plane_bytes + (size_t)l * 16);
  };

  // gather tasks: (local input position l, channel octet); the four lanes of a quad share a position; lane j of the
  // quad keeps the plan entries of iterations k = j, j + 4 and hands them out by shuffle
  const int t_oct = tid & 3;
  constexpr int PLAN_REGS = (ITERS + 3) / 4;
  const float4* const plan_base = p.plan + (size_t)chain * p.D * p.plan_stride + pos0;
  float4 g_plan[PLAN_REGS];
  auto load_plan = [&](int step) {
#pragma unroll
    for (int i = 0; i < PLAN_REGS; ++i) {
      const int k = t_oct + 4 * i;
      const int l = (tid + k * NT) >> 2;
      g_plan[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < ITERS && l < npl_my) g_plan[i] = __ldg(plan_base + (size_t)step * p.plan_stride + l);
    }
  };
  auto plan_of = [&](int k) -> float4 {   // k is a compile-time constant after unrolling
    const float4 src = g_plan[k >> 2];
    const int from = (lane & ~3) | (k & 3);
    float4 r;
    r.x = __shfl_sync(0xffffffffu, src.x, from);
    r.y = __shfl_sync(0xffffffffu, src.y, from);
    r.z = __shfl_sync(0xffffffffu, src.z, from);
    r.w = __shfl_sync(0xffffffffu, src.w, from);
    return r;
  };

  float* const wf_chain = p.wfbuf + (size_t)chain * p.npos * kC;
  float* const y_chain = p.ybuf + (size_t)chain * 2 * p.npos * kC;
  float2* const part_chain = p.part + (size_t)chain * 2 * p.part_stride * kGroups;

  const bool prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
  long long t_prev = prof ? clock64() : 0;
  if (tid < 10) s_prof[tid] = 0;   // (ordered before the first mark by the block barriers of the first step)
#define WIDE_MARK(k)                  \
  do {                                \
    if (prof) {                       \
      const long long _t = clock64(); \
      s_prof[k] += _t - t_prev;       \
      t_prev = _t;                    \
    }                                 \
  } while (0)

  load_plan(1);

  for (int step = 1; step < p.D; ++step) {
    // ================= W: warp the previous hypothesis into the conv0 operand =================
    if (step >= 2) {   // every CTA of the chain has published hypothesis step - 1
      ++sync_k;
      if (tid == 0 && alive) alive = chain_wait(ctr, T * sync_k, p.abort_flag);
      sync_workers();
    }
    WIDE_MARK(0);
    {
      const float* prev = p.vol_in + ((size_t)chain * p.D + (step - 1)) * pixels * kC + 8 * t_oct;
#pragma unroll
      for (int k = 0; k < ITERS; ++k) {
        const int l = (tid + k * NT) >> 2;
        const float4 e = plan_of(k);
        if (l < npl_my) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
          const int fl = __float_as_int(e.w);
          if (fl & 1) {
            const float we = e.y, ws = e.z;
            const float ww = 1.0f - we, wn = 1.0f - ws;
            const float wt[4] = {wn * ww, wn * we, ws * ww, ws * we};
            const int P00 = __float_as_int(e.x);   // output position of the north-west tap
            const int y0 = P00 / PW, x0 = P00 - y0 * PW;
            const int dx = (fl >> 1) & 1, dy = (fl >> 2) & 1;
            const float* p00 = prev + ((size_t)y0 * p.cols + x0) * kC;
            const float* p10 = p00 + (size_t)dy * p.cols * kC;
            const float* tp[4] = {p00, p00 + dx * kC, p10, p10 + dx * kC};
            float a[4][8];
#pragma unroll
            for (int t = 0; t < 4; ++t) ld8(tp[t], a[t]);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaf(a[t][i], wt[t], v[i]);
          }
          if (l >= halo && l < halo + own_n) {
            st8_cg(wf_chain + ((size_t)t_oct * p.npos + (size_t)(pos0 + l - halo)) * 8, v);
          }
          uint4 hi, lo;
          tc::split8(v, &hi, &lo);
          *plane_ptr(PLANE_HI + t_oct, l) = hi;
          *plane_ptr(PLANE_LO + t_oct, l) = lo;
        }
      }
    }
    issue_conv(0);
    WIDE_MARK(1);

    // ===== two normalised layers: raw output -> statistics + boundary rows -> chain barrier -> operand -> conv =====
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      float gs = 0.f, gq = 0.f;
      float* const y_layer = y_chain + (size_t)layer * p.npos * kC;
      // what is added to the accumulators (image half of conv0 + bias0, or bias1), loaded one tile ahead; the raw
      // output goes back into the tile's first accumulator columns (tcgen05.st) and is read again behind the chain
      // barrier, so that nothing per tile stays in registers across it
      auto load_add = [&](int t, float* nx) {
#pragma unroll
        for (int k = 0; k < 8; ++k) nx[k] = 0.f;
        if (layer == 0) {   // [n][D][octet][pixel][8] (launch_image_conv, oct_major)
          if (own_pix[t] >= 0) ld8_nc(p.imgconv + ((((size_t)chain * p.D + step) * 4 + oct_e) * pixels + own_pix[t]) * 8, nx);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) nx[k] = s_bias[0][oct_e * 8 + k];
        }
      };
      float nx[8];
      load_add(0, nx);
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        if (t < nmy) {
          float add[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) add[k] = nx[k];
          if (t + 1 < MT) load_add(t + 1 < MT ? t + 1 : 0, nx);
          float y[8], c[8];
          tc::mbar_wait(&s_bar[t], conv_phase);
          tc::fence_after_sync();
          tmem_ld8x2(tmem_my + (uint32_t)(t * 64), tmem_my + (uint32_t)(t * 64) + 32u, y, c);
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = (y[k] + c[k]) + add[k];
          tmem_st8(tmem_my + (uint32_t)(t * 64), y);
          if (own_pix[t] >= 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              gs += y[k];
              gq += y[k] * y[k];
            }
          }
          if (own_bnd[t]) {
            st8_cg(y_layer + (size_t)(pos0 + t * MTILE + wq * 32 + lane) * kC + oct_e * 8, y);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      conv_phase ^= 1u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, o);
        gq += __shfl_xor_sync(0xffffffffu, gq, o);
      }
      if (lane == 0) s_loc[oct_e][wq] = make_float2(gs, gq);
      tc::fence_before_sync();
      sync_workers();
      WIDE_MARK(2 + 3 * layer);
      ++sync_k;
      if (tid == 0) {
        float2* slot = part_chain + ((size_t)layer * p.part_stride + cta) * kGroups;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const float2 u0 = s_loc[g][0], u1 = s_loc[g][1], u2 = s_loc[g][2], u3 = s_loc[g][3];
          __stcg(slot + g, make_float2((u0.x + u1.x) + (u2.x + u3.x), (u0.y + u1.y) + (u2.y + u3.y)));
        }
        if (!((p.debug & 32) && cta == 0 && chain == 0 && sync_k == 1)) red_release_add(ctr);
        if (alive) alive = chain_wait(ctr, T * sync_k, p.abort_flag);
      }
      sync_workers();
      WIDE_MARK(3 + 3 * layer);
      // the neighbours' boundary rows this thread normalises (two tasks at most: 2 halo * 4 <= 2 * NT): in flight under
      // the coefficients and the own rows
      float hy[2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int h = (tid + i * NT) >> 2;
        const int q = pos0 + (h < halo ? h : h + own_n) - halo;   // the output position the halo position mirrors
#pragma unroll
        for (int k = 0; k < 8; ++k) hy[i][k] = 0.f;
        if (h < 2 * halo && q >= 0 && q < npos_img && q % PW < p.cols) ld8_cg(y_layer + (size_t)q * kC + t_oct * 8, hy[i]);
      }
      // ---- GroupNorm coefficients: warp g adds the chain's slots of group g in one fixed order ----
      if (warp < kGroups) {
        float ts = 0.f, tq = 0.f;
        const float2* slots = part_chain + (size_t)layer * p.part_stride * kGroups + warp;
        for (int c = lane; c < p.T; c += 32) {
          const float2 u = __ldcg(slots + (size_t)c * kGroups);
          ts += u.x;
          tq += u.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ts += __shfl_xor_sync(0xffffffffu, ts, o);
          tq += __shfl_xor_sync(0xffffffffu, tq, o);
        }
        const double mean = (double)ts * (double)inv_count;
        const double var = (double)tq * (double)inv_count - mean * mean;   // cancellation in double
        const float rstd = rsqrtf(fmaxf((float)var, 0.f) + kGnEps);
        if (lane < 8) {
          const int ch = warp * 8 + lane;
          const float ca = s_gamma[layer][ch] * rstd;
          s_ca[ch] = ca;
          s_cb[ch] = s_beta[layer][ch] - (float)mean * ca;
        }
      }
      sync_workers();
      // ---- next operand: x = lrelu(GN(y)) (+ x0 in the residual block) over own + halo positions ----
      {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_ca[oct_e * 8]), a1 = *reinterpret_cast<const float4*>(&s_ca[oct_e * 8 + 4]);
        const float4 c0 = *reinterpret_cast<const float4*>(&s_cb[oct_e * 8]), c1 = *reinterpret_cast<const float4*>(&s_cb[oct_e * 8 + 4]);
        const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float cb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          if (t < nmy) {
            const int own_l = halo + t * MTILE + wq * 32 + lane;
            uint4* ph = plane_ptr(PLANE_HI + oct_e, own_l);
            uint4* plo = plane_ptr(PLANE_LO + oct_e, own_l);
            float x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = 0.f;
            float y[8];
            tmem_ld8(tmem_my + (uint32_t)(t * 64), y);   // (.sync.aligned: the whole warp, outside the per-lane branch)
            if (own_pix[t] >= 0) {
              float xprev[8];
              if (layer == 1) unsplit8(*ph, *plo, xprev);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                x[k] = lrelu(fmaf(y[k], ca[k], cb[k]));
                if (layer == 1) x[k] += xprev[k];
              }
            }
            uint4 hi, lo;
            tc::split8(x, &hi, &lo);
            *ph = hi;
            *plo = lo;
          }
        }
      }
      {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_ca[t_oct * 8]), a1 = *reinterpret_cast<const float4*>(&s_ca[t_oct * 8 + 4]);
        const float4 c0 = *reinterpret_cast<const float4*>(&s_cb[t_oct * 8]), c1 = *reinterpret_cast<const float4*>(&s_cb[t_oct * 8 + 4]);
        const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float cb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int h = (tid + i * NT) >> 2;
          if (h < 2 * halo) {
            const int l = h < halo ? h : h + own_n;   // lower halo, then upper halo
            const int q = pos0 + l - halo;
            uint4* ph = plane_ptr(PLANE_HI + t_oct, l);
            uint4* plo = plane_ptr(PLANE_LO + t_oct, l);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
            if (q >= 0 && q < npos_img && q % PW < p.cols) {
              const float* yy = hy[i];
              float xprev[8];
              if (layer == 1) unsplit8(*ph, *plo, xprev);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                v[e] = lrelu(fmaf(yy[e], ca[e], cb[e]));
                if (layer == 1) v[e] += xprev[e];
              }
            }
            uint4 hi, lo;
            tc::split8(v, &hi, &lo);
            *ph = hi;
            *plo = lo;
          }
        }
      }
      issue_conv(1 + layer);
      WIDE_MARK(4 + 3 * layer);
    }

    // ===== E2: features_step = warped + delta -> the feature volume =====
    const bool more = step + 1 < p.D;
    {
      float* dst_h = p.vol + ((size_t)chain * p.D + step) * pixels * kC + oct_e * 8;
      const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[1][oct_e * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[1][oct_e * 8 + 4]);
      auto load_wf = [&](int t, float* w) {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0.f;
        if (own_pix[t] >= 0) ld8_cg(wf_chain + ((size_t)oct_e * p.npos + (size_t)(pos0 + t * MTILE + wq * 32 + lane)) * 8, w);
      };
      float nw[2][8];   // two tiles ahead
      load_wf(0, nw[0]);
      if (MT > 1) load_wf(MT > 1 ? 1 : 0, nw[1]);
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        if (t < nmy) {
          float w[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) w[k] = nw[t & 1][k];
          if (t + 2 < MT) load_wf(t + 2 < MT ? t + 2 : 0, nw[t & 1]);
          tc::mbar_wait(&s_bar[t], conv_phase);
          tc::fence_after_sync();
          float v[8], c[8];
          tmem_ld8x2(tmem_my + (uint32_t)(t * 64), tmem_my + (uint32_t)(t * 64) + 32u, v, c);
          if (own_pix[t] >= 0) {
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = w[k] + ((v[k] + c[k]) + bb[k]);
            st8(dst_h + (size_t)own_pix[t] * kC, r);
          }
        }
      }
      conv_phase ^= 1u;
    }
    WIDE_MARK(8);
    if (more) load_plan(step + 1);
    tc::fence_before_sync();
    sync_workers();
    if (more && tid == 0) {
      red_release_add(ctr);
    }
    WIDE_MARK(9);
  }
  if (prof) {
    for (int k = 0; k < 10; ++k) p.prof[k] = s_prof[k];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}
#undef WIDE_MARK

struct WidePlan {
  int mt, T, chains_per_launch;
};

// All CTAs of a launch must be co-resident (one per SM: ~215 KB of shared memory): pick the fewest tiles per CTA
// that fit `chains_per_launch` whole chains on the device.
bool plan_wide(int rows, int cols, int n, int sms, WidePlan* out) {
  const WideLayout L5 = make_wide_layout(cols, MAX_MT);
  if (L5.halo > MTILE || rows < 1 || cols < 1) return false;   // boundary rows come from the adjacent CTA's first / last tile
  const int tiles = cdiv(rows * L5.PW, MTILE);
  int max_mt = MAX_MT;
  while (max_mt >= 1 && make_wide_layout(cols, max_mt).total > kSmemBudget) --max_mt;
  if (max_mt < 1) return false;
  const int t_min = cdiv(tiles, max_mt);
  if (t_min > sms) return false;
  const int per_wave = sms / t_min;
  const int waves = cdiv(n, per_wave);
  const int chunk = cdiv(n, waves);
  int mt = max_mt;
  for (int m = 1; m <= max_mt; ++m)
    if (chunk * cdiv(tiles, m) <= sms) {
      mt = m;
      break;
    }
  out->mt = mt;
  out->T = cdiv(tiles, mt);
  out->chains_per_launch = chunk;
  return true;
}

template <int MT>
int launch_mt(const WideParams& p, int chains, size_t smem, cudaStream_t stream) {
  const void* f = reinterpret_cast<const void*>(&sweep_wide_kernel<MT>);
  if (int rc = ensure_func_smem(f, smem)) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.T, chains, 1);
  cfg.blockDim = dim3(NT_ALL, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident, or the launch fails
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, sweep_wide_kernel<MT>, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(std::string("sweep_wide_kernel launch (") + std::to_string(p.T) + " x " + std::to_string(chains) +
              " CTAs): " + cudaGetErrorString(e));
    return -2;
  }
  note_launch();
  return 0;
}

}  // namespace

bool sweep_wide_supported(int rows, int cols) {
  int sms = 0;
  if (current_device_sm_count(&sms) != 0 || sms < 1) return false;   // (one chain must fit the device it will run on)
  WidePlan pl;
  return plan_wide(rows, cols, 1, sms, &pl);
}

void sweep_wide_scratch(int rows, int cols, int n, size_t* wf_floats, size_t* y_floats, size_t* part_float2,
                        size_t* counters) {
  const int tiles = cdiv(rows * (cols + 2), MTILE);
  const size_t npos = (size_t)tiles * MTILE;
  *wf_floats = (size_t)n * npos * kC;
  *y_floats = (size_t)n * 2 * npos * kC;
  *part_float2 = (size_t)n * 2 * tiles * kGroups;
  *counters = (size_t)n + 1;   // + the watchdog's abort flag
}

int launch_sweep_wide(const RecurrenceArgs& a, const WideScratch& s, cudaStream_t stream) {
  if (a.D < 2 || a.n <= 0) return 0;
  int sms = 0;
  if (int rc = current_device_sm_count(&sms)) return rc;
  WidePlan pl;
  if (!plan_wide(a.rows, a.cols, a.n, sms, &pl)) {
    set_error("launch_sweep_wide: shape not supported");
    return -1;
  }
  const WideLayout L = make_wide_layout(a.cols, pl.mt);
  const int tiles = cdiv(a.rows * L.PW, MTILE);
  B200MVS_CUDA_OK(cudaMemsetAsync(s.ctr, 0, ((size_t)a.n + 1) * sizeof(unsigned), stream));
  WideParams p;
  p.vol_in = a.vol;
  p.vol = a.vol;
  p.w16 = a.w16;
  p.bias1 = a.bias1;
  p.bias2 = a.bias2;
  p.gamma0 = a.gamma0;
  p.beta0 = a.beta0;
  p.gamma1 = a.gamma1;
  p.beta1 = a.beta1;
  p.imgconv = a.imgconv;
  p.plan = reinterpret_cast<const float4*>(a.plan);
  p.plan_stride = recurrence_plan_stride(a.rows, a.cols);
  p.wfbuf = s.wf;
  p.ybuf = s.y;
  p.part = reinterpret_cast<float2*>(s.part);
  p.ctr = s.ctr;
  p.abort_flag = s.ctr + a.n;
  p.part_stride = tiles;
  p.npos = tiles * MTILE;
  p.D = a.D;
  p.rows = a.rows;
  p.cols = a.cols;
  p.tiles = tiles;
  p.T = pl.T;
  p.prof = a.prof;
  p.debug = a.debug;
  for (int c0 = 0; c0 < a.n; c0 += pl.chains_per_launch) {
    const int chains = a.n - c0 < pl.chains_per_launch ? a.n - c0 : pl.chains_per_launch;
    p.chain0 = c0;
    int rc = 0;
    switch (pl.mt) {
      case 1: rc = launch_mt<1>(p, chains, L.total, stream); break;
      case 2: rc = launch_mt<2>(p, chains, L.total, stream); break;
      case 3: rc = launch_mt<3>(p, chains, L.total, stream); break;
      case 4: rc = launch_mt<4>(p, chains, L.total, stream); break;
      default: rc = launch_mt<5>(p, chains, L.total, stream); break;
    }
    if (rc != 0) return rc;
    p.prof = nullptr;
  }
  // the watchdog's verdict travels to the host behind the sweep; the caller looks at it after its next synchronise
  if (s.abort_host != nullptr)
    B200MVS_CUDA_OK(cudaMemcpyAsync(s.abort_host, p.abort_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  return 0;
}

}  // namespace b200mvs
