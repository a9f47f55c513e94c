// The depth-sweep feature recurrence (IncrementalFastGeometryAwareFeatureNetwork.forward loop,
// multi_view_stereonet.py:279-290) as ONE persistent kernel: a thread-block cluster per
// (image group, view) runs all D-1 dependent steps with the working set in shared memory.
//
//   step i:  wf   = warp(features_{i-1}, H_{i-1}^-1 H_i)            (bilinear gather, zero outside)
//            img  = warp(right image at 1/16 scale, H_i)
//            y0   = conv3x3(cat[img, wf]) + b ; x0 = lrelu(GN(y0))
//            y1   = conv3x3(x0) + b           ; x1 = lrelu(GN(y1)) + x0
//            features_i = wf + conv3x3(x1) + b                        (FeatureRefiner, :424-440)
//
// Decomposition.  The 1/16-scale image is linearised with a zero column on each side (pitch
// PW = w + 2); CTA r of the cluster owns output positions [128 r, 128 r + 128) = one UMMA M-tile.
// Each 3x3 conv is 9 taps x K/16 k-steps of tcgen05.mma.kind::f16 (M=128, N=32) whose A operand
// for tap (ky, kx) is the staged activation plane started ky*PW + kx positions later (see
// conv_tc.cu).  The 1/16-scale stages need ~fp32 operand accuracy, so every
// operand is split x = hi + lo into two fp16 values and each k-step issues three MMAs
// (hi*hi + lo*hi + hi*lo), fp32 accumulation in TMEM: ~22 significant bits.
// GroupNorm needs whole-image statistics twice per step: CTAs push their partial sums into every
// CTA's shared memory (DSMEM) and meet at a hardware cluster barrier; the raw conv outputs of the
// PW+1 positions next to a tile boundary are pushed into the neighbour CTA's halo buffer at the same
// time, so each CTA normalises its own tile plus halo locally.  Three cluster barriers per step.
// All weights (hi and lo, three layers, 126 KB) stay resident in shared memory for all steps.
#include <cuda_fp16.h>

#include <vector>

#include "conv.cuh"
#include "recurrence.cuh"

namespace b200mvs {
namespace {

constexpr int NT = 256;
constexpr int MTILE = 128;
constexpr int W0_BLOCKS = 9 * 3;   // conv0: taps x k-steps (32 feature + 3 image channels, padded to 48)
constexpr int W1_BLOCKS = 9 * 2;
constexpr int W_TOTAL_BYTES = (W0_BLOCKS + 2 * W1_BLOCKS) * 2 * 1024;  // hi + lo

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr uint32_t kIdescF16 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(kIdescF16), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t raddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cluster_f1(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

// x = hi + lo with both halves fp16 (round to nearest): ~22 significant bits.
__device__ __forceinline__ void split8(const float* v, uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * k] - back.x, v[2 * k + 1] - back.y);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
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

struct Layout {
  int PW, halo, npl, npl_pad;
  uint32_t plane_bytes;
  // byte offsets into dynamic shared memory
  uint32_t off_w, off_planes, off_own, off_halo, off_wf, total;
};

// Planes (each npl_pad x 16 B): [hi f0..f3][hi extra][zero][lo f0..f3][lo extra][zero]
constexpr int PLANE_HI = 0, PLANE_HI_X = 4, PLANE_LO = 6, PLANE_LO_X = 10, NUM_PLANES = 12;

__host__ __device__ inline Layout make_layout(int cols) {
  Layout L;
  L.PW = cols + 2;
  L.halo = L.PW + 1;
  L.npl = MTILE + 2 * L.halo;
  L.npl_pad = (L.npl + 7) & ~7;
  L.plane_bytes = (uint32_t)L.npl_pad * 16u;
  uint32_t o = 0;
  L.off_w = o;
  o += W_TOTAL_BYTES;
  L.off_planes = o;
  o += NUM_PLANES * L.plane_bytes;
  L.off_own = o;           // raw conv output of the own 128 positions, fp32 [128][32]
  o += MTILE * kC * 4;
  L.off_halo = o;          // [layer 2][side 2][halo][32] fp32, written by the neighbour CTAs
  o += 2 * 2 * (uint32_t)L.halo * kC * 4;
  L.off_wf = o;            // warped features of the own positions, fp32 [128][32]
  o += MTILE * kC * 4;
  L.total = o;
  return L;
}

struct RecParams {
  const float* vol_in;   // feature volume [n][D][rows*cols][32]; hypothesis 0 filled
  float* vol;            // same buffer (written for hypotheses 1..D-1)
  GeomOut geo;
  ViewPtrs right_l4;
  const uint8_t* w16;    // packed weights (pack_recurrence_weights)
  const float* bias0;    // [32] x3
  const float* bias1;
  const float* bias2;
  const float* gamma0;
  const float* beta0;
  const float* gamma1;
  const float* beta1;
  int D, rows, cols, n_tiles;
};

__global__ void __launch_bounds__(NT, 1) recurrence_kernel(const RecParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_part[2][16][2 * kGroups];  // per-layer partial (sum, sumsq) of every CTA of the cluster
  __shared__ float s_red[NT / 32][2 * kGroups];
  __shared__ float s_a[kC], s_b[kC], s_bias[3][kC];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t csize = gridDim.x;  // cluster == all CTAs of blockIdx.y
  const int n = blockIdx.y;
  const Layout L = make_layout(p.cols);
  const int PW = L.PW, halo = L.halo, npl = L.npl;
  const int pixels = p.rows * p.cols;
  const bool active = (int)rank < p.n_tiles;
  const int pos0 = (int)rank * MTILE;  // first own output position; also the input position of local l = 0

  uint8_t* s_w = smem + L.off_w;
  uint8_t* s_planes = smem + L.off_planes;
  float* s_own = reinterpret_cast<float*>(smem + L.off_own);
  float* s_halo = reinterpret_cast<float*>(smem + L.off_halo);
  float* s_wf = reinterpret_cast<float*>(smem + L.off_wf);

  // ---- one-time setup ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kC) {
    s_bias[0][tid] = __ldg(p.bias0 + tid);
    s_bias[1][tid] = __ldg(p.bias1 + tid);
    s_bias[2][tid] = __ldg(p.bias2 + tid);
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w16);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < W_TOTAL_BYTES / 16; i += NT) dst[i] = __ldg(src + i);
    // zero every plane once: the two zero planes and the padding lanes of the extra planes stay zero
    uint4* pl = reinterpret_cast<uint4*>(s_planes);
    for (int i = tid; i < NUM_PLANES * L.npl_pad; i += NT) pl[i] = make_uint4(0, 0, 0, 0);
    float4* hz = reinterpret_cast<float4*>(s_halo);
    for (int i = tid; i < 2 * 2 * halo * kC / 4; i += NT) hz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  uint32_t bar_phase = 0;
  cluster_sync_all();  // every CTA's shared memory is initialised before anyone pushes into it

  const uint32_t a_base = smem_u32(s_planes);
  const uint32_t w_base = smem_u32(s_w);
  const float inv_count = 1.0f / (8.0f * (float)pixels);

  // Issues one conv's MMAs: ksteps k-steps per tap, three split terms per k-step.
  auto issue_conv = [&](int w_block0, int ksteps) {
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t acc = 0;
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const uint32_t pos = (uint32_t)((tap / 3) * PW + (tap % 3));
        for (int ks = 0; ks < ksteps; ++ks) {
          // k-steps 0,1 read feature planes (2ks, 2ks+1); k-step 2 reads (extra, zero)
          const int pl = (ks < 2) ? 2 * ks : PLANE_HI_X;
          const uint32_t a_hi = a_base + (uint32_t)(PLANE_HI + pl) * L.plane_bytes + pos * 16u;
          const uint32_t a_lo = a_base + (uint32_t)(PLANE_LO + pl) * L.plane_bytes + pos * 16u;
          const uint32_t b_hi = w_base + (uint32_t)(w_block0 + tap * ksteps + ks) * 2048u;
          const uint32_t b_lo = b_hi + 1024u;
          const uint64_t da_hi = umma_desc(a_hi, L.plane_bytes, 128u), da_lo = umma_desc(a_lo, L.plane_bytes, 128u);
          const uint64_t db_hi = umma_desc(b_hi, 512u, 128u), db_lo = umma_desc(b_lo, 512u, 128u);
          mma_f16(tmem_base, da_hi, db_hi, acc);
          acc = 1;
          mma_f16(tmem_base, da_lo, db_hi, 1u);
          mma_f16(tmem_base, da_hi, db_lo, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&s_bar))
                   : "memory");
    }
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(bar), "r"(bar_phase)
          : "memory");
    }
    bar_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  // This thread's slice of the accumulator tile: one output position, 16 channels.
  const int wq = warp & 3, chalf = warp >> 2;
  const int jl = wq * 32 + lane;                 // local output position
  const int jg = pos0 + jl;                      // global output position
  const int oy = jg / PW, ox = jg % PW;
  const bool real_out = active && ox < p.cols && oy < p.rows;
  const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(chalf * 16);

  // Epilogue of conv0 / conv1: raw output (+bias) -> own buffer, neighbours' halo buffers, statistics.
  auto epilogue_raw = [&](int layer) {
    float v[16];
    float gs[2] = {0.f, 0.f}, gq[2] = {0.f, 0.f};
    if (active) {
      tmem_ld16(tmem_my, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] += s_bias[layer][chalf * 16 + k];
      float4* own = reinterpret_cast<float4*>(s_own + jl * kC + chalf * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) own[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      // halo pushes: the neighbour indexes its halo buffer [layer][side][i][32]
      if (jl < halo && rank > 0) {  // upper halo (side 1) of rank-1, index jl
        const uint32_t la = smem_u32(s_halo + (((layer * 2 + 1) * halo) + jl) * kC + chalf * 16);
        const uint32_t ra = map_to_rank(la, rank - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) st_cluster_f4(ra + 16u * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
      if (jl >= MTILE - halo && (int)rank + 1 < p.n_tiles) {  // lower halo (side 0) of rank+1
        const int idx = jl - (MTILE - halo);
        const uint32_t la = smem_u32(s_halo + (((layer * 2 + 0) * halo) + idx) * kC + chalf * 16);
        const uint32_t ra = map_to_rank(la, rank + 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) st_cluster_f4(ra + 16u * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
      }
      if (real_out) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gs[0] += v[k];
          gq[0] += v[k] * v[k];
          gs[1] += v[8 + k];
          gq[1] += v[8 + k] * v[8 + k];
        }
      }
    }
    // CTA partial: warp shuffle, then 8 warps -> one value per (group, moment)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], o);
      gq[0] += __shfl_xor_sync(0xffffffffu, gq[0], o);
      gs[1] += __shfl_xor_sync(0xffffffffu, gs[1], o);
      gq[1] += __shfl_xor_sync(0xffffffffu, gq[1], o);
    }
    if (lane == 0) {
      // this warp covers groups 2*chalf and 2*chalf + 1
      float* r = s_red[warp];
#pragma unroll
      for (int k = 0; k < 2 * kGroups; ++k) r[k] = 0.f;
      r[(2 * chalf) * 2 + 0] = gs[0];
      r[(2 * chalf) * 2 + 1] = gq[0];
      r[(2 * chalf + 1) * 2 + 0] = gs[1];
      r[(2 * chalf + 1) * 2 + 1] = gq[1];
    }
    __syncthreads();
    if (tid < (int)csize * 2 * kGroups) {
      const int dst = tid / (2 * kGroups), k = tid % (2 * kGroups);
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tot += s_red[w][k];
      const uint32_t la = smem_u32(&s_part[layer][rank][k]);
      st_cluster_f1(map_to_rank(la, (uint32_t)dst), tot);
    }
  };

  // GroupNorm coefficients from the cluster-wide partials (fixed summation order: deterministic).
  auto gn_coeffs = [&](int layer, const float* gamma, const float* beta) {
    if (tid < kC) {
      const int g = tid >> 3;
      double sum = 0.0, sq = 0.0;
      for (int r = 0; r < p.n_tiles; ++r) {
        sum += (double)s_part[layer][r][2 * g];
        sq += (double)s_part[layer][r][2 * g + 1];
      }
      const double mean = sum * (double)inv_count;
      double var = sq * (double)inv_count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const double rstd = rsqrt(var + (double)kGnEps);
      s_a[tid] = (float)((double)__ldg(gamma + tid) * rstd);
      s_b[tid] = (float)((double)__ldg(beta + tid) - mean * (double)__ldg(gamma + tid) * rstd);
    }
    __syncthreads();
  };

  // Raw conv output of local input position l (own tile or a neighbour's pushed halo).
  auto raw_ptr = [&](int layer, int l) -> const float* {
    if (l < halo) return s_halo + ((layer * 2 + 0) * halo + l) * kC;
    if (l < halo + MTILE) return s_own + (l - halo) * kC;
    return s_halo + ((layer * 2 + 1) * halo + (l - halo - MTILE)) * kC;
  };
  auto is_real = [&](int l, int* gy, int* gx) -> bool {
    const int Lg = pos0 + l;
    *gy = Lg / PW - 1;
    *gx = Lg % PW - 1;
    return *gy >= 0 && *gy < p.rows && *gx >= 0 && *gx < p.cols;
  };

  for (int step = 1; step < p.D; ++step) {
    // ================= W: warp previous features and the 1/16 image into the conv0 operand =========
    if (active) {
      const float* prev = p.vol_in + ((size_t)n * p.D + (step - 1)) * pixels * kC;
      const float* Hinc = p.geo.Hinc + ((size_t)n * p.D + step) * 9;
      const float* Hd = p.geo.H + ((size_t)n * p.D + step) * 9;
      for (int i = tid; i < npl * 5; i += NT) {
        const int oct = i % 5, l = i / 5;
        int gy, gx;
        const bool real = is_real(l, &gy, &gx);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
        if (oct < 4) {
          if (real) {
            const WarpCoord c = homography_coord(Hinc, (float)gx, (float)gy, p.rows, p.cols);
            if (!c.invalid) {
              const Bilinear bl = bilinear_setup(c, p.rows, p.cols);
              const float* b00 = prev + ((size_t)bl.y0 * p.cols + bl.x0) * kC + 8 * oct;
              const float* b01 = prev + ((size_t)bl.y0 * p.cols + bl.x1) * kC + 8 * oct;
              const float* b10 = prev + ((size_t)bl.y1 * p.cols + bl.x0) * kC + 8 * oct;
              const float* b11 = prev + ((size_t)bl.y1 * p.cols + bl.x1) * kC + 8 * oct;
#pragma unroll
              for (int hq = 0; hq < 2; ++hq) {
                const float4 a = __ldcg(reinterpret_cast<const float4*>(b00) + hq);
                const float4 b = __ldcg(reinterpret_cast<const float4*>(b01) + hq);
                const float4 cc = __ldcg(reinterpret_cast<const float4*>(b10) + hq);
                const float4 d = __ldcg(reinterpret_cast<const float4*>(b11) + hq);
                v[4 * hq + 0] = a.x * bl.w00 + b.x * bl.w01 + cc.x * bl.w10 + d.x * bl.w11;
                v[4 * hq + 1] = a.y * bl.w00 + b.y * bl.w01 + cc.y * bl.w10 + d.y * bl.w11;
                v[4 * hq + 2] = a.z * bl.w00 + b.z * bl.w01 + cc.z * bl.w10 + d.z * bl.w11;
                v[4 * hq + 3] = a.w * bl.w00 + b.w * bl.w01 + cc.w * bl.w10 + d.w * bl.w11;
              }
            }
          }
          if (l >= halo && l < halo + MTILE) {
            float4* wfp = reinterpret_cast<float4*>(s_wf + (l - halo) * kC + 8 * oct);
            wfp[0] = make_float4(v[0], v[1], v[2], v[3]);
            wfp[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
          uint4 hi, lo;
          split8(v, &hi, &lo);
          *reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_HI + oct) * L.plane_bytes + (size_t)l * 16) = hi;
          *reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_LO + oct) * L.plane_bytes + (size_t)l * 16) = lo;
        } else {
          if (real) {
            const WarpCoord c = homography_coord(Hd, (float)gx, (float)gy, p.rows, p.cols);
            if (!c.invalid) {
              const Bilinear bl = bilinear_setup(c, p.rows, p.cols);
              const float* img = p.right_l4.p[n % p.right_l4.views] + (size_t)(n / p.right_l4.views) * 3 * pixels;
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                const float* pl = img + (size_t)ch * pixels;
                v[ch] = __ldg(pl + bl.y0 * p.cols + bl.x0) * bl.w00 + __ldg(pl + bl.y0 * p.cols + bl.x1) * bl.w01 +
                        __ldg(pl + bl.y1 * p.cols + bl.x0) * bl.w10 + __ldg(pl + bl.y1 * p.cols + bl.x1) * bl.w11;
              }
            }
          }
          uint4 hi, lo;
          split8(v, &hi, &lo);
          *reinterpret_cast<uint4*>(s_planes + (size_t)PLANE_HI_X * L.plane_bytes + (size_t)l * 16) = hi;
          *reinterpret_cast<uint4*>(s_planes + (size_t)PLANE_LO_X * L.plane_bytes + (size_t)l * 16) = lo;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (active) issue_conv(0, 3);
    epilogue_raw(0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // A: statistics and halos of y0 are everywhere

    // ================= S1: x0 = lrelu(GN(y0)) over own + halo -> conv1 operand ====================
    gn_coeffs(0, p.gamma0, p.beta0);
    if (active) {
      for (int i = tid; i < npl * 4; i += NT) {
        const int oct = i & 3, l = i >> 2;
        int gy, gx;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
        if (is_real(l, &gy, &gx)) {
          const float4* src = reinterpret_cast<const float4*>(raw_ptr(0, l) + 8 * oct);
          const float4 a = src[0], b = src[1];
          const float y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = lrelu(fmaf(y[k], s_a[8 * oct + k], s_b[8 * oct + k]));
        }
        uint4 hi, lo;
        split8(v, &hi, &lo);
        *reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_HI + oct) * L.plane_bytes + (size_t)l * 16) = hi;
        *reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_LO + oct) * L.plane_bytes + (size_t)l * 16) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (active) issue_conv(W0_BLOCKS, 2);
    epilogue_raw(1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // C: statistics and halos of y1 are everywhere

    // ================= S2: x1 = lrelu(GN(y1)) + x0 over own + halo -> conv_final operand ==========
    gn_coeffs(1, p.gamma1, p.beta1);
    if (active) {
      for (int i = tid; i < npl * 4; i += NT) {
        const int oct = i & 3, l = i >> 2;
        int gy, gx;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
        uint4* ph = reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_HI + oct) * L.plane_bytes + (size_t)l * 16);
        uint4* plo = reinterpret_cast<uint4*>(s_planes + (size_t)(PLANE_LO + oct) * L.plane_bytes + (size_t)l * 16);
        if (is_real(l, &gy, &gx)) {
          float x0[8];
          unsplit8(*ph, *plo, x0);
          const float4* src = reinterpret_cast<const float4*>(raw_ptr(1, l) + 8 * oct);
          const float4 a = src[0], b = src[1];
          const float y[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = lrelu(fmaf(y[k], s_a[8 * oct + k], s_b[8 * oct + k])) + x0[k];
        }
        uint4 hi, lo;
        split8(v, &hi, &lo);
        *ph = hi;
        *plo = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (active) issue_conv(W0_BLOCKS + W1_BLOCKS, 2);

    // ================= E2: features_step = wf + delta -> global =====================================
    if (active) {
      float v[16];
      tmem_ld16(tmem_my, v);
      if (real_out) {
        float* dst = p.vol + (((size_t)n * p.D + step) * pixels + (size_t)oy * p.cols + ox) * kC + chalf * 16;
        const float* wfp = s_wf + jl * kC + chalf * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 r;
          r.x = wfp[4 * q + 0] + (v[4 * q + 0] + s_bias[2][chalf * 16 + 4 * q + 0]);
          r.y = wfp[4 * q + 1] + (v[4 * q + 1] + s_bias[2][chalf * 16 + 4 * q + 1]);
          r.z = wfp[4 * q + 2] + (v[4 * q + 2] + s_bias[2][chalf * 16 + 4 * q + 2]);
          r.w = wfp[4 * q + 3] + (v[4 * q + 3] + s_bias[2][chalf * 16 + 4 * q + 3]);
          __stcg(reinterpret_cast<float4*>(dst) + q, r);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __threadfence();
    cluster_sync_all();  // E: hypothesis `step` is visible to every CTA's gathers
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
  }
}

}  // namespace

// Weight blocks of 2 KB: [hi 1 KB][lo 1 KB], each [k half (2)][n (32)][8 fp16] (UMMA K-major, no swizzle).
// Order: conv0 [tap][kstep 3] (k-step 2 = image channels 0..2 then zeros), conv1 [tap][kstep 2], conv2.
void pack_recurrence_weights(const float* w0_oihw35, const float* w1_oihw32, const float* w2_oihw32,
                             std::vector<uint8_t>* out) {
  out->assign(W_TOTAL_BYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  auto put = [&](int block, int k, int nn, float w) {
    const size_t e = (size_t)block * 1024 + (size_t)(k / 8) * 256 + (size_t)nn * 8 + (size_t)(k % 8);  // in halves
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    h[e] = hi;
    h[e + 512] = lo;
  };
  for (int tap = 0; tap < 9; ++tap)
    for (int nn = 0; nn < 32; ++nn) {
      // reference conv0 input order: [image 0..2, features 0..31]  (multi_view_stereonet.py:425)
      for (int c = 0; c < 32; ++c) put(tap * 3 + c / 16, c % 16, nn, w0_oihw35[((size_t)nn * 35 + 3 + c) * 9 + tap]);
      for (int c = 0; c < 3; ++c) put(tap * 3 + 2, c, nn, w0_oihw35[((size_t)nn * 35 + c) * 9 + tap]);
      for (int c = 0; c < 32; ++c) {
        put(W0_BLOCKS + tap * 2 + c / 16, c % 16, nn, w1_oihw32[((size_t)nn * 32 + c) * 9 + tap]);
        put(W0_BLOCKS + W1_BLOCKS + tap * 2 + c / 16, c % 16, nn, w2_oihw32[((size_t)nn * 32 + c) * 9 + tap]);
      }
    }
}

bool recurrence_supported(int rows, int cols, int* n_tiles, size_t* smem_bytes) {
  const Layout L = make_layout(cols);
  const int tiles = cdiv(rows * L.PW, MTILE);
  if (n_tiles != nullptr) *n_tiles = tiles;
  if (smem_bytes != nullptr) *smem_bytes = L.total;
  return tiles <= 16 && L.halo <= MTILE && L.total <= 225 * 1024;
}

int launch_recurrence(const RecurrenceArgs& a, cudaStream_t stream) {
  int n_tiles = 0;
  size_t smem = 0;
  if (!recurrence_supported(a.rows, a.cols, &n_tiles, &smem)) {
    set_error("launch_recurrence: shape not supported by the persistent kernel");
    return -1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B200MVS_CUDA_OK(cudaFuncSetAttribute(recurrence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         225 * 1024));
    B200MVS_CUDA_OK(cudaFuncSetAttribute(recurrence_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  RecParams p;
  p.vol_in = a.vol;
  p.vol = a.vol;
  p.geo = a.geo;
  p.right_l4 = a.right_l4;
  p.w16 = a.w16;
  p.bias0 = a.bias0;
  p.bias1 = a.bias1;
  p.bias2 = a.bias2;
  p.gamma0 = a.gamma0;
  p.beta0 = a.beta0;
  p.gamma1 = a.gamma1;
  p.beta1 = a.beta1;
  p.D = a.D;
  p.rows = a.rows;
  p.cols = a.cols;
  p.n_tiles = n_tiles;

  // Cluster size: one CTA per M-tile; if that size cannot be scheduled, pad with idle CTAs.
  static int good_cluster[17] = {0};
  int candidates[3] = {n_tiles, (n_tiles + 1) & ~1, 16};
  if (good_cluster[n_tiles] != 0) candidates[0] = candidates[1] = candidates[2] = good_cluster[n_tiles];
  cudaError_t e = cudaErrorUnknown;
  for (int c = 0; c < 3; ++c) {
    const int cs = candidates[c];
    if (cs < n_tiles || cs > 16) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, a.n, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (good_cluster[n_tiles] == 0) {
      int max_clusters = 0;
      e = cudaOccupancyMaxActiveClusters(&max_clusters, recurrence_kernel, &cfg);
      if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        e = cudaErrorLaunchOutOfResources;
        continue;
      }
    }
    e = cudaLaunchKernelEx(&cfg, recurrence_kernel, p);
    if (e == cudaSuccess) {
      good_cluster[n_tiles] = cs;
      break;
    }
    cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error(std::string("recurrence_kernel launch (") + std::to_string(n_tiles) +
              " tiles): " + cudaGetErrorString(e));
    return -2;
  }
  note_launch();
  return 0;
}

}  // namespace b200mvs
