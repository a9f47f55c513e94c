// 3x3 (dilated) 32->32 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Implicit GEMM without an im2col copy.  One CTA owns a TH x TW tile of output pixels of one image:
//   * the halo-extended input tile ((TH+2d) rows of PW = 64 positions, TW = PW - 2d of them valid
//     outputs) is staged ONCE in shared memory, previous layer's GroupNorm + LeakyReLU (+ residual)
//     applied on the way, converted to fp16 and laid out as four planes of [position][8 channels]
//     (16 bytes per position).  With the UMMA "no swizzle, K-major" canonical layout (core matrix =
//     8 rows x 16 bytes, contiguous) a plane IS a valid A operand whose row m is position m, and
//     the A operand of filter tap (ky, kx) is the same plane started (ky*d*PW + kx*d) positions
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

#include <vector>

#include "conv.cuh"
#include "conv_tc.cuh"

namespace b200mvs {
namespace {

constexpr int PW = 64;        // positions per tile row (valid outputs: PW - 2*dil)
constexpr int NT = 256;       // threads per CTA
constexpr int W_BYTES = 9 * 2 * 1024;   // fp16 weights: [tap][kstep][half(2)][n(32)][8]

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start address, leading (K) byte offset,
// stride (M/N, 8-row group) byte offset, all in 16-byte units; version = 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = F16, K-major both, N = 32, M = 128.
constexpr uint32_t kIdescF16 = (1u << 4) | (0u << 7) | (0u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__host__ __device__ inline int tc_npos(int TH, int dil) {
  int n = (TH + 2 * dil) * PW + 2 * dil;
  return (n + 7) & ~7;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int TH>
__global__ void __launch_bounds__(NT) conv3x3_tc_kernel(const ConvParams p, const uint8_t* __restrict__ w16) {
  constexpr int MT = TH * PW / 128;   // M-tiles (two output rows each)
  constexpr int TMEM_COLS = MT * 32;  // power of two >= 32 for TH in {4, 8, 16}
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_a[kC], s_b[kC], s_bias[kC];
  __shared__ double s_stats[2 * kGroups];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = blockIdx.y;
  const int d = p.dil;
  const int TW = PW - 2 * d;
  const int tiles_x = cdiv(p.Wo, TW);
  const int tx0 = (blockIdx.x % tiles_x) * TW;
  const int ty0 = (blockIdx.x / tiles_x) * TH;
  const int npos = tc_npos(TH, d);
  const uint32_t plane_bytes = (uint32_t)npos * 16u;
  uint8_t* s_w = smem;              // W_BYTES
  uint8_t* s_in = smem + W_BYTES;   // 4 planes x npos x 16 B

  // ---- one-time setup ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kC) {
    s_bias[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
    if (p.feat.mode >= FEAT_GN) {
      const int grp = tid >> 3;
      const double sum = p.feat.stats[(img * kGroups + grp) * 2 + 0];
      const double sq = p.feat.stats[(img * kGroups + grp) * 2 + 1];
      const double mean = sum * p.feat.inv_count;
      double var = sq * p.feat.inv_count - mean * mean;
      var = var > 0.0 ? var : 0.0;
      const double rstd = rsqrt(var + (double)kGnEps);
      s_a[tid] = (float)((double)p.feat.gamma[tid] * rstd);
      s_b[tid] = (float)((double)p.feat.beta[tid] - mean * (double)p.feat.gamma[tid] * rstd);
    }
  }
  if (tid < 2 * kGroups) s_stats[tid] = 0.0;
  __syncthreads();

  // ---- stage weights (already in the canonical fp16 layout) and the transformed input tile ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(w16);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < W_BYTES / 16; i += NT) dst[i] = __ldg(src + i);
  }
  {
    const size_t vol = (size_t)p.Hi * p.Wi;
    const float* fbase = p.feat.ptr + (size_t)(img / p.feat.img_div) * vol * kC;
    const float* rbase = p.feat.resid != nullptr ? p.feat.resid + (size_t)img * vol * kC : nullptr;
    float* xbase = p.feat.x_out != nullptr ? p.feat.x_out + (size_t)img * vol * kC : nullptr;
    const int rows_in = TH + 2 * d;
    const int mode = p.feat.mode;
    constexpr int BATCH = 4;   // tasks whose global loads are all issued before any is consumed
    for (int i0 = tid; i0 < npos * 4; i0 += NT * BATCH) {
      float4 ya[BATCH], yb[BATCH], ra[BATCH], rb[BATCH];
      size_t off[BATCH];
      bool inb[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * NT;
        const int c8 = i & 3;
        const int L = i >> 2;
        const int iy = L / PW, ix = L % PW;
        const int gy = ty0 - d + iy, gx = tx0 - d + ix;
        inb[k] = i < npos * 4 && iy < rows_in && gy >= 0 && gy < p.Hi && gx >= 0 && gx < p.Wi;
        off[k] = inb[k] ? ((size_t)gy * p.Wi + gx) * kC + 8 * c8 : 0;
        if (inb[k]) {
          ya[k] = __ldg(reinterpret_cast<const float4*>(fbase + off[k]));
          yb[k] = __ldg(reinterpret_cast<const float4*>(fbase + off[k] + 4));
          if (mode == FEAT_GN_RES) {
            ra[k] = __ldg(reinterpret_cast<const float4*>(rbase + off[k]));
            rb[k] = __ldg(reinterpret_cast<const float4*>(rbase + off[k] + 4));
          }
        }
      }
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
          v[0] = ya[k].x; v[1] = ya[k].y; v[2] = ya[k].z; v[3] = ya[k].w;
          v[4] = yb[k].x; v[5] = yb[k].y; v[6] = yb[k].z; v[7] = yb[k].w;
          if (mode >= FEAT_GN) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(v[e], s_a[8 * c8 + e], s_b[8 * c8 + e]));
            if (mode == FEAT_GN_RES) {
              v[0] += ra[k].x; v[1] += ra[k].y; v[2] += ra[k].z; v[3] += ra[k].w;
              v[4] += rb[k].x; v[5] += rb[k].y; v[6] += rb[k].z; v[7] += rb[k].w;
            }
            const int iy = L / PW, ix = L % PW;
            if (xbase != nullptr && iy >= d && iy < d + TH && ix >= d && ix < d + TW) {
              *reinterpret_cast<float4*>(xbase + off[k]) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(xbase + off[k] + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
        }
        uint4 h;
        h.x = pack_half2(v[0], v[1]);
        h.y = pack_half2(v[2], v[3]);
        h.z = pack_half2(v[4], v[5]);
        h.w = pack_half2(v[6], v[7]);
        *reinterpret_cast<uint4*>(s_in + (size_t)c8 * plane_bytes + (size_t)L * 16) = h;
      }
    }
  }
  // generic-proxy writes -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;

  // ---- one thread issues every MMA of the tile, then commits to the mbarrier ----
  if (tid == 0) {
    const uint32_t a0 = smem_u32(s_in);
    const uint32_t w0 = smem_u32(s_w);
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t dcol = tmem_base + (uint32_t)(mt * 32);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const uint32_t pos = (uint32_t)(mt * 128 + ky * d * PW + kx * d);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t adesc = umma_desc(a0 + (uint32_t)(2 * ks) * plane_bytes + pos * 16u, plane_bytes, 128u);
          const uint64_t bdesc = umma_desc(w0 + (uint32_t)((tap * 2 + ks) * 1024), 512u, 128u);
          mma_f16(dcol, adesc, bdesc, (tap | ks) != 0 ? 1u : 0u);
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar))
                 : "memory");
  }
  // ---- everyone waits for the accumulators ----
  {
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(bar), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM -> registers -> bias, statistics, channels-last store ----
  const int wq = warp & 3;  // TMEM lane quarter this warp may read
  float gsum[kGroups], gsq[kGroups];
#pragma unroll
  for (int g = 0; g < kGroups; ++g) gsum[g] = gsq[g] = 0.f;
  const size_t ovol = (size_t)p.Ho * p.Wo;
  const size_t ostride = p.out_img_stride != 0 ? (size_t)p.out_img_stride : ovol * kC;
  for (int mt = warp >> 2; mt < MT; mt += NT / 128) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(mt * 32), v);
    const int j = mt * 128 + wq * 32 + lane;
    const int oy = ty0 + j / PW, ox_t = j % PW;
    const int ox = tx0 + ox_t;
    if (ox_t < TW && ox < p.Wo && oy < p.Ho) {
      const size_t opix = (size_t)oy * p.Wo + ox;
      float* o = p.out + (size_t)img * ostride + opix * kC;
      const float* add = p.add_src != nullptr ? p.add_src + ((size_t)img * ovol + opix) * kC : nullptr;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 r;
        r.x = v[4 * q + 0] + s_bias[4 * q + 0];
        r.y = v[4 * q + 1] + s_bias[4 * q + 1];
        r.z = v[4 * q + 2] + s_bias[4 * q + 2];
        r.w = v[4 * q + 3] + s_bias[4 * q + 3];
        if (add != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(add + 4 * q));
          r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
        }
        *reinterpret_cast<float4*>(o + 4 * q) = r;
        gsum[q >> 1] += (r.x + r.y) + (r.z + r.w);
        gsq[q >> 1] += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
      }
    }
  }
  if (p.out_stats != nullptr) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gsum[g] += __shfl_xor_sync(0xffffffffu, gsum[g], o);
        gsq[g] += __shfl_xor_sync(0xffffffffu, gsq[g], o);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        atomicAdd(&s_stats[2 * g + 0], (double)gsum[g]);
        atomicAdd(&s_stats[2 * g + 1], (double)gsq[g]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.out_stats != nullptr && tid < 2 * kGroups)
    atomicAdd(p.out_stats + (size_t)img * 2 * kGroups + tid, s_stats[tid]);
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

template <int TH>
int launch_th(const ConvParams& p, const uint8_t* w16, cudaStream_t stream) {
  const size_t smem = W_BYTES + (size_t)4 * tc_npos(TH, p.dil) * 16;
  static bool attr_set = false;
  if (!attr_set) {
    B200MVS_CUDA_OK(cudaFuncSetAttribute(conv3x3_tc_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024));
    attr_set = true;
  }
  const int TW = PW - 2 * p.dil;
  dim3 grid(cdiv(p.Wo, TW) * cdiv(p.Ho, TH), p.n_img);
  if (p.tag != TAG_NONE) probe_before(p.tag, stream);
  conv3x3_tc_kernel<TH><<<grid, NT, smem, stream>>>(p, w16);
  if (p.tag != TAG_NONE) probe_after(p.tag, stream);
  B200MVS_LAUNCH_OK("conv3x3_tc_kernel");
  return 0;
}

}  // namespace

void pack_conv3x3_tc_weights(const float* w_oihw, std::vector<uint8_t>* out) {
  out->assign(W_BYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  for (int tap = 0; tap < 9; ++tap)
    for (int c = 0; c < 32; ++c)
      for (int n = 0; n < 32; ++n) {
        const int ks = c / 16, k = c % 16;
        const size_t byte = (size_t)(tap * 2 + ks) * 1024 + (size_t)(k / 8) * 512 + (size_t)n * 16 + (size_t)(k % 8) * 2;
        h[byte / 2] = __float2half_rn(w_oihw[((size_t)n * 32 + c) * 9 + tap]);
      }
}

bool conv3x3_tc_supported(const ConvParams& p) {
  return p.feat.mode != FEAT_NONE && p.extra.n == 0 && p.Di == 1 && p.Do == 1 && p.Hi == p.Ho && p.Wi == p.Wo &&
         p.dil >= 1 && p.dil <= 8;
}

int launch_conv3x3_tc(const ConvParams& p, const uint8_t* w16, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  if (!conv3x3_tc_supported(p)) {
    set_error("launch_conv3x3_tc: unsupported configuration");
    return -1;
  }
  // Tile height: two CTAs per SM where the halo allows it.
  if (p.dil <= 2) return launch_th<16>(p, w16, stream);
  if (p.dil <= 4) return launch_th<8>(p, w16, stream);
  return launch_th<16>(p, w16, stream);
}

}  // namespace b200mvs
