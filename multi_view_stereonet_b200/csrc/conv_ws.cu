// Persistent, warp-specialised 3x3 (dilated) 32->32 convolution for the residual blocks of IDepthmapRefiner at the
// large pyramid levels (multi_view_stereonet.py:473-478; utils/resnet.py:93-109): fp16 activations in HBM, fp16
// operands, fp32 accumulation in TMEM.  Same implicit-GEMM formulation as conv_tc.cu (one fp16 plane per channel
// octet, nine shifted-window descriptors), but one CTA per SM walks over tiles with three roles running
// concurrently:
//   producers (8 warps)  y, resid -> x = lrelu(GN(y)) + resid -> x_out (tile interior) and the fp16 planes of the
//                        next shared-memory stage
//   MMA (1 warp)         9 taps x 2 k-steps x M-tiles of tcgen05.mma per tile into one of two TMEM accumulator sets;
//                        tcgen05.commit frees the stage and publishes the accumulators
//   epilogue (4 warps)   TMEM -> registers -> bias, GroupNorm statistics of the raw output, fp16 store
// so the loads of tile i+1, the MMAs of tile i and the stores of tile i-1 overlap; mbarriers carry every hand-off.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "conv.cuh"
#include "conv_ws.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int PW = 64;                 // positions per tile row (valid outputs: PW - 2 * dil)
constexpr int TH = 8;                  // output rows per tile
constexpr int MT = TH * PW / 128;      // M-tiles per tile
constexpr int N_EPI = 4, N_PROD = 8;   // warps; the producers are 64 columns x 4 channel octets
constexpr int NT = (N_EPI + 1 + N_PROD) * 32;
constexpr int NPROD_T = N_PROD * 32;
constexpr int PROD_T0 = (N_EPI + 1) * 32;
constexpr uint32_t W_BYTES = 9u * 2u * 1024u;
constexpr int MAX_STAGES = 3;
constexpr int ACC_COLS = MT * 32;      // TMEM columns of one accumulator set

__host__ __device__ inline int ws_npos(int dil) {
  const int n = (TH + 2 * dil) * PW + 2 * dil;
  return (n + 7) & ~7;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
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

// Optional timeline of CTA 0 (globaltimer ns): [tile < 8][producer stage free, producer filled, MMA operands seen,
// MMA committed, epilogue accumulators seen, epilogue done]; written only when B200MVS_WS_PROFILE is set.
__device__ long long g_ws_prof[8 * 6 + 2 + 32];
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct WsParams {
  const __half* y;        // raw output of the previous conv [n][H][W][32]
  const __half* resid;    // residual stream (FEAT_GN_RES) or null (FEAT_GN)
  __half* x_out;          // transformed input, written once per pixel
  __half* out;            // raw output of this conv
  const double* stats;    // [n][4][2] of y
  double* out_stats;      // [n][4][2] of out
  const float* gamma;
  const float* beta;
  const float* bias;
  double inv_count;
  int n_img, H, W, dil;
  int tiles_x, tiles_y, stages;
  int prof, hints;
  int dbg;   // timing ablations (wrong results): 1 no loads, 2 no x_out stores, 4 no output stores, 8 no MMAs
};

__global__ void __launch_bounds__(NT, 1) conv3x3_ws_kernel(const WsParams p, const uint8_t* __restrict__ w16) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_full[MAX_STAGES], s_empty[MAX_STAGES], s_tfull[2], s_tempty[2];
  __shared__ float s_bias[kC];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int d = p.dil;
  const int TW = PW - 2 * d;
  const int npos = ws_npos(d);
  const uint32_t plane_bytes = (uint32_t)npos * 16u;
  const uint32_t stage_bytes = 4u * plane_bytes;
  const int S = p.stages;
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int total = tiles_img * p.n_img;
  uint8_t* s_w = smem;
  uint8_t* s_st = smem + W_BYTES;

  // ---- prologue: nothing here depends on earlier kernels ----
  if (warp == 0) tc::tmem_alloc(&s_tmem, 2u * ACC_COLS);
  if (tid == 32) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      tc::mbar_init(&s_full[s], NPROD_T);
      tc::mbar_init(&s_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&s_tfull[a], 1);
      tc::mbar_init(&s_tempty[a], N_EPI * 32);
    }
    tc::mbar_init_fence();
  }
  if (tid < kC) s_bias[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
  {
    const uint4* src = reinterpret_cast<const uint4*>(w16);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < (int)(W_BYTES / 16); i += NT) dst[i] = __ldg(src + i);
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_launch_dependents();
  pdl_wait();
  if (p.prof && blockIdx.x == 0 && tid == 0) g_ws_prof[48] = gtime();
  const uint32_t tmem_base = s_tmem;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_dead = l2_policy_evict_first();
  const size_t img_elems = (size_t)p.H * p.W * kC;

  if (warp >= N_EPI + 1) {
    // =========================== producers ===========================
    // Thread = (tile column ix, channel octet c8), walking down the rows of the halo-extended tile in batches of RB
    // rows: addresses advance by one image row per task, nothing else is recomputed.  The global loads of batch
    // g + 1 are issued before batch g is transformed and stored (across tile boundaries too), so two batches per
    // thread are always in flight.  Within a warp lane = c8 * 8 + (ix & 7): a quarter-warp writes 128 contiguous
    // bytes of one plane (no bank conflicts) and the warp reads 8 pixels x 64 B contiguous from global memory.
    const int pw = warp - (N_EPI + 1);
    const int c8 = lane >> 3;
    const int ix = pw * 8 + (lane & 7);
    const int rows_in = TH + 2 * d;
    const bool has_res = p.resid != nullptr;
    constexpr int RB = 4;
    const int nb = cdiv(rows_in, RB);                               // batches per tile
    const int my_tiles = blockIdx.x < total ? (total - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int nwork = my_tiles * nb;
    const bool ptid0 = pw == 0 && lane == 0;
    const size_t row_elems = (size_t)p.W * kC;
    struct Regs {
      uint4 ya[RB], ra[RB];
      unsigned inb;
    };
    float ca[8], cb[8];
    int cur_img = -1;
    // per-tile state of the two pipeline ends (issue runs one batch ahead of consume)
    struct TileState {
      int it, b, img, gy0;
      bool col_ok, col_int;
      size_t off0;       // element offset of (row gy0, column gx, octet c8) inside the image
    };
    auto advance = [&](TileState& T) {   // next batch; recompute the tile origin when a new tile starts
      if (++T.b == nb || T.it < 0) {
        T.b = 0;
        ++T.it;
        const int t = blockIdx.x + T.it * gridDim.x;
        T.img = t / tiles_img;
        const int tt = t - T.img * tiles_img;
        const int tyi = tt / p.tiles_x;
        const int tx0 = (tt - tyi * p.tiles_x) * TW;
        T.gy0 = tyi * TH - d;
        const int gx = tx0 - d + ix;
        T.col_ok = gx >= 0 && gx < p.W;
        T.col_int = ix >= d && ix < d + TW;
        T.off0 = (size_t)T.img * img_elems + (size_t)(T.col_ok ? gx : 0) * kC + 8 * c8;
      }
    };
    auto issue = [&](const TileState& T, Regs& R) {
      R.inb = 0;
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        const int r = T.b * RB + k;
        const int gy = T.gy0 + r;
        if (T.col_ok && r < rows_in && gy >= 0 && gy < p.H && !(p.dbg & 1)) {
          R.inb |= 1u << k;
          const size_t off = T.off0 + (size_t)gy * row_elems;
          R.ya[k] = ldg_hint(p.y + off, pol_dead);
          if (has_res) R.ra[k] = ldg_hint(p.resid + off, pol_dead);
        }
      }
    };
    int cons_n = 0;
    auto consume = [&](const TileState& T, const Regs& R) {
      const int s = T.it % S;
      if (p.prof && blockIdx.x == 0 && ptid0 && cons_n < 16) g_ws_prof[50 + 2 * cons_n] = gtime();
      if (T.b == 0) {
        if (T.img != cur_img) {
          cur_img = T.img;
          const double sum = p.stats[(T.img * kGroups + c8) * 2 + 0];   // octet == GroupNorm group
          const double sq = p.stats[(T.img * kGroups + c8) * 2 + 1];
          const double mean = sum * p.inv_count;
          double var = sq * p.inv_count - mean * mean;
          var = var > 0.0 ? var : 0.0;
          const double rstd = rsqrt(var + (double)kGnEps);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const double gm = (double)__ldg(p.gamma + 8 * c8 + e);
            ca[e] = (float)(gm * rstd);
            cb[e] = (float)((double)__ldg(p.beta + 8 * c8 + e) - mean * gm * rstd);
          }
        }
        tc::mbar_wait_warp(&s_empty[s], (uint32_t)(((T.it / S) & 1) ^ 1));   // the MMAs that read this stage completed
        if (p.prof && blockIdx.x == 0 && ptid0 && T.it < 8) g_ws_prof[T.it * 6 + 0] = gtime();
      }
      uint8_t* plane = s_st + (size_t)s * stage_bytes + (size_t)c8 * plane_bytes + (size_t)ix * 16;
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        const int r = T.b * RB + k;
        if (r >= rows_in) continue;
        uint4 h = make_uint4(0, 0, 0, 0);
        if (R.inb & (1u << k)) {
          float v[8];
          unpack8(R.ya[k], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(v[e], ca[e], cb[e]));
          if (has_res) {
            float rr[8];
            unpack8(R.ra[k], rr);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += rr[e];
          }
          h = pack8(v);
          if (p.x_out != nullptr && T.col_int && r >= d && r < d + TH && !(p.dbg & 2))
            stg_hint(p.x_out + T.off0 + (size_t)(T.gy0 + r) * row_elems, h, pol_keep);
        }
        *reinterpret_cast<uint4*>(plane + (size_t)r * (PW * 16)) = h;
      }
      if (T.b == nb - 1) {
        tc::fence_proxy_async();
        mbar_arrive(&s_full[s]);
        if (p.prof && blockIdx.x == 0 && ptid0 && T.it < 8) g_ws_prof[T.it * 6 + 1] = gtime();
      }
      if (p.prof && blockIdx.x == 0 && ptid0 && cons_n < 16) g_ws_prof[51 + 2 * cons_n] = gtime();
      ++cons_n;
    };
    // the tail positions past the last staged row are read by the garbage columns only; keep them zero
    for (int s = 0; s < S; ++s)
      for (int i = tid - PROD_T0; i < 4 * (npos - rows_in * PW); i += NPROD_T) {
        const int pl = i / (npos - rows_in * PW), r = i % (npos - rows_in * PW);
        *reinterpret_cast<uint4*>(s_st + (size_t)s * stage_bytes + (size_t)pl * plane_bytes +
                                  (size_t)(rows_in * PW + r) * 16) = make_uint4(0, 0, 0, 0);
      }
    Regs A, B;
    TileState ti, tc_;
    ti.it = -1; ti.b = 0;
    tc_.it = -1; tc_.b = 0;
    if (nwork > 0) {
      advance(ti);
      issue(ti, A);
    }
    for (int g = 0; g < nwork; g += 2) {
      if (g + 1 < nwork) {
        advance(ti);
        issue(ti, B);
      }
      advance(tc_);
      consume(tc_, A);
      if (g + 2 < nwork) {
        advance(ti);
        issue(ti, A);
      }
      if (g + 1 < nwork) {
        advance(tc_);
        consume(tc_, B);
      }
    }
  } else if (warp == N_EPI) {
    // =========================== MMA issue ===========================
    if (tc::elect_one()) {
      const uint32_t plane_u16 = plane_bytes >> 4;
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_st), plane_bytes, 128u);
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 512u, 128u);
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        const int a = it & 1;
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&s_tempty[a], aph ^ 1u);   // the epilogue has drained this accumulator set
        tc::mbar_wait(&s_full[s], ph);           // the producers have filled this stage
        if (p.prof && blockIdx.x == 0 && it < 8) g_ws_prof[it * 6 + 2] = gtime();
        tc::fence_after_sync();
        const uint64_t da = da0 + (uint64_t)(s * 4) * plane_u16;
#pragma unroll 1
        for (int mt = 0; mt < ((p.dbg & 8) ? 0 : MT); ++mt) {
          const uint32_t dcol = tmem_base + (uint32_t)(a * ACC_COLS + mt * 32);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t pos = (uint32_t)(mt * 128 + (tap / 3) * d * PW + (tap % 3) * d);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_f16(dcol, da + (uint64_t)(2u * ks * plane_u16 + pos), db0 + (uint64_t)((tap * 2 + ks) * 64),
                          tc::idesc_f16(32), (tap | ks) != 0 ? 1u : 0u);
          }
        }
        tc::mma_commit(&s_empty[s]);
        tc::mma_commit(&s_tfull[a]);
        if (p.prof && blockIdx.x == 0 && it < 8) g_ws_prof[it * 6 + 3] = gtime();
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int wq = warp;   // TMEM lane quarter
    float gsum[kGroups], gsq[kGroups];
#pragma unroll
    for (int g = 0; g < kGroups; ++g) gsum[g] = gsq[g] = 0.f;
    int cur_img = -1;
    auto flush = [&](int img) {
      if (p.out_stats == nullptr || img < 0) return;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gsum[g] += __shfl_xor_sync(0xffffffffu, gsum[g], o);
          gsq[g] += __shfl_xor_sync(0xffffffffu, gsq[g], o);
        }
        if (lane == 0) {
          atomicAdd(p.out_stats + ((size_t)img * kGroups + g) * 2 + 0, (double)gsum[g]);
          atomicAdd(p.out_stats + ((size_t)img * kGroups + g) * 2 + 1, (double)gsq[g]);
        }
        gsum[g] = gsq[g] = 0.f;
      }
    };
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      const int img = t / tiles_img, tt = t % tiles_img;
      const int tx0 = (tt % p.tiles_x) * TW, ty0 = (tt / p.tiles_x) * TH;
      if (img != cur_img) {
        flush(cur_img);
        cur_img = img;
      }
      tc::mbar_wait_warp(&s_tfull[a], aph);
      tc::fence_after_sync();
      if (p.prof && blockIdx.x == 0 && tid == 0 && it < 8) g_ws_prof[it * 6 + 4] = gtime();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int j = mt * 128 + wq * 32 + lane;
        const int oy = ty0 + j / PW, ox_t = j % PW, ox = tx0 + ox_t;
        const bool valid = ox_t < TW && ox < p.W && oy < p.H;
        float v[32];
        const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * ACC_COLS + mt * 32);
        tc::tmem_ld16(ta, v);
        tc::tmem_ld16(ta + 16u, v + 16);
        if (valid && !(p.dbg & 4)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            v[c] += s_bias[c];
            gsum[c >> 3] += v[c];
            gsq[c >> 3] = fmaf(v[c], v[c], gsq[c >> 3]);
          }
          __half* o = p.out + (size_t)img * img_elems + ((size_t)oy * p.W + ox) * kC;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (p.hints) stg_hint(o + 8 * q, pack8(v + 8 * q), pol_keep);
            else *reinterpret_cast<uint4*>(o + 8 * q) = pack8(v + 8 * q);
          }
        }
      }
      tc::fence_before_sync();
      mbar_arrive(&s_tempty[a]);
      if (p.prof && blockIdx.x == 0 && tid == 0 && it < 8) g_ws_prof[it * 6 + 5] = gtime();
    }
    flush(cur_img);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (p.prof && blockIdx.x == 0 && tid == 0) g_ws_prof[49] = gtime();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 2u * ACC_COLS);
}

size_t ws_stage_bytes(int dil) { return (size_t)4 * ws_npos(dil) * 16; }

}  // namespace

bool conv3x3_ws_supported(const ConvParams& p) {
  if (!(p.feat.mode == FEAT_GN || p.feat.mode == FEAT_GN_RES) || !p.feat.half_io || !p.out_half || p.extra.n != 0)
    return false;
  if (p.Di != 1 || p.Do != 1 || p.Hi != p.Ho || p.Wi != p.Wo || p.dil < 1 || p.dil > 8) return false;
  if (p.add_src != nullptr || p.out_img_stride != 0 || p.feat.img_div != 1) return false;
  const size_t budget = 220 * 1024 - W_BYTES;
  if (2 * ws_stage_bytes(p.dil) > budget) return false;
  // Worth it only when every SM gets at least a couple of tiles; smaller layers are latency bound either way.
  const long long tiles = (long long)cdiv(p.Wo, PW - 2 * p.dil) * cdiv(p.Ho, TH) * p.n_img;
  return tiles >= 148;
}

int launch_conv3x3_ws(const ConvParams& p, const uint8_t* w16, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  if (!conv3x3_ws_supported(p)) {
    set_error("launch_conv3x3_ws: unsupported configuration");
    return -1;
  }
  WsParams q;
  q.y = reinterpret_cast<const __half*>(p.feat.ptr);
  q.resid = p.feat.mode == FEAT_GN_RES ? reinterpret_cast<const __half*>(p.feat.resid) : nullptr;
  q.x_out = reinterpret_cast<__half*>(p.feat.x_out);
  q.out = reinterpret_cast<__half*>(p.out);
  q.stats = p.feat.stats;
  q.out_stats = p.out_stats;
  q.gamma = p.feat.gamma;
  q.beta = p.feat.beta;
  q.bias = p.bias;
  q.inv_count = p.feat.inv_count;
  q.n_img = p.n_img;
  q.H = p.Hi;
  q.W = p.Wi;
  q.dil = p.dil;
  q.tiles_x = cdiv(p.Wo, PW - 2 * p.dil);
  q.tiles_y = cdiv(p.Ho, TH);
  const size_t budget = 220 * 1024 - W_BYTES;
  int stages = (int)(budget / ws_stage_bytes(p.dil));
  stages = stages > MAX_STAGES ? MAX_STAGES : stages;
  q.stages = stages;
  static const bool prof = getenv("B200MVS_WS_PROFILE") != nullptr;
  q.prof = prof ? 1 : 0;
  static const int hints = getenv("B200MVS_L2_HINTS") ? atoi(getenv("B200MVS_L2_HINTS")) : 1;
  q.hints = hints;
  static const int dbg = getenv("B200MVS_WS_DEBUG") ? atoi(getenv("B200MVS_WS_DEBUG")) : 0;
  q.dbg = dbg;
  const size_t smem = W_BYTES + (size_t)stages * ws_stage_bytes(p.dil);
  static bool attr_set = false;
  if (!attr_set) {
    B200MVS_CUDA_OK(cudaFuncSetAttribute(conv3x3_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    B200MVS_CUDA_OK(cudaGetDevice(&dev));
    B200MVS_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long total = (long long)q.tiles_x * q.tiles_y * q.n_img;
  const int grid = (int)(total < num_sms ? total : num_sms);
  if (p.tag != TAG_NONE) probe_before(p.tag, stream);
  launch_pdl(conv3x3_ws_kernel, dim3(grid), dim3(NT), smem, stream, q, w16);
  if (p.tag != TAG_NONE) probe_after(p.tag, stream);
  B200MVS_LAUNCH_OK("conv3x3_ws_kernel");
  if (prof) {
    long long h[82];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_ws_prof, sizeof(h));
    fprintf(stderr, "ws dil=%d grid=%d tiles=%lld: start 0 end %lld ns\n", p.dil, grid, total, h[49] - h[48]);
    for (int t = 0; t < 8 && t * (long long)grid < total; ++t)
      fprintf(stderr, "  tile %d: free %lld filled %lld | mma seen %lld committed %lld | epi seen %lld done %lld\n", t,
              h[t * 6] - h[48], h[t * 6 + 1] - h[48], h[t * 6 + 2] - h[48], h[t * 6 + 3] - h[48], h[t * 6 + 4] - h[48],
              h[t * 6 + 5] - h[48]);
    fprintf(stderr, "  consume [start,end]:");
    for (int i = 0; i < 16; ++i) fprintf(stderr, " [%lld,%lld]", h[50 + 2 * i] - h[48], h[51 + 2 * i] - h[48]);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace b200mvs
