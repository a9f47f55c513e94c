// Persistent, warp-specialised 3x3 (dilated) 32->32 convolution for the residual blocks of IDepthmapRefiner at the
// large pyramid levels (multi_view_stereonet.py:473-478; utils/resnet.py:93-109): fp16 activations in HBM, fp16
// operands, fp32 accumulation in TMEM.  Same implicit-GEMM formulation as conv_tc.cu (one fp16 plane of
// [position][8 channels] per channel octet, nine shifted-window descriptors), but one CTA per SM walks over tiles
// with four roles running concurrently, every hand-off an mbarrier:
//   TMA (1 lane)            streams the halo-extended tile two rows at a time through a 4-slot ring: per chunk one
//                           cp.async.bulk.tensor box [2 rows][64 positions][32 channels] of the raw previous output y
//                           and one of the residual stream (64-byte inner rows; out-of-image positions arrive as
//                           zeros)
//   transform (16 warps)    ring -> operand planes of the current stage: x = lrelu(GN(y)) + resid (zero outside the
//                           image: the convolution pads x, not y); x_out for the tile interior
//   MMA (1 lane)            9 taps x 2 k-steps x M-tiles of tcgen05.mma per tile into one of two TMEM accumulator sets;
//                           tcgen05.commit frees the stage and publishes the accumulators
//   epilogue (4 warps)      TMEM -> registers -> bias, GroupNorm statistics of the raw output, fp16 store
// so the loads of tile i+1, the transform / MMAs of tile i and the stores of tile i-1 overlap.
//
// Dilated layers (dilation d = 2, 4, 8) run as d vertical polyphase components: a tile is TH rows of the sub-image
// made of image rows py, py + d, py + 2d, ... (TMA element stride d along y), on which the vertical taps are +-1
// row -- a halo of 2 rows instead of 2d (at d = 8 a tile loaded and transformed 24 rows for 8 rows of output, now
// 10).  Horizontally the dilation stays in the descriptors (tap offset kx * d positions, TW = 64 - 2d valid
// columns), so loads and stores remain contiguous along x.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "conv.cuh"
#include "conv_ws.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int PW = 64;                 // positions per tile row (valid outputs: PW - 2 * dil)
constexpr int N_EPI = 8, N_XF = 16;    // warps: epilogue, transform (+ 1 MMA + 1 TMA)
constexpr int NT = (N_EPI + 2 + N_XF) * 32;
constexpr int NXF_T = N_XF * 32;
constexpr uint32_t W_BYTES = 9u * 2u * 1024u;
constexpr int MAX_STAGES = 2;
constexpr int MAX_RING = 6;                    // raw chunks (2 tile rows of y and of resid) in flight, at most
constexpr uint32_t CHUNK_HALF = 2u * PW * 64u;  // one tensor's part of a chunk: 2 rows x 64 positions x 64 B
__host__ __device__ inline uint32_t ring_bytes(int ring) { return (uint32_t)ring * 2u * CHUNK_HALF; }
constexpr size_t kSmemBudget = 220 * 1024;

// positions per plane, padded to 2 (mod 8): consecutive planes then start 32 bytes apart modulo 128, so the four
// octet planes a quarter-warp writes at once fall into distinct banks
// vs = vertical tap distance in tile rows (1 in polyphase mode, the dilation otherwise)
__host__ __device__ inline int ws_npos(int th, int vs, int dil) {
  const int n = (th + 2 * vs) * PW + 2 * dil;
  return ((n + 5) & ~7) + 2;
}
__host__ __device__ inline uint32_t ws_stage_bytes(int th, int vs, int dil) {
  return 4u * (uint32_t)ws_npos(th, vs, dil) * 16u;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          tc::smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(tc::smem_u32(bar))
      : "memory");
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

// Optional timeline of CTA 0 (globaltimer ns, B200MVS_WS_PROFILE=1): per tile < 8:
// [TMA issued, boxes landed, transform done, MMA operands seen, MMA committed, epilogue accumulators seen, epilogue done]
__device__ long long g_ws_prof[8 * 7 + 2];
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WS_STAMP(cond, slot)                                                          \
  do {                                                                                \
    if (p.prof && blockIdx.x == 0 && (cond) && it < 8) g_ws_prof[it * 7 + (slot)] = gtime(); \
  } while (0)

struct WsParams {
  __half* x_out;          // transformed input, written once per pixel (or null)
  __half* out;            // raw output of this conv
  const double* stats;    // [n][4][2] of y
  double* out_stats;      // [n][4][2] of out
  const float* gamma;
  const float* beta;
  const float* bias;
  double inv_count;
  int n_img, H, W, dil;
  int tiles_x, tiles_y, stages, ring, has_res, prof;
  int poly;   // 1: vertical polyphase (tile rows are image rows py + dil * k); tiles_y counts rows of one component
  int dbg;   // timing ablations (wrong results): 1 transform warps only wait / arrive, 2 no MMAs, 4 no output stores
};

template <int TH>
__global__ void __launch_bounds__(NT, 1) conv3x3_ws_kernel(const WsParams p, const uint8_t* __restrict__ w16,
                                                           const __grid_constant__ CUtensorMap tm_y,
                                                           const __grid_constant__ CUtensorMap tm_r) {
  constexpr int MT = TH * PW / 128;      // M-tiles per tile
  constexpr int ACC_COLS = MT * 32;      // TMEM columns of one accumulator set
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  extern __shared__ __align__(128) uint8_t smem[];
  // s_full / s_tfull are per M-tile: an M-tile's MMAs start when the chunks under its three tap rows are transformed
  // (not the whole tile), and its epilogue when ITS MMAs have completed -- the tile's pipeline fill shrinks from
  // (transform + MMA + epilogue) of a tile to roughly that of one M-tile
  constexpr int MAX_MT = 4;
  static_assert(MT <= MAX_MT, "tile height");
  __shared__ __align__(8) uint64_t s_raw[MAX_RING], s_rfree[MAX_RING], s_full[MAX_STAGES][MAX_MT], s_empty[MAX_STAGES],
      s_tfull[2][MAX_MT], s_tempty[2];
  __shared__ __align__(8) uint64_t s_wbar;   // bulk copy of the weights
  __shared__ float s_bias[kC];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const int d = p.dil;
  const int vs = p.poly ? 1 : d;     // vertical tap distance in tile rows
  const int rs = p.poly ? d : 1;     // image rows per tile row
  const int TW = PW - 2 * d;
  const int rows_in = TH + 2 * vs;
  const int npos = ws_npos(TH, vs, d);
  const uint32_t plane_bytes = (uint32_t)npos * 16u;
  const uint32_t stage_bytes = ws_stage_bytes(TH, vs, d);
  const int S = p.stages;
  const int tiles_comp = p.tiles_x * p.tiles_y;            // tiles of one polyphase component (or of the image)
  const int tiles_img = tiles_comp * (p.poly ? d : 1);
  const int total = tiles_img * p.n_img;
  // tile t -> image, first image row of tile row 0 (incl. the halo row(s) above), first column, component
  struct TileAt { int img, tx0, ty0, y_first; };
  auto tile_at = [&](int t) {
    TileAt a;
    a.img = t / tiles_img;
    int tt = t - a.img * tiles_img;
    const int py = p.poly ? tt / tiles_comp : 0;
    tt -= py * tiles_comp;
    const int tyi = tt / p.tiles_x;
    a.tx0 = (tt - tyi * p.tiles_x) * TW;
    a.ty0 = tyi * TH;                                  // first output row, in rows of the component
    a.y_first = py + rs * (a.ty0 - vs);                // image row of tile row 0
    return a;
  };
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + W_BYTES;            // [slot][y | resid][2 rows][64 positions][32 channels]
  const int RING = p.ring;
  uint8_t* s_st = smem + W_BYTES + ring_bytes(RING);

  // ---- prologue: nothing here depends on earlier kernels ----
  if (warp == 0) tc::tmem_alloc(&s_tmem, TMEM_COLS);
  if (tid == 32) {
    for (int q = 0; q < MAX_RING; ++q) {
      tc::mbar_init(&s_raw[q], 1);
      tc::mbar_init(&s_rfree[q], NXF_T);
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      for (int m = 0; m < MAX_MT; ++m) tc::mbar_init(&s_full[s][m], NXF_T);
      tc::mbar_init(&s_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      for (int m = 0; m < MAX_MT; ++m) tc::mbar_init(&s_tfull[a][m], 1);
      tc::mbar_init(&s_tempty[a], N_EPI * 32);
    }
    tc::mbar_init(&s_wbar, 1);
    tc::mbar_init_fence();
    // constant data, fetched by the copy engine while the CTA sets itself up; only the MMA lane waits for it.
    // Two of these CTAs cannot share an SM, so this prologue runs AFTER the previous layer's CTA has left the SM --
    // on the critical path between two layers.
    tc::bulk_load_weights(s_w, w16, W_BYTES, &s_wbar);
  }
  if (tid < kC) s_bias[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
  {
    // the tail positions past the last staged row are read by the garbage columns only; keep them zero
    const int tail = npos - rows_in * PW;
    for (int i = tid; i < S * 4 * tail; i += NT) {
      const int s = i / (4 * tail), pl = (i / tail) % 4, r = i % tail;
      *reinterpret_cast<uint4*>(s_st + (size_t)s * stage_bytes + (size_t)pl * plane_bytes + (size_t)(rows_in * PW + r) * 16) =
          make_uint4(0, 0, 0, 0);
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_launch_dependents();
  pdl_wait();
  if (p.prof && blockIdx.x == 0 && tid == 0) g_ws_prof[56] = gtime();
  const uint32_t tmem_base = s_tmem;
  const uint64_t pol_keep = l2_policy_evict_last();
  const size_t img_elems = (size_t)p.H * p.W * kC;

  if (warp >= N_EPI + 2) {
    // =========================== transform ===========================
    // Per chunk (2 tile rows) thread = (row, tile column ix, channel octet c8) with lane = (ix & 7) * 4 + c8: a
    // quarter-warp reads 128 contiguous bytes of the ring, writes 32 bytes into each of the four planes (distinct
    // banks, see ws_npos) and the warp stores 8 pixels x 64 B contiguous to x_out.
    const int xw = warp - (N_EPI + 2);
    const int c8 = lane & 3;
    const int ix = (xw & 7) * 8 + (lane >> 2);
    const int rr = xw >> 3;
    const int nchunks = rows_in / 2;
    float ca[8], cb[8];
    int cur_img = -1;
    int it = 0;
    uint32_t q = 0, qph = 0;   // ring slot of the running chunk counter and the parity of its lap
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int s = it % S;
      const TileAt ta = tile_at(t);
      const int img = ta.img, tx0 = ta.tx0;
      if (img != cur_img) {
        cur_img = img;
        const double sum = p.stats[(img * kGroups + c8) * 2 + 0];   // octet == GroupNorm group
        const double sq = p.stats[(img * kGroups + c8) * 2 + 1];
        const double mean = sum * p.inv_count;
        double var = sq * p.inv_count - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double rstd = gn_rstd(var);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const double gm = (double)__ldg(p.gamma + 8 * c8 + e);
          ca[e] = (float)(gm * rstd);
          cb[e] = (float)((double)__ldg(p.beta + 8 * c8 + e) - mean * gm * rstd);
        }
      }
      const int gx = tx0 - d + ix;
      const bool col_ok = gx >= 0 && gx < p.W;
      const bool col_int = col_ok && ix >= d && ix < d + TW;
      uint8_t* xpl = s_st + (size_t)s * stage_bytes + (size_t)c8 * plane_bytes + (size_t)ix * 16;
      // x_out address of tile row 0 (may lie above the image; only rows inside it are dereferenced)
      __half* xg = p.x_out != nullptr ? p.x_out + (ptrdiff_t)img * (ptrdiff_t)img_elems +
                                            ((ptrdiff_t)ta.y_first * p.W + (col_ok ? gx : 0)) * kC + 8 * c8
                                      : nullptr;
      tc::mbar_wait_warp(&s_empty[s], (uint32_t)(((it / S) & 1) ^ 1));   // the MMAs that read this stage completed
      for (int c = 0; c < nchunks; ++c) {
        tc::mbar_wait_warp(&s_raw[q], qph);   // this chunk's TMA boxes have landed
        if (c == 0) WS_STAMP(xw == 0 && lane == 0, 1);
        const int r = 2 * c + rr;
        const int gy = ta.y_first + rs * r;
        const uint8_t* raw = s_ring + (size_t)q * 2u * CHUNK_HALF + (size_t)(rr * PW + ix) * 64 + c8 * 16;
        uint4 h = make_uint4(0, 0, 0, 0);
        if (col_ok && gy >= 0 && gy < p.H && !(p.dbg & 1)) {
          float v[8];
          unpack8(*reinterpret_cast<const uint4*>(raw), v);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(v[e], ca[e], cb[e]));
          if (p.has_res) {
            float rs[8];
            unpack8(*reinterpret_cast<const uint4*>(raw + CHUNK_HALF), rs);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += rs[e];
          }
          h = pack8(v);
          if (xg != nullptr && col_int && r >= vs && r < vs + TH) stg_hint(xg + (ptrdiff_t)r * rs * p.W * kC, h, pol_keep);
        }
        if (!(p.dbg & 1)) *reinterpret_cast<uint4*>(xpl + (size_t)r * (PW * 16)) = h;
        mbar_arrive(&s_rfree[q]);   // this thread is done reading the slot
        if (++q == (uint32_t)RING) {
          q = 0;
          qph ^= 1u;
        }
        // M-tile mt reads tile rows 2 mt .. 2 mt + 1 + 2 vs, i.e. chunks up to mt + vs: complete with this chunk
        if (c >= vs) {
          tc::fence_proxy_async();
          mbar_arrive(&s_full[s][c - vs]);
        }
      }
      WS_STAMP(xw == 0 && lane == 0, 2);
    }
  } else if (warp == N_EPI + 1) {
    // =========================== TMA issue ===========================
    if (tc::elect_one()) {
      const uint32_t tx_bytes = (p.has_res ? 2u : 1u) * CHUNK_HALF;
      const int nchunks = rows_in / 2;
      int it = 0;
      uint32_t q = 0, qph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const TileAt ta = tile_at(t);
        const int img = ta.img, tx0 = ta.tx0;
        for (int c = 0; c < nchunks; ++c) {
          tc::mbar_wait(&s_rfree[q], qph ^ 1u);   // every transform warp has read the slot
          mbar_arrive_tx(&s_raw[q], tx_bytes);
          if (c == 0) WS_STAMP(true, 0);
          uint8_t* slot = s_ring + (size_t)q * 2u * CHUNK_HALF;
          // (polyphase: the map's element stride along y is the dilation, a box is image rows y, y + dil)
          tma_load_4d(slot, &tm_y, 0, tx0 - d, ta.y_first + rs * 2 * c, img, &s_raw[q]);
          if (p.has_res) tma_load_4d(slot + CHUNK_HALF, &tm_r, 0, tx0 - d, ta.y_first + rs * 2 * c, img, &s_raw[q]);
          if (++q == (uint32_t)RING) {
            q = 0;
            qph ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == N_EPI) {
    // =========================== MMA issue ===========================
    if (tc::elect_one()) {
      const uint32_t plane_u16 = plane_bytes >> 4;
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_st), plane_bytes, 128u);   // stage and plane strides: 16 B units
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 512u, 128u);
      tc::mbar_wait(&s_wbar, 0u);   // the bulk-copied weights have landed
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        const int a = it & 1;
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&s_tempty[a], aph ^ 1u);   // the epilogue has drained this accumulator set
        const uint64_t da = da0 + (uint64_t)(s * (stage_bytes >> 4));
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          tc::mbar_wait(&s_full[s][mt], ph);     // the transform warps have finished the rows under this M-tile
          if (mt == 0) WS_STAMP(true, 3);
          tc::fence_after_sync();
          const uint32_t dcol = tmem_base + (uint32_t)(a * ACC_COLS + mt * 32);
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t pos = (uint32_t)(mt * 128 + (tap / 3) * vs * PW + (tap % 3) * d);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                tc::mma_f16(dcol, da + (uint64_t)(2u * ks * plane_u16 + pos), db0 + (uint64_t)((tap * 2 + ks) * 64),
                            tc::idesc_f16(32), (tap | ks) != 0 ? 1u : 0u);
            }
          }
          tc::mma_commit(&s_tfull[a][mt]);       // this M-tile's accumulators are complete
        }
        tc::mma_commit(&s_empty[s]);
        WS_STAMP(true, 4);
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int wq = warp & 3;    // TMEM lane quarter
    const int mth = warp >> 2;  // this warp takes M-tiles mth, mth + 2, ...
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
      const TileAt ta = tile_at(t);
      const int img = ta.img, tx0 = ta.tx0;
      const int y_out0 = ta.y_first + rs * vs;   // image row of the tile's first output row
      if (img != cur_img) {
        flush(cur_img);
        cur_img = img;
      }
#pragma unroll 1
      for (int mt = mth; mt < MT; mt += N_EPI / 4) {
        tc::mbar_wait_warp(&s_tfull[a][mt], aph);
        tc::fence_after_sync();
        if (mt == mth) WS_STAMP(tid == 0, 5);
        const int j = mt * 128 + wq * 32 + lane;
        const int oy = y_out0 + rs * (j / PW), ox_t = j % PW, ox = tx0 + ox_t;
        const bool valid = ox_t < TW && ox < p.W && oy < p.H;
        float v[32];
        const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(a * ACC_COLS + mt * 32);
        tc::tmem_ld16(ta, v);
        tc::tmem_ld16(ta + 16u, v + 16);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            v[c] += s_bias[c];
            gsum[c >> 3] += v[c];
            gsq[c >> 3] = fmaf(v[c], v[c], gsq[c >> 3]);
          }
        }
        // Store through a lane transpose: in every store instruction two adjacent lanes write the two 16-byte halves
        // of one 32-byte sector (a lane writing its own pixel's 64 bytes with four instructions would touch every
        // sector twice, half-filled).  Instruction k covers pixels 16 (k >> 1) .. + 15, channel octets 2 (k & 1) + {0, 1}.
        uint4 ch[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ch[q] = pack8(v + 8 * q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int src = (lane >> 1) + 16 * (k >> 1);
          uint4 lo, hi;
          lo.x = __shfl_sync(0xffffffffu, ch[2 * (k & 1)].x, src);
          lo.y = __shfl_sync(0xffffffffu, ch[2 * (k & 1)].y, src);
          lo.z = __shfl_sync(0xffffffffu, ch[2 * (k & 1)].z, src);
          lo.w = __shfl_sync(0xffffffffu, ch[2 * (k & 1)].w, src);
          hi.x = __shfl_sync(0xffffffffu, ch[2 * (k & 1) + 1].x, src);
          hi.y = __shfl_sync(0xffffffffu, ch[2 * (k & 1) + 1].y, src);
          hi.z = __shfl_sync(0xffffffffu, ch[2 * (k & 1) + 1].z, src);
          hi.w = __shfl_sync(0xffffffffu, ch[2 * (k & 1) + 1].w, src);
          const int js = mt * 128 + wq * 32 + src;
          const int oys = y_out0 + rs * (js / PW), oxts = js % PW, oxs = tx0 + oxts;
          if (oxts < TW && oxs < p.W && oys < p.H && !(p.dbg & 4)) {
            __half* o = p.out + (size_t)img * img_elems + ((size_t)oys * p.W + oxs) * kC + 8 * (2 * (k & 1) + (lane & 1));
            stg_hint(o, (lane & 1) ? hi : lo, pol_keep);
          }
        }
      }
      tc::fence_before_sync();
      mbar_arrive(&s_tempty[a]);
      WS_STAMP(tid == 0, 6);
    }
    flush(cur_img);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (p.prof && blockIdx.x == 0 && tid == 0) g_ws_prof[57] = gtime();
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    cudaGetLastError();
  }
  return fn;
}

// [n][H][W][32] fp16, box = [1][2][64][32]: two rows of a halo-extended tile, zeros outside the image
bool make_map(const void* base, int n, int H, int W, int ystride, CUtensorMap* out) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)kC * 2, (cuuint64_t)W * kC * 2, (cuuint64_t)H * W * kC * 2};
  // element stride s along y: the box spans 2 s image rows and delivers every s-th of them, i.e. two rows
  const cuuint32_t box[4] = {(cuuint32_t)kC, (cuuint32_t)PW, (cuuint32_t)(2 * ystride), 1};
  const cuuint32_t estr[4] = {1, 1, (cuuint32_t)ystride, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WsPlan {
  int th, stages, ring, poly;
};
// Shared memory = weights + ring + stages.  Preference (measured on level 0, 512x640): two stages so that the loads /
// transform of the next tile overlap this tile's MMAs, and the taller the tile the smaller the share of halo rows
// that is loaded and transformed twice ((TH + 2 dil) / TH); a four-slot ring costs ~5 % against six slots, far less
// than a second stage or a taller tile gain at dilation 4 (level 0: 42.5 -> 34.0 us per layer).
bool plan_for(int dil, int H, int W, int n_img, WsPlan* plan) {
  // (16-row single-stage tiles at dilation 8 were measured equal to 8-row ones: 54.7 vs 54.2 us per level-0 layer)
  static const WsPlan prefs[] = {{8, 2, 6, 0}, {8, 2, 4, 0}, {4, 2, 6, 0}, {8, 1, 6, 0}, {4, 2, 4, 0}, {4, 1, 6, 0}};
  static const int force = getenv("B200MVS_WS_PLAN") ? atoi(getenv("B200MVS_WS_PLAN")) : -1;   // A/B: index into prefs
  static const bool no_poly = getenv("B200MVS_WS_NOPOLY") != nullptr;                          // A/B: round-1 tiles
  const bool poly = dil > 1 && !no_poly;
  const int vs = poly ? 1 : dil;
  for (int k = 0; k < (int)(sizeof(prefs) / sizeof(prefs[0])); ++k) {
    WsPlan c = prefs[k];
    c.poly = poly ? 1 : 0;
    if (force >= 0 && k < force && dil >= 4) continue;
    // worth it only when every SM gets at least one tile; smaller layers are latency bound either way
    const long long rows_tiles = poly ? (long long)dil * cdiv(cdiv(H, dil), c.th) : cdiv(H, c.th);
    static const int min_tiles = getenv("B200MVS_WS_MIN_TILES") ? atoi(getenv("B200MVS_WS_MIN_TILES")) : 148;   // A/B
    if ((long long)cdiv(W, PW - 2 * dil) * rows_tiles * n_img < min_tiles) continue;
    if (W_BYTES + ring_bytes(c.ring) + (size_t)c.stages * ws_stage_bytes(c.th, vs, dil) <= kSmemBudget) {
      *plan = c;
      return true;
    }
  }
  return false;
}

template <int TH>
int launch_th(const WsParams& q, const uint8_t* w16, const CUtensorMap& tm_y, const CUtensorMap& tm_r, int grid,
              size_t smem, int tag, cudaStream_t stream) {
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&conv3x3_ws_kernel<TH>), kSmemBudget)) return rc;
  if (tag != TAG_NONE) probe_before(tag, stream);
  launch_pdl(conv3x3_ws_kernel<TH>, dim3(grid), dim3(NT), smem, stream, q, w16, tm_y, tm_r);
  if (tag != TAG_NONE) probe_after(tag, stream);
  B200MVS_LAUNCH_OK("conv3x3_ws_kernel");
  if (q.prof) {
    long long h[58];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_ws_prof, sizeof(h));
    fprintf(stderr, "ws TH=%d dil=%d S=%d grid=%d: body %lld ns\n", TH, q.dil, q.stages, grid, h[57] - h[56]);
    for (int t = 0; t < 8 && t * (long long)grid < (long long)q.tiles_x * q.tiles_y * q.n_img; ++t) {
      fprintf(stderr, "  tile %d:", t);
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %lld", h[t * 7 + k] - h[56]);
      fprintf(stderr, "\n");
    }
  }
  return 0;
}

}  // namespace

bool conv3x3_ws_supported(const ConvParams& p) {
  if (!(p.feat.mode == FEAT_GN || p.feat.mode == FEAT_GN_RES) || !p.feat.half_io || !p.out_half || p.extra.n != 0)
    return false;
  if (p.Di != 1 || p.Do != 1 || p.Hi != p.Ho || p.Wi != p.Wo || p.dil < 1 || p.dil > 8) return false;
  if (p.add_src != nullptr || p.out_img_stride != 0 || p.feat.img_div != 1) return false;
  WsPlan plan;
  return plan_for(p.dil, p.Hi, p.Wi, p.n_img, &plan) && encode_fn() != nullptr;
}

int launch_conv3x3_ws(const ConvParams& p, const uint8_t* w16, cudaStream_t stream) {
  if (p.n_img <= 0) return 0;
  if (!conv3x3_ws_supported(p)) {
    set_error("launch_conv3x3_ws: unsupported configuration");
    return -1;
  }
  const bool has_res = p.feat.mode == FEAT_GN_RES;
  WsPlan plan;
  plan_for(p.dil, p.Hi, p.Wi, p.n_img, &plan);
  WsParams q;
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
  q.poly = plan.poly;
  q.tiles_y = plan.poly ? cdiv(cdiv(p.Ho, p.dil), plan.th) : cdiv(p.Ho, plan.th);
  q.stages = plan.stages;
  q.ring = plan.ring;
  q.has_res = has_res ? 1 : 0;
  static const bool prof = getenv("B200MVS_WS_PROFILE") != nullptr;
  q.prof = prof ? 1 : 0;
  static const int dbg = getenv("B200MVS_WS_DEBUG") ? atoi(getenv("B200MVS_WS_DEBUG")) : 0;
  q.dbg = dbg;
  CUtensorMap tm_y, tm_r;
  const int ystride = plan.poly ? p.dil : 1;
  if (!make_map(p.feat.ptr, p.n_img, p.Hi, p.Wi, ystride, &tm_y) ||
      !make_map(has_res ? (const void*)p.feat.resid : (const void*)p.feat.ptr, p.n_img, p.Hi, p.Wi, ystride, &tm_r)) {
    set_error("launch_conv3x3_ws: cuTensorMapEncodeTiled failed");
    return -1;
  }
  const size_t smem = W_BYTES + ring_bytes(plan.ring) +
                      (size_t)plan.stages * ws_stage_bytes(plan.th, plan.poly ? 1 : p.dil, p.dil);
  int num_sms = 0;
  if (int rc = current_device_sm_count(&num_sms)) return rc;
  const long long total = (long long)q.tiles_x * q.tiles_y * (plan.poly ? p.dil : 1) * q.n_img;
  const int grid = (int)(total < num_sms ? total : num_sms);
  if (plan.th == 8) return launch_th<8>(q, w16, tm_y, tm_r, grid, smem, p.tag, stream);
  return launch_th<4>(q, w16, tm_y, tm_r, grid, smem, p.tag, stream);
}

}  // namespace b200mvs
