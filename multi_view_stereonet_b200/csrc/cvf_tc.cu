// Conv3d 3x3x3 32->32 of the cost-volume filter (CostVolumeFilter.forward, multi_view_stereonet.py:341-353)
// on tcgen05 tensor cores with split-fp16 operands (fp32-class accuracy, see recurrence.cu).
//
// A CTA owns RT output rows of a column strip of width cw (RT * (cw+2) <= 128 positions = one UMMA M-tile; images
// wider than 48 pixels are cut into equal strips whose halo columns come from the neighbouring strip) of a chunk of
// depth slices of one volume and streams through depth: every input slice tile ((RT+2) rows, previous layer's GroupNorm + LeakyReLU
// applied, split into hi/lo fp16 planes) is staged ONCE into a 4-slot shared-memory ring and used by the three
// output slices that touch it.  Output slice q = 27 taps x 2 k-steps x 2 MMAs reading ring slots q, q+1, q+2 with
// the shifted-window descriptors of conv_tc.cu; accumulators are double-buffered in TMEM.
// Warp-specialised: eight worker warps stage slices and drain accumulators, a ninth warp only issues MMAs.  One
// slice's 108 MMAs take ~4.9 k cycles of the tensor pipe and nearly as long to ISSUE (tools/mma_bench.cu: 45 cycles
// per MMA at N = 64 / 32); with the issue loop on a worker warp every iteration paid issue time plus staging plus
// epilogue (~9 k cycles, round 1), now the tensor pipe runs back to back and the workers hide under it.  The roles
// meet only at mbarriers: full[slot] (workers -> issuer), acc_full[buf] (tcgen05.commit -> workers), acc_empty[buf]
// (workers -> issuer).
#include <cstdlib>
#include <vector>

#include "conv.cuh"
#include "cvf_tc.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int NW = 256;            // worker threads (8 warps): staging + epilogue
constexpr int NT = NW + 32;        // + the MMA-issuing warp
constexpr int MAX_TASKS = 4;
constexpr int W_BLOCKS = 27 * 2;
constexpr int W_BYTES = W_BLOCKS * 2048;
constexpr int RING = 4;

struct Geo {
  int PW, RT, NP, np_pad;
  uint32_t plane_bytes, slot_bytes, total;
};
__host__ __device__ inline Geo make_geo(int w) {
  Geo g;
  g.PW = w + 2;
  g.RT = 128 / g.PW;
  g.NP = (g.RT + 2) * g.PW + 2;
  // the MMA reads 128 rows from every tap offset (up to 2*PW + 2) even when RT * PW < 128
  if (g.NP < 128 + 2 * g.PW + 2) g.NP = 128 + 2 * g.PW + 2;
  // padded to 2 (mod 8) positions: consecutive planes then start 32 bytes apart modulo 128, so the four octet planes
  // the lanes of a quad write at once fall into distinct shared-memory banks
  g.np_pad = ((g.NP + 5) & ~7) + 2;
  g.plane_bytes = (uint32_t)g.np_pad * 16u;
  g.slot_bytes = 8u * g.plane_bytes;  // hi planes 0..3, lo planes 4..7
  g.total = (uint32_t)W_BYTES + RING * g.slot_bytes;
  return g;
}

constexpr int kMaxStrip = 48;   // widest strip whose 4-slot ring fits next to the weights in shared memory
__host__ __device__ inline int strip_count(int w) { return (w + kMaxStrip - 1) / kMaxStrip; }
__host__ __device__ inline int strip_width(int w) { return (w + strip_count(w) - 1) / strip_count(w); }

struct CvfParams {
  CvfArgs a;
  int row_tiles, col_tiles, cw;
  int share;   // work units (one output slice of one (volume, row tile, column strip) column) per CTA
  int chunks;  // > 0: column-aligned decomposition, CTA = chunk (blockIdx.x % chunks) of column (blockIdx.x / chunks)
  int dbg;     // timing ablations (wrong results): 1 no MMAs, 2 no operand staging stores, 4 no output stores, 8 no input loads
};

__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NW) : "memory"); }

// Work decomposition.  A column = all D output slices of one (volume n, row tile, column strip); the n * row_tiles *
// col_tiles columns are laid end to end and cut into gridDim.x equal ranges of `share` units, so every SM gets the
// same number of slices whatever the batch size (a column-per-CTA grid left 60 of 148 SMs idle at batch 8).  A CTA
// walks its range as segments (the part of a column it covers); each segment restarts the depth pipeline (two extra
// staged slices), the ring / accumulator counters run on across segments.
__global__ void __launch_bounds__(NT, 1) cvf_tc_kernel(const CvfParams P) {
  const CvfArgs& p = P.a;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_a[kC], s_b[kC], s_bias[kC];
  __shared__ double s_stats[2 * kGroups];
  __shared__ __align__(8) uint64_t s_full[RING], s_acc_full[2], s_acc_empty[2], s_wbar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
  const Geo g = make_geo(P.cw);
  const int PW = g.PW, RT = g.RT, NP = g.NP;
  const int cols_per_n = P.row_tiles * P.col_tiles;
  const long long total = (long long)p.n * cols_per_n * p.D;
  long long u_begin, u_end;
  if (P.chunks > 0) {
    const long long col = blockIdx.x / P.chunks;
    u_begin = col * p.D + (long long)(blockIdx.x % P.chunks) * P.share;
    u_end = u_begin + P.share < (col + 1) * p.D ? u_begin + P.share : (col + 1) * p.D;
  } else {
    u_begin = (long long)blockIdx.x * P.share;
    u_end = u_begin + P.share < total ? u_begin + P.share : total;
  }
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + W_BYTES;

  if (warp == 0) tc::tmem_alloc(&s_tmem, 128u);
  if (tid == 32) {
    for (int i = 0; i < RING; ++i) tc::mbar_init(&s_full[i], NW / 32);   // one arrival per worker warp
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_acc_full[i], 1);
      tc::mbar_init(&s_acc_empty[i], NW / 32);
    }
    tc::mbar_init(&s_wbar, 1);
    tc::mbar_init_fence();
    tc::bulk_load_weights(s_w, p.w16, (uint32_t)W_BYTES, &s_wbar);   // constant data: before griddepcontrol.wait
  }
  if (tid < kC) s_bias[tid] = p.bias != nullptr ? __ldg(p.bias + tid) : 0.f;
  if (tid < 2 * kGroups) s_stats[tid] = 0.0;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_launch_dependents();   // after the TMEM allocation (see common.cuh)
  // (no ring initialisation: every position an MMA reads, [0, NP), is written by the staging of its slice)
  pdl_wait();
  const uint32_t tmem_base = s_tmem;
  const uint32_t plane_u16 = g.plane_bytes >> 4, slot_u16 = g.slot_bytes >> 4;
  const size_t slice_elems = (size_t)p.h * p.w * kC;

  if (warp == NW / 32) {
    // ================= MMA issuer: output slice q of a segment once its input slices q .. q+2 are staged and its
    //                   accumulator is free
    if (tc::elect_one()) {
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_ring), g.plane_bytes, 128u);
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 1024u, 128u);
      tc::mbar_wait(&s_wbar, 0u);   // the bulk-copied weights have landed
      uint32_t sc = 0, oc = 0;      // staged slices / output slices of all earlier segments
      for (long long u = u_begin; u < u_end;) {
        const int d0 = (int)(u % p.D);
        const int dcount = (int)((long long)(p.D - d0) < u_end - u ? (long long)(p.D - d0) : u_end - u);
        for (int q = 0; q < dcount; ++q) {
          const uint32_t S = sc + (uint32_t)q + 2u, O = oc + (uint32_t)q;
          tc::mbar_wait(&s_full[S & (RING - 1)], (S / RING) & 1u);
          if (O >= 2) tc::mbar_wait(&s_acc_empty[O & 1], ((O >> 1) - 1u) & 1u);
          tc::fence_after_sync();
          const uint32_t acc = tmem_base + (O & 1u) * 64u;
#pragma unroll
          for (int kz = 0; kz < ((P.dbg & 1) ? 0 : 3); ++kz) {
            const uint64_t da_slot = da0 + (uint64_t)(((sc + (uint32_t)(q + kz)) & (RING - 1)) * slot_u16);
#pragma unroll
            for (int t2 = 0; t2 < 9; ++t2) {
              const uint32_t pos = (uint32_t)((t2 / 3) * PW + (t2 % 3));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t a_hi = da_slot + (uint64_t)(2 * ks * plane_u16 + pos);
                const uint64_t a_lo = a_hi + (uint64_t)(4 * plane_u16);
                const uint64_t b = db0 + (uint64_t)((((kz * 9 + t2) * 2) + ks) * 128);
                tc::mma_f16(acc, a_hi, b, tc::idesc_f16(64), (kz | t2 | ks) != 0 ? 1u : 0u);
                tc::mma_f16(acc, a_lo, b, tc::idesc_f16(32), 1u);
              }
            }
          }
          tc::mma_commit(&s_acc_full[O & 1]);   // implies tcgen05.fence::before_thread_sync
        }
        sc += (uint32_t)dcount + 2u;
        oc += (uint32_t)dcount;
        u += dcount;
      }
    }
    __syncwarp();
  } else {
    // ================= workers: staging + epilogue
    const int t_oct = tid & 3;
    const int wq = warp & 3, chalf = warp >> 2;
    const int jl = wq * 32 + lane;
    const int e_oy = jl / PW, e_ox = jl % PW;
    const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(chalf * 16);
    uint32_t sc = 0, oc = 0;
    int cur_n = -1;
    for (long long u = u_begin; u < u_end;) {
      const int col = (int)(u / p.D), d0 = (int)(u % p.D);
      const int dcount = (int)((long long)(p.D - d0) < u_end - u ? (long long)(p.D - d0) : u_end - u);
      const int n = col / cols_per_n, rc = col - n * cols_per_n;
      const int rt = rc / P.col_tiles, ct = rc - rt * P.col_tiles;
      const int y0 = rt * RT, x0 = ct * P.cw;
      if (n != cur_n) {
        // GroupNorm coefficients of volume n (the previous segment's staging finished before its last epilogue)
        worker_barrier();
        if (tid < kC && p.mode >= FEAT_GN) {
          const int grp = tid >> 3;
          const double sum = p.stats[(n * kGroups + grp) * 2 + 0];
          const double sq = p.stats[(n * kGroups + grp) * 2 + 1];
          const double mean = sum * p.inv_count;
          const double var = sq * p.inv_count - mean * mean;
          const double rstd = gn_rstd(var);
          s_a[tid] = (float)((double)p.gamma[tid] * rstd);
          s_b[tid] = (float)((double)p.beta[tid] - mean * (double)p.gamma[tid] * rstd);
        }
        worker_barrier();
        cur_n = n;
      }
      // staging tasks of this thread (same positions for every slice of the segment)
      int t_l[MAX_TASKS];
      size_t t_off[MAX_TASKS];
      bool t_in[MAX_TASKS], t_real[MAX_TASKS];
#pragma unroll
      for (int k = 0; k < MAX_TASKS; ++k) {
        const int i = tid + k * NW;
        t_l[k] = i >> 2;
        const int iy = t_l[k] / PW, ix = t_l[k] % PW;
        const int gy = y0 - 1 + iy, gx = x0 + ix - 1;
        t_in[k] = t_l[k] < NP;
        t_real[k] = t_in[k] && iy < RT + 2 && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
        t_off[k] = t_real[k] ? ((size_t)gy * p.w + gx) * kC + 8 * t_oct : 0;
      }
      // epilogue slice of this thread: one output position, 16 channels
      const bool e_real = e_oy < RT && e_ox < P.cw && (x0 + e_ox) < p.w && (y0 + e_oy) < p.h;
      const size_t e_off = e_real ? ((size_t)(y0 + e_oy) * p.w + x0 + e_ox) * kC + chalf * 16 : 0;
      const float* in_n = p.in + (size_t)n * p.D * slice_elems;
      float* out_n = p.out + (size_t)n * p.D * slice_elems;
      // statistics: float32 within one slice (8 values per group), float64 across slices -- the result must not
      // depend on how many slices a segment happens to hold
      double gs[2] = {0.0, 0.0}, gq[2] = {0.0, 0.0};

      // Software pipeline over the segment's depth range.  Iteration `it`:
      //   transform + stage input slice it (its global loads were issued one iteration earlier) -> next ring slot,
      //     last read by the MMAs of the output slice four back, whose completion this thread observed in its epilogue
      //   issue the global loads of input slice it + 1
      //   epilogue of output slice it - 3
      float y8[MAX_TASKS][8];
      bool loaded_valid = false;
      auto issue_loads = [&](int slice) {
        const int din = d0 - 1 + slice;
        loaded_valid = slice <= dcount + 1 && din >= 0 && din < p.D && !(P.dbg & 8);
        if (loaded_valid) {
          const float* src = in_n + (size_t)din * slice_elems;
#pragma unroll
          for (int k = 0; k < MAX_TASKS; ++k) {
            if (t_real[k]) ld8_nc(src + t_off[k], y8[k]);   // one 256-bit load per (position, octet)
          }
        }
      };
      issue_loads(0);
      for (int it = 0; it <= dcount + 2; ++it) {
        if (it <= dcount + 1) {
          const uint32_t S = sc + (uint32_t)it;
          uint8_t* slot = s_ring + (size_t)(S & (RING - 1)) * g.slot_bytes;
#pragma unroll
          for (int k = 0; k < MAX_TASKS; ++k) {
            if (t_in[k]) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = 0.f;
              if (t_real[k] && loaded_valid) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = y8[k][e];
                if (p.mode >= FEAT_GN) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(v[e], s_a[8 * t_oct + e], s_b[8 * t_oct + e]));
                }
              }
              uint4 hi, lo;
              tc::split8(v, &hi, &lo);
              if (!(P.dbg & 2)) {
                *reinterpret_cast<uint4*>(slot + (size_t)t_oct * g.plane_bytes + (size_t)t_l[k] * 16) = hi;
                *reinterpret_cast<uint4*>(slot + (size_t)(4 + t_oct) * g.plane_bytes + (size_t)t_l[k] * 16) = lo;
              }
            }
          }
          tc::fence_proxy_async();   // this thread's operand stores -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&s_full[S & (RING - 1)]);
        }

        issue_loads(it + 1);

        const int qe = it - 3;
        if (qe >= 0 && qe < dcount) {
          const uint32_t O = oc + (uint32_t)qe;
          tc::mbar_wait_warp(&s_acc_full[O & 1], (O >> 1) & 1u);
          tc::fence_after_sync();
          float v[16], c[16];
          const uint32_t acc = tmem_my + (O & 1u) * 64u;
          tc::tmem_ld16(acc, v);
          tc::tmem_ld16(acc + 32u, c);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&s_acc_empty[O & 1]);   // the accumulator may be overwritten
          if (e_real && !(P.dbg & 4)) {
            float* dst = out_n + (size_t)(d0 + qe) * slice_elems + e_off;
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = (v[k] + c[k]) + s_bias[chalf * 16 + k];
            st8(dst, v);
            st8(dst + 8, v + 8);
            float fs[2] = {0.f, 0.f}, fq[2] = {0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              fs[0] += v[k];
              fq[0] += v[k] * v[k];
              fs[1] += v[8 + k];
              fq[1] += v[8 + k] * v[8 + k];
            }
            gs[0] += (double)fs[0];
            gq[0] += (double)fq[0];
            gs[1] += (double)fs[1];
            gq[1] += (double)fq[1];
          }
        }
      }
      // ---- GroupNorm statistics of what this segment stored (volume n) ----
      if (p.out_stats != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs[0] += __shfl_xor_sync(0xffffffffu, gs[0], o);
          gq[0] += __shfl_xor_sync(0xffffffffu, gq[0], o);
          gs[1] += __shfl_xor_sync(0xffffffffu, gs[1], o);
          gq[1] += __shfl_xor_sync(0xffffffffu, gq[1], o);
        }
        if (lane == 0) {
          atomicAdd(&s_stats[(2 * chalf) * 2 + 0], gs[0]);
          atomicAdd(&s_stats[(2 * chalf) * 2 + 1], gq[0]);
          atomicAdd(&s_stats[(2 * chalf + 1) * 2 + 0], gs[1]);
          atomicAdd(&s_stats[(2 * chalf + 1) * 2 + 1], gq[1]);
        }
        worker_barrier();
        if (tid < 2 * kGroups) {
          atomicAdd(p.out_stats + (size_t)n * 2 * kGroups + tid, s_stats[tid]);
          s_stats[tid] = 0.0;
        }
        // No warp can reach the next segment's statistics before these threads are through (every staged slice of that
        // segment needs an arrival of this warp, which comes after the lines above in program order); the barrier makes
        // that explicit for compute-sanitizer's racecheck, once per segment.
        worker_barrier();
      }
      sc += (uint32_t)dcount + 2u;
      oc += (uint32_t)dcount;
      u += dcount;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 128u);
}

}  // namespace

void pack_cvf_tc_weights(const float* w_oidhw, std::vector<uint8_t>* out) {
  out->assign(W_BYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  for (int tap = 0; tap < 27; ++tap)
    for (int nn = 0; nn < 32; ++nn)
      for (int c = 0; c < 32; ++c)
        tc::put_split_weight(h, tap * 2 + c / 16, c % 16, nn, w_oidhw[((size_t)nn * 32 + c) * 27 + tap]);
}

bool cvf_tc_supported(int h, int w) {
  const Geo g = make_geo(strip_width(w));
  return w >= 1 && g.PW <= 128 && g.RT >= 1 && g.NP * 4 <= MAX_TASKS * NW && g.total + 2048 <= 227 * 1024 && h >= 1;
}

int launch_cvf_tc(const CvfArgs& a, cudaStream_t stream) {
  if (a.n <= 0) return 0;
  if (!cvf_tc_supported(a.h, a.w)) {
    set_error("launch_cvf_tc: shape not supported");
    return -1;
  }
  const Geo g = make_geo(strip_width(a.w));
  if (int rc = ensure_func_smem(reinterpret_cast<const void*>(&cvf_tc_kernel), g.total)) return rc;
  CvfParams P;
  P.a = a;
  P.row_tiles = cdiv(a.h, g.RT);
  P.col_tiles = strip_count(a.w);
  P.cw = strip_width(a.w);
  static const int dbg = getenv("B200MVS_CVF_DEBUG") ? atoi(getenv("B200MVS_CVF_DEBUG")) : 0;
  P.dbg = dbg;
  // One CTA per SM (220 KB of shared memory).  Every CTA gets the same number of output slices; a range shorter than
  // ~4 slices would spend more on its two halo slices and the pipeline fill than on its outputs.
  int num_sms = 0;
  if (int rc = current_device_sm_count(&num_sms)) return rc;
  const long long columns = (long long)a.n * P.row_tiles * P.col_tiles;
  const long long total = columns * a.D;
  int grid_x;
  // Few columns (small batch): cut every column into the same number of depth chunks if that fills the chip -- a
  // CTA is then one segment.  Otherwise equal ranges of the concatenated columns (a range may span two columns).
  int chunks = (int)(num_sms / columns);
  if (chunks > a.D) chunks = a.D;
  if (chunks >= 1 && columns * chunks * 100 >= (long long)num_sms * 90) {
    P.share = cdiv(a.D, chunks);
    P.chunks = cdiv(a.D, P.share);
    grid_x = (int)(columns * P.chunks);
  } else {
    long long ctas = total / 4 < 1 ? 1 : total / 4;
    if (ctas > num_sms) ctas = num_sms;
    P.share = (int)((total + ctas - 1) / ctas);
    P.chunks = 0;
    grid_x = (int)((total + P.share - 1) / P.share);
  }
  dim3 grid(grid_x, 1);
  if (a.tag != TAG_NONE) probe_before(a.tag, stream);
  launch_pdl(cvf_tc_kernel, grid, dim3(NT), (size_t)g.total, stream, P);
  if (a.tag != TAG_NONE) probe_after(a.tag, stream);
  B200MVS_LAUNCH_OK("cvf_tc_kernel");
  return 0;
}

}  // namespace b200mvs
