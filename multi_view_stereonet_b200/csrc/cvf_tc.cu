// Conv3d 3x3x3 32->32 of the cost-volume filter (CostVolumeFilter.forward, multi_view_stereonet.py:341-353)
// on tcgen05 tensor cores with split-fp16 operands (fp32-class accuracy, see recurrence.cu).
//
// A CTA owns RT output rows of a column strip of width cw (RT * (cw+2) <= 128 positions = one UMMA M-tile; images
// wider than 48 pixels are cut into equal strips whose halo columns come from the neighbouring strip) of a chunk of
// depth slices of one volume and streams through depth: every input slice tile ((RT+2) rows, previous layer's GroupNorm + LeakyReLU
// applied, split into hi/lo fp16 planes) is staged ONCE into a 4-slot shared-memory ring and read by ONE batch of MMAs.
//
// Sliding accumulator window: input slice i contributes to output slices i - 1, i, i + 1 through the taps kz = 2, 1, 0.
// The accumulators of consecutive output slices sit in consecutive 32-column slots of tensor memory (a ring of 16), and
// the weights of the three kz taps of one in-plane tap are stacked along N (96 rows, kz = 2 first), so one N = 96 MMA
// per (in-plane tap, k-step, split-precision product) adds a slice's contribution to all three outputs at once:
// 9 x 2 x 3 = 54 MMAs per slice instead of 108 (N = 64 / 32) per output slice, and the 4 KB A operand is fetched once
// for three outputs.  The MMA rate is set by the operand bytes read from shared memory, (A + B) / 128 per clock
// (tools/mma_bench.cu: N = 32 43 cycles, 64 55, 96 60, 128 69); a batch is 3.25 k cycles against 4.87 k before.  The
// output slice a batch starts takes its first product as an MMA of its own with accumulate = 0; a window that wraps around
// the ring (2 batches of 16) is two runs (N = 64 + 32), at the ends of a segment the entries outside it are left out
// (N = 32 / 64).  An output slice is complete after the batch of the input slice behind it: one tcgen05.commit per batch.
// Warp-specialised: sixteen worker warps stage slices and drain accumulators (8 channels of one position per thread), a
// seventeenth warp only issues MMAs.  The roles meet only at mbarriers: full[slot] (workers -> issuer), acc_full[slot]
// (tcgen05.commit -> workers), acc_empty[slot] (workers -> issuer).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv.cuh"
#include "cvf_tc.cuh"
#include "tc_common.cuh"

namespace b200mvs {
namespace {

constexpr int NW = 512;            // worker threads (16 warps): staging + epilogue
constexpr int NT = NW + 32;        // + the MMA-issuing warp
constexpr int MAX_TASKS = 2;
constexpr int W_BLOCKS = 27 * 2;
constexpr int W_BYTES = W_BLOCKS * 2048;
constexpr int RING = 4;
constexpr int ACC = 16;            // accumulator slots in tensor memory (32 columns each, a ring over the output slices)
// The workers drain output slice it - LAG in the iteration that stages input slice it: the batch that completes it
// (input slice it - LAG + 2) was issued two iterations earlier, so neither role waits for the other in steady state.
constexpr int LAG = 4;
constexpr uint32_t TMEM_COLS = 32u * ACC;

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

__device__ __forceinline__ uint32_t idesc_n(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

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
  __shared__ __align__(8) uint64_t s_full[RING], s_free[RING], s_acc_full[ACC], s_acc_empty[ACC], s_wbar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = tc::uniform_warp_index();
#ifdef CVF_PROF
  unsigned long long pk_t0, pk_t1, pk_t2;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pk_t0));
#endif
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

  if (warp == 0) tc::tmem_alloc(&s_tmem, TMEM_COLS);
  if (tid == 32) {
    for (int i = 0; i < RING; ++i) {
      tc::mbar_init(&s_full[i], NW / 32);   // one arrival per worker warp
      tc::mbar_init(&s_free[i], 1);         // tcgen05.commit behind the batch that read the slot
    }
    for (int i = 0; i < ACC; ++i) {
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
#ifdef CVF_PROF
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pk_t1));
#endif
  const uint32_t tmem_base = s_tmem;
  const uint32_t plane_u16 = g.plane_bytes >> 4, slot_u16 = g.slot_bytes >> 4;
  const size_t slice_elems = (size_t)p.h * p.w * kC;

  if (warp == NW / 32) {
    // ================= MMA issuer: the batch of input slice i (its contributions to output slices i - 2, i - 1, i)
    //                   once the slice is staged and the accumulator slot of the output it starts is drained
    if (tc::elect_one()) {
      const uint64_t da0 = tc::umma_desc(tc::smem_u32(s_ring), g.plane_bytes, 128u);
      const uint64_t db0 = tc::umma_desc(tc::smem_u32(s_w), 96u * 16u, 128u);
      tc::mbar_wait(&s_wbar, 0u);   // the bulk-copied weights have landed
      uint32_t sc = 0, oc = 0;      // staged slices / output slices of all earlier segments
#ifdef CVF_PROF
      long long pf_full = 0, pf_empty = 0, pf_issue = 0, pf_n = 0, pf_t0 = clock64();
#endif
      for (long long u = u_begin; u < u_end;) {
        const int d0 = (int)(u % p.D);
        const int dcount = (int)((long long)(p.D - d0) < u_end - u ? (long long)(p.D - d0) : u_end - u);
        for (int i = 0; i <= dcount + 1; ++i) {
          const uint32_t S = sc + (uint32_t)i;
          const bool fresh = i < dcount;   // output slice i receives its first contribution (kz = 0) from this slice
#ifdef CVF_PROF
          const long long pa = clock64();
#endif
          tc::mbar_wait(&s_full[S & (RING - 1)], (S / RING) & 1u);
#ifdef CVF_PROF
          const long long pb = clock64();
#endif
          if (fresh) {
            const uint32_t O = oc + (uint32_t)i;
            if (O >= ACC) tc::mbar_wait(&s_acc_empty[O & (ACC - 1)], ((O / ACC) - 1u) & 1u);
          }
#ifdef CVF_PROF
          const long long pc = clock64();
#endif
          tc::fence_after_sync();
          // Window entry j = 0, 1, 2 is output slice i - 2 + j (weights kz = 2 - j) in accumulator slot (oc + i - 2 + j)
          // mod ACC; all three split-precision products (A_hi W_hi, A_lo W_hi, A_hi W_lo) add into the same columns.
          const uint32_t slot0 = (oc + (uint32_t)i - 2u) & (ACC - 1);
          const uint64_t da_slot = da0 + (uint64_t)((S & (RING - 1)) * slot_u16);
          if (!(P.dbg & 1)) {   // (bit 0: timing ablation without MMAs)
            // Entries inside the segment: jlo .. jhi.  Run A = those up to the end of the ring, run B = the rest of a
            // window that wraps around it (slot 0 on); one MMA per run and product, N = 32 per entry.  The single
            // issuing thread needs ~6 cycles per uniform-datapath instruction, predicated off or not, so the two cases
            // are separate unrolled loops without predicates.
            const int jlo = i >= 2 ? 0 : 2 - i;
            const int jhi = dcount + 1 - i < 2 ? dcount + 1 - i : 2;
            const uint32_t slotA = (slot0 + (uint32_t)jlo) & (ACC - 1);
            int lenA = jhi - jlo + 1;
            if ((int)(ACC - slotA) < lenA) lenA = (int)(ACC - slotA);
            const int lenB = jhi - jlo + 1 - lenA;
            const uint32_t dA = tmem_base + 32u * slotA, dB = tmem_base;
            const uint64_t bA = (uint64_t)(32 * jlo), bB = (uint64_t)(32 * (jlo + lenA));   // 32 weight rows per entry
            const uint32_t iA = idesc_n(32u * (uint32_t)lenA), iB = idesc_n(32u * (uint32_t)(lenB > 0 ? lenB : 1));
            // First product (A_hi W_hi of tap 0, k-step 0).  A batch that starts an output slice (entry 2) overwrites
            // that entry and accumulates into the others: one N = 32 MMA per entry, once per batch.
            if (fresh) {
#pragma unroll
              for (int j = 0; j < 3; ++j)
                if (j >= jlo)
                  tc::mma_f16(tmem_base + 32u * ((slot0 + (uint32_t)j) & (ACC - 1)), da_slot, db0 + (uint64_t)(32 * j),
                              tc::idesc_f16(32), j == 2 ? 0u : 1u);
            } else {
              tc::mma_f16(dA, da_slot, db0 + bA, iA, 1u);
              if (lenB > 0) tc::mma_f16(dB, da_slot, db0 + bB, iB, 1u);
            }
            if (lenB == 0) {
#pragma unroll
              for (int t2 = 0; t2 < 9; ++t2) {
                const uint32_t pos = (uint32_t)((t2 / 3) * PW + (t2 % 3));
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t a_hi = da_slot + (uint64_t)(2 * ks * plane_u16 + pos);
                  const uint64_t a_lo = a_hi + (uint64_t)(4 * plane_u16);
                  const uint64_t b_hi = db0 + (uint64_t)((t2 * 2 + ks) * 384) + bA;   // [W_hi | W_lo] x 2 k-octets x 96 rows
                  const uint64_t b_lo = b_hi + 192u;
                  if ((t2 | ks) != 0) tc::mma_f16(dA, a_hi, b_hi, iA, 1u);
                  tc::mma_f16(dA, a_lo, b_hi, iA, 1u);
                  tc::mma_f16(dA, a_hi, b_lo, iA, 1u);
                }
              }
            } else {
#pragma unroll
              for (int t2 = 0; t2 < 9; ++t2) {
                const uint32_t pos = (uint32_t)((t2 / 3) * PW + (t2 % 3));
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t a_hi = da_slot + (uint64_t)(2 * ks * plane_u16 + pos);
                  const uint64_t a_lo = a_hi + (uint64_t)(4 * plane_u16);
                  const uint64_t b_hi = db0 + (uint64_t)((t2 * 2 + ks) * 384);
                  const uint64_t b_lo = b_hi + 192u;
                  if ((t2 | ks) != 0) {
                    tc::mma_f16(dA, a_hi, b_hi + bA, iA, 1u);
                    tc::mma_f16(dB, a_hi, b_hi + bB, iB, 1u);
                  }
                  tc::mma_f16(dA, a_lo, b_hi + bA, iA, 1u);
                  tc::mma_f16(dB, a_lo, b_hi + bB, iB, 1u);
                  tc::mma_f16(dA, a_hi, b_lo + bA, iA, 1u);
                  tc::mma_f16(dB, a_hi, b_lo + bB, iB, 1u);
                }
              }
            }
          }
          // output slice i - 2 is complete (implies tcgen05.fence::before_thread_sync)
          if (i >= 2) tc::mma_commit(&s_acc_full[(oc + (uint32_t)i - 2u) & (ACC - 1)]);
          tc::mma_commit(&s_free[S & (RING - 1)]);   // the ring slot may be overwritten
#ifdef CVF_PROF
          pf_full += pb - pa;
          pf_empty += pc - pb;
          pf_issue += clock64() - pc;
          ++pf_n;
#endif
        }
        sc += (uint32_t)dcount + 2u;
        oc += (uint32_t)dcount;
        u += dcount;
      }
#ifdef CVF_PROF
      if (blockIdx.x == 0 || blockIdx.x == 77)
        printf("cvf issuer cta %d: %lld batches, per batch: wait full %lld, wait acc_empty %lld, issue %lld, total %lld\n",
               (int)blockIdx.x, pf_n, pf_full / pf_n, pf_empty / pf_n, pf_issue / pf_n, (clock64() - pf_t0) / pf_n);
#endif
    }
    __syncwarp();
  } else {
    // ================= workers: staging + epilogue
    const int t_oct = tid & 3;
    const int wq = warp & 3, cq = warp >> 2;   // TMEM lane quarter, GroupNorm group (8 output channels)
    const int jl = wq * 32 + lane;
    const int e_oy = jl / PW, e_ox = jl % PW;
    const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cq * 8);
    uint32_t sc = 0, oc = 0;
    int cur_n = -1;
#ifdef CVF_PROF
    long long pw_stage = 0, pw_loads = 0, pw_wait = 0, pw_epi = 0, pw_n = 0, pw_t0 = clock64();
#endif
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
      // epilogue slice of this thread: one output position, 8 channels (one GroupNorm group)
      const bool e_real = e_oy < RT && e_ox < P.cw && (x0 + e_ox) < p.w && (y0 + e_oy) < p.h;
      const size_t e_off = e_real ? ((size_t)(y0 + e_oy) * p.w + x0 + e_ox) * kC + cq * 8 : 0;
      const float* in_n = p.in + (size_t)n * p.D * slice_elems;
      float* out_n = p.out + (size_t)n * p.D * slice_elems;
      // statistics: float32 within one slice (8 values per group), float64 across slices -- the result must not
      // depend on how many slices a segment happens to hold
      double gs = 0.0, gq = 0.0;

      // Software pipeline over the segment's depth range.  Iteration `it`:
      //   transform + stage input slice it (its global loads were issued one iteration earlier) -> next ring slot,
      //     once the batch of the slice four back has released it (free[slot])
      //   issue the global loads of input slice it + 1
      //   epilogue of output slice it - LAG
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
      for (int it = 0; it <= dcount + LAG - 1; ++it) {
#ifdef CVF_PROF
        const long long wa = clock64();
#endif
        if (it <= dcount + 1) {
          const uint32_t S = sc + (uint32_t)it;
          uint8_t* slot = s_ring + (size_t)(S & (RING - 1)) * g.slot_bytes;
          if (S >= RING) tc::mbar_wait_warp(&s_free[S & (RING - 1)], ((S / RING) - 1u) & 1u);
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

#ifdef CVF_PROF
        const long long wb = clock64();
#endif
        issue_loads(it + 1);
#ifdef CVF_PROF
        const long long wc = clock64();
        long long wd = wc;
#endif

        const int qe = it - LAG;
        if (qe >= 0 && qe < dcount) {
          const uint32_t O = oc + (uint32_t)qe;
          tc::mbar_wait_warp(&s_acc_full[O & (ACC - 1)], (O / ACC) & 1u);
#ifdef CVF_PROF
          wd = clock64();
#endif
          tc::fence_after_sync();
          float v[8];
          tc::tmem_ld8(tmem_my + (O & (ACC - 1)) * 32u, v);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&s_acc_empty[O & (ACC - 1)]);   // the accumulator may be overwritten
          if (e_real && !(P.dbg & 4)) {
            float* dst = out_n + (size_t)(d0 + qe) * slice_elems + e_off;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += s_bias[cq * 8 + k];
            st8(dst, v);
            float fs = 0.f, fq = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              fs += v[k];
              fq += v[k] * v[k];
            }
            gs += (double)fs;
            gq += (double)fq;
          }
        }
#ifdef CVF_PROF
        pw_stage += wb - wa;
        pw_loads += wc - wb;
        pw_wait += wd - wc;
        pw_epi += clock64() - wd;
        ++pw_n;
#endif
      }
      // ---- GroupNorm statistics of what this segment stored (volume n) ----
      if (p.out_stats != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs += __shfl_xor_sync(0xffffffffu, gs, o);
          gq += __shfl_xor_sync(0xffffffffu, gq, o);
        }
        if (lane == 0) {
          atomicAdd(&s_stats[cq * 2 + 0], gs);
          atomicAdd(&s_stats[cq * 2 + 1], gq);
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
#ifdef CVF_PROF
    if ((blockIdx.x == 0 || blockIdx.x == 77) && (tid == 0 || tid == 511))
      printf("cvf worker cta %d tid %d: %lld iterations, per iteration: stage %lld, issue loads %lld, wait acc_full %lld, "
             "epilogue %lld, total %lld\n", (int)blockIdx.x, tid, pw_n, pw_stage / pw_n, pw_loads / pw_n, pw_wait / pw_n,
             pw_epi / pw_n, (clock64() - pw_t0) / pw_n);
#endif
  }
  tc::fence_before_sync();
  __syncthreads();
#ifdef CVF_PROF
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pk_t2));
  if (tid == 0 && blockIdx.x % 21 == 0)
    printf("cvf cta %d: entry %llu ns, after pdl_wait +%llu, exit +%llu, units %lld\n", (int)blockIdx.x, pk_t0 % 100000000ull,
           pk_t1 - pk_t0, pk_t2 - pk_t0, u_end - u_begin);
#endif
  if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

void pack_cvf_tc_weights(const float* w_oidhw, std::vector<uint8_t>* out) {
  out->assign(W_BYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  // per (in-plane tap t2, k-step ks) one 6 KB block: [W_hi, W_lo][k octet (2)][row (96)][8 fp16], row = 32 j + output
  // channel with j = 2 - kz -- the window of output slices (i - 2, i - 1, i) an input slice i contributes to
  for (int tap = 0; tap < 27; ++tap)
    for (int nn = 0; nn < 32; ++nn)
      for (int c = 0; c < 32; ++c) {
        const int kz = tap / 9, t2 = tap % 9, ks = c / 16, k = c % 16;
        const float w = w_oidhw[((size_t)nn * 32 + c) * 27 + tap];
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const size_t e = (size_t)(t2 * 2 + ks) * 3072 + (size_t)(k / 8) * 768 + (size_t)((2 - kz) * 32 + nn) * 8 + (size_t)(k % 8);
        h[e] = hi;
        h[e + 1536] = lo;
      }
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
