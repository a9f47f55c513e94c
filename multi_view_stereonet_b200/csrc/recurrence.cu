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
// The image half of the first convolution, conv3x3(img, W0[:, 0:3]) + b, does not depend on the recurrence: it is
// computed for all hypotheses beforehand (image_conv_kernel, misc.cu) and added in the first epilogue, so the
// tensor core only sees 32-channel operands.
//
// Decomposition.  The 1/16-scale image is linearised with a zero column on each side (pitch
// PW = w + 2); CTA r of the cluster owns output positions [128 r, 128 r + 128) = one UMMA M-tile.
// Each 3x3 conv is 9 taps x 2 k-steps of tcgen05.mma.kind::f16 (M=128, N=32) whose A operand
// for tap (ky, kx) is the staged activation plane started ky*PW + kx positions later (see
// conv_tc.cu).  The 1/16-scale stages need ~fp32 operand accuracy, so every
// operand is split x = hi + lo into two fp16 values and each k-step issues the three products
// (hi*hi + lo*hi + hi*lo) as two MMAs, fp32 accumulation in TMEM: ~22 significant bits.
// GroupNorm needs whole-image statistics twice per step: every CTA pushes its (sum, sumsq) per channel group into
// every CTA's shared memory with st.async, and the raw conv outputs of the PW + 1 positions next to a tile boundary go to
// the neighbour CTA as one bulk shared::cta -> shared::cluster copy; both complete BYTES ON THE RECEIVER'S mbarrier,
// so each CTA waits for its own expected byte count (no cluster-wide barrier) and normalises its tile plus halo
// locally.  The previous hypothesis -- the gather source of the next step -- lives in a position-indexed shared-memory
// window that is filled where the data is produced: own rows by the CTA's last epilogue, side rows by the neighbours'
// bulk copies; global memory only receives the result (and serves taps that leave the window, behind progress flags).
// The gather plan of every step (taps, bilinear weights) depends on the cameras only and is precomputed
// (gather_plan_kernel).  All weights (hi and lo, three layers, 108 KB) stay resident in shared memory for all steps.
// The level-4 tail of the feature network (l4_tail_kernel, below) reuses the same decomposition.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include <vector>

#include "conv.cuh"
#include "recurrence.cuh"

namespace b200mvs {
namespace {

constexpr int NT = 512;            // worker threads: 16 warps = (4 lane quarters of the M-tile) x (4 channel octets)
constexpr bool kMmaWarp = false;   // a 17th warp that only issues the MMAs (caps the kernel at 96 registers: 5 warps on one SMSP)
constexpr int NT_ALL = NT + (kMmaWarp ? 32 : 0);
constexpr int MTILE = 128;
constexpr int W0_BLOCKS = 9 * 2;   // taps x k-steps (32 input channels)
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
// kind::f16 instruction descriptors: D = F32, A = B = F16 (K-major), M = 128, N = 32 / 64.
constexpr uint32_t kIdescN32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescN64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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
  // Window of the previous hypothesis kept in shared memory: output positions [pos0 - wm, pos0 + MTILE + wm), one
  // 128-byte row (fp32 [32]) per position.  wm = halo + one image row + 2: the taps of the own + halo positions under an
  // incremental motion of up to one row / two columns.
  int wm, win_rows;
  uint32_t plane_bytes;
  // byte offsets into dynamic shared memory
  uint32_t off_w, off_planes, off_halo, off_wf, off_stage, total;
};

// Planes (each npl_pad x 16 B): [hi f0..f3][lo f0..f3]
constexpr int PLANE_HI = 0, PLANE_LO = 4, NUM_PLANES = 8;
constexpr int MAX_TASKS = 2;   // (position, channel octet) staging tasks per thread: npl * 4 <= 2 * NT
constexpr uint32_t kSmemBudget = 227u * 1024u - 8u * 1024u;   // dynamic part; static arrays take the rest

__host__ __device__ inline Layout make_layout(int rows, int cols) {
  Layout L;
  L.PW = cols + 2;
  L.halo = L.PW + 1;
  L.npl = MTILE + 2 * L.halo;
  // padded to 2 (mod 8) positions: consecutive planes then start 32 bytes apart modulo 128, so the four octet planes
  // the lanes of a quad write at once fall into distinct shared-memory banks
  L.npl_pad = ((L.npl + 5) & ~7) + 2;
  L.plane_bytes = (uint32_t)L.npl_pad * 16u;
  L.wm = L.halo + L.PW + 2;
  L.win_rows = MTILE + 2 * L.wm;
  uint32_t o = 0;
  L.off_w = o;
  o += W_TOTAL_BYTES;
  L.off_planes = o;
  o += NUM_PLANES * L.plane_bytes;
  L.off_halo = o;          // [layer 2][lower, upper][halo][32] fp32, written by the neighbour CTAs
  o += 2 * 2 * (uint32_t)L.halo * kC * 4;
  L.off_wf = o;            // warped features of the own positions, fp32 [128][32]
  o += MTILE * kC * 4;
  L.off_stage = o;         // the window
  o += (uint32_t)L.win_rows * kC * 4;
  L.total = o;
  return L;
}

// Do the taps of a gather (north-west tap at output position P00, flags as in the plan) lie in the window of CTA r?
__host__ __device__ inline bool taps_in_window(int P00, int fl, int r, const Layout& L) {
  const int w00 = P00 - (r * MTILE - L.wm);
  return w00 >= 0 && w00 + ((fl & 4) ? L.PW : 0) + ((fl >> 1) & 1) < L.win_rows;
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
  const float* imgconv;  // [n][D][rows*cols][32]: image half of conv0 + bias0 (image_conv_kernel)
  const float4* plan;    // [n][D][plan_stride]: gather plan of every step (gather_plan_kernel)
  int plan_stride;
  int* flags;            // [n][17]: [rank] = last hypothesis CTA `rank` has made visible in global memory (only kept up
                         // to date when [16] != 0); [16] != 0: some tap of the sweep lies outside its CTA's window
  int D, rows, cols, n_tiles;
  long long* prof;       // optional [16][12] phase cycle totals (debug)
  int debug;             // timing ablations (wrong results): 1 skip MMAs, 2 skip gathers, 8 no generic->async proxy
                         // fences before the MMAs, 16 no halo exchange
};


// Two 8-column slices of this thread's TMEM lane (the hi*hi+lo*hi and hi*lo accumulators): both loads in flight,
// one wait.
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

// One conv = 9 taps x ksteps k-steps; per k-step two MMAs implement the split product
//   D[:, 0:32]  += A_hi * W_hi      D[:, 32:64] += A_hi * W_lo      (one N=64 MMA on [W_hi | W_lo])
//   D[:, 0:32]  += A_lo * W_hi                                      (one N=32 MMA)
// so the A operand -- the shared-memory-bandwidth-bound side at N=32 -- is read twice, not three times.
// A single thread issues the MMAs, so the issue loop itself is on the critical path: the (A_hi, A_lo, B)
// descriptors of every k-step are the same in every step of the sweep (fixed shared-memory addresses), are built
// once into a shared-memory table, and the loop is three loads and two MMAs per k-step.
// A single elected lane of warp 0 issues the MMAs from warp-uniform control flow with descriptors computed from
// warp-uniform values, so that ptxas keeps them in uniform registers and emits back-to-back UTCHMMA.  (Issuing
// from an `if (tid == 0)` branch, or loading descriptors from memory, makes it wrap every MMA in an
// ELECT / R2UR / BRA.U.ANY waterfall loop: ~70 cycles per MMA, more than the MMA itself.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void issue_conv_mmas(uint64_t da_hi0, uint64_t da_lo0, uint64_t db, uint32_t plane_u16,
                                                uint32_t PW, uint32_t tmem_base) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t pos = (uint32_t)(tap / 3) * PW + (uint32_t)(tap % 3);   // 16-byte units
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t a_off = (uint64_t)(2u * ks * plane_u16 + pos);
      const uint64_t b = db + (uint64_t)((tap * 2 + ks) * (2048 / 16));
      mma_f16(tmem_base, da_hi0 + a_off, b, kIdescN64, (tap | ks) != 0 ? 1u : 0u);
      mma_f16(tmem_base, da_lo0 + a_off, b, kIdescN32, 1u);
    }
  }
}

__device__ __forceinline__ void st_async_f2(uint32_t raddr, float2 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
               "f"(v.x), "f"(v.y), "r"(rbar)
               : "memory");
}
// A CTA's (sum, sumsq) of one group to CTA `dst` of the cluster, 8 bytes counted on that CTA's barrier.  An image of a
// single tile runs as a "cluster" of one CTA, for which shared::cluster addressing is not defined: it stores locally
// and completes the bytes on its own barrier (release: fence, then the relaxed complete_tx).
__device__ __forceinline__ void push_partial(float2* slot, uint64_t* bar, float2 v, uint32_t dst, bool single) {
  if (single) {
    *slot = v;
    __threadfence_block();
    asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(8u) : "memory");
  } else {
    st_async_f2(map_to_rank(smem_u32(slot), dst), v, map_to_rank(smem_u32(bar), dst));
  }
}
// Rows of 32 floats (128 bytes) whose eight 16-byte chunks are XOR-swizzled with the row index: a warp that reads or
// writes the same chunk of 32 consecutive rows (thread = position, one channel octet) touches every bank group
// instead of one (a 32-way conflict in the linear layout).
__device__ __forceinline__ float4* swz_ptr(float* base, int row, int chunk) {
  return reinterpret_cast<float4*>(base + (size_t)row * 32 + (size_t)((chunk ^ (row & 7)) << 2));
}
__device__ __forceinline__ const float4* swz_ptr(const float* base, int row, int chunk) {
  return reinterpret_cast<const float4*>(base + (size_t)row * 32 + (size_t)((chunk ^ (row & 7)) << 2));
}
__device__ __forceinline__ void mbar_arm_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Default (.acquire.cta) wait: enough for bytes that st.async / cp.async.bulk completed on this CTA's own barrier (the
// data is in shared memory); an .acquire.cluster wait makes ptxas add CCTL.IVALL (an L1 invalidation) to every poll.
__device__ __forceinline__ void mbar_wait_cta(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

// shared::cta -> another CTA's shared memory, bytes counted on an mbarrier of the destination CTA
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Per-step schedule of one CTA (one 128-position M-tile of one view):
//   W    wait for the neighbours' rows of the previous hypothesis (bulk copies into the window's side rows), gather
//        own + halo positions from the window (plan loaded during the previous step; taps outside the window: global
//        memory) -> split-fp16 conv0 operand
//   MMA0 | E0: raw output + image half -> (sum, sumsq) per group to every CTA (st.async), boundary rows to the two
//        neighbours (one bulk copy each), each completing bytes on the receiver's mbarriers
//   wait for the statistics | S1: coefficients, normalise own rows, wait for the boundary rows, normalise the halo
//        -> operand | MMA1 | E1 / wait / S2 | MMA2
//   E2   features = warped + delta -> own rows of the window; first / last wm rows -> the neighbours' windows (bulk
//        copies); the tile -> global memory as coalesced 16-byte chunks; next step's plan loads
// No cluster-wide barrier inside the loop: every wait is on bytes that land in this CTA's own shared memory.
template <bool PROF>
__global__ void __launch_bounds__(NT_ALL, 1) recurrence_kernel(const RecParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // GroupNorm partial (sum, sumsq) of [layer][group][source CTA], pushed by the source CTA; unused slots stay zero
  __shared__ __align__(16) float2 s_part[2][kGroups][16];
  __shared__ __align__(16) float2 s_loc[kGroups][4];   // this CTA's warp partials [group][warp quarter]
  __shared__ __align__(16) float s_bias[2][kC], s_gamma[2][kC], s_beta[2][kC];   // biases of conv1, conv2 (conv0's is in imgconv)
  __shared__ __align__(8) uint64_t s_bar;        // MMA completion
  __shared__ __align__(8) uint64_t s_sbar[2];    // per layer: GroupNorm partials pushed into this CTA by the cluster
  __shared__ __align__(8) uint64_t s_hbar[2];    // per layer: boundary rows copied into this CTA by its neighbours
  __shared__ __align__(8) uint64_t s_tbar;       // bytes the neighbours copied into this CTA's window
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int n = blockIdx.y;
  const Layout L = make_layout(p.rows, p.cols);
  const int PW = L.PW, halo = L.halo, npl = L.npl;
  const int pixels = p.rows * p.cols;
  const bool active = (int)rank < p.n_tiles;
  const int pos0 = (int)rank * MTILE;  // first own output position; also the input position of local l = 0

  uint8_t* s_w = smem + L.off_w;
  uint8_t* s_planes = smem + L.off_planes;
  float* s_halo = reinterpret_cast<float*>(smem + L.off_halo);
  float* s_wf = reinterpret_cast<float*>(smem + L.off_wf);
  uint8_t* const s_win = smem + L.off_stage;   // window rows [pos0 - wm, pos0 + MTILE + wm), see Layout

  // ---- one-time setup ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_sbar[0])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_sbar[1])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_hbar[0])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_hbar[1])), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_tbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 2 * kGroups * 16) (&s_part[0][0][0])[tid] = make_float2(0.f, 0.f);
  if (tid < kC) {
    s_bias[0][tid] = __ldg(p.bias1 + tid);
    s_bias[1][tid] = __ldg(p.bias2 + tid);
    s_gamma[0][tid] = __ldg(p.gamma0 + tid);
    s_beta[0][tid] = __ldg(p.beta0 + tid);
    s_gamma[1][tid] = __ldg(p.gamma1 + tid);
    s_beta[1][tid] = __ldg(p.beta1 + tid);
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w16);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < W_TOTAL_BYTES / 16; i += NT_ALL) dst[i] = __ldg(src + i);
    // zero every plane once: padding lanes stay zero
    uint4* pl = reinterpret_cast<uint4*>(s_planes);
    for (int i = tid; i < NUM_PLANES * L.npl_pad; i += NT_ALL) pl[i] = make_uint4(0, 0, 0, 0);
    float4* hz = reinterpret_cast<float4*>(s_halo);
    for (int i = tid; i < 2 * 2 * halo * kC / 4; i += NT_ALL) hz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  uint32_t bar_phase = 0;
  pdl_launch_dependents();
  cluster_sync_all();  // every CTA's shared memory and mbarriers are initialised before anyone pushes into them

  const float inv_count = 1.0f / (8.0f * (float)pixels);
  const uint32_t plane_u16 = L.plane_bytes >> 4;
  const uint64_t da_hi0 = umma_desc(smem_u32(s_planes) + (uint32_t)PLANE_HI * L.plane_bytes, L.plane_bytes, 128u);
  const uint64_t da_lo0 = umma_desc(smem_u32(s_planes) + (uint32_t)PLANE_LO * L.plane_bytes, L.plane_bytes, 128u);
  const uint64_t db_c0 = umma_desc(smem_u32(s_w), 1024u, 128u);

  // Tensor-core conv.  Warp 16 does nothing but issue: it sleeps at named barrier 7 until the 512 workers have staged
  // an operand (they only ARRIVE there and go on to wait for the MMAs' completion barrier), issues the 36 MMAs of the
  // conv from one elected lane and commits them to s_bar.  (tcgen05.mma issue blocks at the rate the tensor pipe
  // drains; issued from a worker warp, that warp entered every epilogue ~400 cycles behind the other fifteen, and
  // the cluster waited for it.)
  auto issue_conv = [&](int layer) {   // warp-uniform caller
    if (active && elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (!(p.debug & 1))
        issue_conv_mmas(da_hi0, da_lo0, db_c0 + (uint64_t)(layer * W1_BLOCKS * (2048 / 16)), plane_u16, (uint32_t)PW,
                        tmem_base);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&s_bar))
                   : "memory");
    }
    __syncwarp();
  };
  if (kMmaWarp && warp == 16) {
    for (int step = 1; step < p.D; ++step) {
#pragma unroll 1
      for (int layer = 0; layer < 3; ++layer) {
        asm volatile("bar.sync 7, %0;" ::"n"(NT_ALL) : "memory");
        issue_conv(layer);
      }
    }
    cluster_sync_all();   // the final cluster barrier counts every thread
    return;   // (TMEM is freed by warp 0 behind a barrier of the worker warps only)
  }
  // workers: operand staged (generic-proxy writes fenced by the caller) -> wake the MMA warp, do not wait
  // (without the extra warp: warp 0 waits for the others' arrival and issues; nobody else waits)
  auto operands_ready = [&](int layer) {
    if (!kMmaWarp && warp == 0) {
      asm volatile("bar.sync 7, %0;" ::"n"(NT_ALL) : "memory");
      issue_conv(layer);
    } else {
      asm volatile("bar.arrive 7, %0;" ::"n"(NT_ALL) : "memory");
    }
  };
  // every thread waits on the completion barrier itself (a hardware sleep, ~60 cycles from arrive to wake-up): no
  // block-wide barrier between the tensor core and the epilogue
  auto wait_conv = [&]() {
    if (active) mbar_wait_cta(&s_bar, bar_phase);
    bar_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  // This thread's slice of the accumulator tile: one output position, ONE channel octet (= one GroupNorm group).
  const int wq = warp & 3, oct_e = warp >> 2;
  const int jl = wq * 32 + lane;                 // local output position
  const int jg = pos0 + jl;                      // global output position
  const int oy = jg / PW, ox = jg % PW;
  const bool real_out = active && ox < p.cols && oy < p.rows;
  const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(oct_e * 8);
  const int own_l = jl + halo;                   // local input position of the own output
  const size_t own_pix = real_out ? (size_t)oy * p.cols + ox : 0;

  // Gather tasks of this thread, fixed for all steps: (local input position l, channel octet).  The four lanes of
  // a quad share the position and take one octet each.
  const int t_oct = tid & 3;  // NT % 4 == 0: the octet is the same for every task of a thread
  int t_l[MAX_TASKS];
#pragma unroll
  for (int k = 0; k < MAX_TASKS; ++k) t_l[k] = (tid + k * NT) >> 2;
  // halo staging task: halo position (tid & 127) of the thread's own octet; 2 * halo <= 128
  const int h_idx = tid & 127;                              // lower halo then upper halo, as laid out in s_halo
  const bool h_in = active && h_idx < 2 * halo;
  const int h_l = h_idx < halo ? h_idx : h_idx + MTILE;     // local input position
  bool h_real;
  {
    const int Lg = pos0 + h_l;
    const int gy = Lg / PW - 1, gx = Lg % PW - 1;
    h_real = h_in && gy >= 0 && gy < p.rows && gx >= 0 && gx < p.cols;
  }
  // bytes the cluster pushes into this CTA per layer: every active CTA's (sum, sumsq) of the four groups and the
  // boundary rows of the two neighbours
  const uint32_t sbytes = (uint32_t)p.n_tiles * (uint32_t)kGroups * 8u;
  const uint32_t hbytes = (rank > 0 && !(p.debug & 16) ? (uint32_t)halo * kC * 4u : 0u) +
                          ((int)rank + 1 < p.n_tiles && !(p.debug & 16) ? (uint32_t)halo * kC * 4u : 0u);
  // Window of the previous hypothesis (Layout): filled where the data is produced -- the own rows by this CTA's last
  // epilogue, the wm rows on either side by the two neighbours' bulk copies.  A window row is 128 bytes whose eight
  // 16-byte chunks are XOR-swizzled with the (global) output position, so rows copy verbatim between CTAs, the last
  // epilogue (thread = position) writes without bank conflicts and the gather (quad = position) reads without.
  const int Wm = L.wm;
  const int win_lo = pos0 - Wm;   // output position of window row 0
  const bool has_prev = active && rank > 0, has_next = active && (int)rank + 1 < p.n_tiles;
  const uint32_t wbytes = ((has_prev ? 1u : 0u) + (has_next ? 1u : 0u)) * (uint32_t)Wm * kC * 4u;
  int* const gflags = p.flags + (size_t)n * 17;
  // The real pixels of the own tile are one contiguous pixel range of the image (rows of the tile follow each other in
  // memory).  The last epilogue leaves the new features in shared memory and the tile goes out as 16-byte chunks
  // in pixel order: 32 lanes = 512 contiguous bytes.  (One thread = one position stored 32 bytes at a 128-byte
  // stride: 32 sectors per warp instruction, ~1 k cycles of store issue per step.)
  int co_src[2], co_dst[2];   // byte offset into the window (swizzled) / float offset into the hypothesis; -1: nothing
  {
    int first = pos0 % PW < p.cols ? pos0 : (pos0 / PW + 1) * PW;
    int last = pos0 + MTILE - 1 < p.rows * PW - 1 ? pos0 + MTILE - 1 : p.rows * PW - 1;
    if (last % PW >= p.cols) last = (last / PW) * PW + p.cols - 1;
    const bool any = active && first <= last;
    const int px_lo = any ? (first / PW) * p.cols + first % PW : 0;
    const int px_hi = any ? (last / PW) * p.cols + last % PW : -1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * NT, px = px_lo + (i >> 3), chunk = i & 7;
      co_src[k] = co_dst[k] = -1;
      if (px <= px_hi) {
        const int row = (px / p.cols) * PW + px % p.cols - pos0;   // local output position
        co_src[k] = (Wm + row) * 128 + ((chunk ^ (row & 7)) << 4);
        co_dst[k] = px * kC + chunk * 4;
      }
    }
  }

  auto plane_ptr = [&](int plane, int l) -> uint4* {
    return reinterpret_cast<uint4*>(s_planes + (size_t)plane * L.plane_bytes + (size_t)l * 16);
  };

  long long t_prev = PROF ? clock64() : 0;
  long long acc_t[PROF ? 12 : 1];
#pragma unroll
  for (int k = 0; k < (PROF ? 12 : 1); ++k) acc_t[k] = 0;
#define PROF_MARK(k)                          \
  do {                                        \
    if (PROF) {                               \
      const long long _t = clock64();         \
      acc_t[k] += _t - t_prev;                \
      t_prev = _t;                            \
    }                                         \
  } while (0)

  // Per-warp timeline of one step (PROF only): clock64 of lane 0 of every warp at ~30 points, [rank][warp][32]
  // behind the [16][12] phase totals.  All CTAs leave the cluster barrier within a few cycles of each other, so
  // point 0 (taken right behind it) is the common origin across CTAs.
  long long* const trace = (PROF && p.prof != nullptr && blockIdx.y == 0) ? p.prof + 16 * 12 + ((size_t)rank * 16 + warp) * 32 : nullptr;
  const int trace_step = p.D / 2;
  int cur_step = 0;
#define TRACE(k)                                                                   \
  do {                                                                             \
    if (PROF && trace != nullptr && cur_step == trace_step && lane == 0) trace[k] = clock64(); \
  } while (0)

  pdl_wait();  // everything above ran under the previous kernel's tail; from here on its outputs are read

  // Gather plan of the coming step: one float4 per (step, padded input position) -- first tap's pixel index, east and
  // south weights, flags -- precomputed for all steps by gather_plan_kernel (it depends on the homographies only).
  // Loaded one step ahead, under conv0's MMAs.
  const float4* const plan_base = p.plan + (size_t)n * p.D * p.plan_stride + pos0;
  float4 g_plan[MAX_TASKS];
  auto load_plan = [&](int step) {
#pragma unroll
    for (int k = 0; k < MAX_TASKS; ++k) {
      g_plan[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && t_l[k] < npl && !(p.debug & 2)) g_plan[k] = __ldg(plan_base + (size_t)step * p.plan_stride + t_l[k]);
    }
  };
  // hypothesis 0 (written by the feature network): the whole window from global memory, once
  if (active) {
    const float4* src0 = reinterpret_cast<const float4*>(p.vol_in + (size_t)n * p.D * pixels * kC);
    for (int i = tid; i < L.win_rows * 8; i += NT) {
      const int w = i >> 3, chunk = i & 7, P = win_lo + w;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (P >= 0) {
        const int y = P / PW, x = P - y * PW;
        if (y < p.rows && x < p.cols) v = __ldg(src0 + ((size_t)y * p.cols + x) * 8 + chunk);
      }
      *reinterpret_cast<float4*>(s_win + (size_t)w * 128 + ((chunk ^ (P & 7)) << 4)) = v;
    }
  }
  const bool slow = __ldg(gflags + 16) != 0;   // some tap lies outside its window: global memory + progress flags
  asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");
  load_plan(1);

  float x0own[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x0own[k] = 0.f;

  for (int step = 1; step < p.D; ++step) {
    cur_step = step;
    TRACE(0);
    if (active && tid == 0) {
      mbar_arm_tx(&s_sbar[0], sbytes);
      mbar_arm_tx(&s_sbar[1], sbytes);
      if (hbytes != 0) {
        mbar_arm_tx(&s_hbar[0], hbytes);
        mbar_arm_tx(&s_hbar[1], hbytes);
      }
      if (step >= 2 && wbytes != 0) mbar_arm_tx(&s_tbar, wbytes);   // (bytes that are already here count ahead)
    }
    // ================= W: warp previous features into the conv0 operand ============================
    if (active && step >= 2 && wbytes != 0) mbar_wait_cta(&s_tbar, (uint32_t)(step & 1));
    TRACE(1);
    {
      const float* prev = p.vol_in + ((size_t)n * p.D + (step - 1)) * pixels * kC + 8 * t_oct;
#pragma unroll
      for (int k = 0; k < MAX_TASKS; ++k) {
        if (active && t_l[k] < npl) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          const int fl = __float_as_int(g_plan[k].w);
          if (fl & 1) {
            const float we = g_plan[k].y, ws = g_plan[k].z;
            const float ww = 1.0f - we, wn = 1.0f - ws;
            const float wt[4] = {wn * ww, wn * we, ws * ww, ws * we};
            const int P00 = __float_as_int(g_plan[k].x);   // output position of the north-west tap
            const int dx = (fl >> 1) & 1, dyP = (fl & 4) ? PW : 0;
            const int offP[4] = {0, dx, dyP, dyP + dx};
            if (taps_in_window(P00, fl, (int)rank, L)) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int Pt = P00 + offP[t];
                const uint8_t* rowp = s_win + (size_t)(Pt - win_lo) * 128;
                const uint32_t ca = (uint32_t)(((2 * t_oct) ^ (Pt & 7)) << 4);
                const float4 a = *reinterpret_cast<const float4*>(rowp + ca);
                const float4 b = *reinterpret_cast<const float4*>(rowp + (ca ^ 16u));
                v[0] = fmaf(a.x, wt[t], v[0]); v[1] = fmaf(a.y, wt[t], v[1]); v[2] = fmaf(a.z, wt[t], v[2]); v[3] = fmaf(a.w, wt[t], v[3]);
                v[4] = fmaf(b.x, wt[t], v[4]); v[5] = fmaf(b.y, wt[t], v[5]); v[6] = fmaf(b.z, wt[t], v[6]); v[7] = fmaf(b.w, wt[t], v[7]);
              }
            } else {
              // a tap outside the window (incremental motion beyond one row): global memory, once the CTA that owns the
              // pixel has published the previous hypothesis (gather_plan_kernel found such taps: `slow` is set and
              // every CTA keeps its progress flag up to date)
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int Pt = P00 + offP[t];
                if (step >= 2) {
                  int seen;
                  do {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(gflags + Pt / MTILE) : "memory");
                  } while (seen < step - 1);
                }
                const int y = Pt / PW, x = Pt - y * PW;
                const float4* gp = reinterpret_cast<const float4*>(prev + ((size_t)y * p.cols + x) * kC);
                const float4 a = __ldcg(gp), b = __ldcg(gp + 1);
                v[0] = fmaf(a.x, wt[t], v[0]); v[1] = fmaf(a.y, wt[t], v[1]); v[2] = fmaf(a.z, wt[t], v[2]); v[3] = fmaf(a.w, wt[t], v[3]);
                v[4] = fmaf(b.x, wt[t], v[4]); v[5] = fmaf(b.y, wt[t], v[5]); v[6] = fmaf(b.z, wt[t], v[6]); v[7] = fmaf(b.w, wt[t], v[7]);
              }
            }
          }
          const int l = t_l[k];
          if (l >= halo && l < halo + MTILE) {
            const int jo = l - halo;
            *swz_ptr(s_wf, jo, 2 * t_oct) = make_float4(v[0], v[1], v[2], v[3]);
            *swz_ptr(s_wf, jo, 2 * t_oct + 1) = make_float4(v[4], v[5], v[6], v[7]);
          }
          uint4 hi, lo;
          split8(v, &hi, &lo);
          *plane_ptr(PLANE_HI + t_oct, l) = hi;
          *plane_ptr(PLANE_LO + t_oct, l) = lo;
        }
      }
    }
    TRACE(2);
    if (!(p.debug & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    operands_ready(0);
    PROF_MARK(0);
    TRACE(3);
    // image half of conv0 (+ bias) at this thread's output position, consumed in the first epilogue: in flight under
    // conv0's MMAs
    float4 ic0 = make_float4(0.f, 0.f, 0.f, 0.f), ic1 = ic0;
    if (real_out) {
      const float4* icp =
          reinterpret_cast<const float4*>(p.imgconv + (((size_t)n * p.D + step) * pixels + own_pix) * kC + oct_e * 8);
      ic0 = __ldg(icp);
      ic1 = __ldg(icp + 1);
    }
    TRACE(4);
    wait_conv();
    PROF_MARK(1);
    TRACE(5);

    // ===== two normalised layers: raw output -> statistics + halo exchange -> wait -> operand -> conv =====
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      // ---- epilogue: raw output (+bias) stays in registers, is pushed to the neighbours' halos, and is reduced
      float y[8];
      if (active) {
        float c[8];
        tmem_ld8x2(tmem_my, tmem_my + 32u, y, c);
        TRACE(6 + 8 * layer);
        if (layer == 0) {
          const float add[8] = {ic0.x, ic0.y, ic0.z, ic0.w, ic1.x, ic1.y, ic1.z, ic1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = (y[k] + c[k]) + add[k];
        } else {
          const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[0][oct_e * 8]);
          const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[0][oct_e * 8 + 4]);
          const float add[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = (y[k] + c[k]) + add[k];
        }
        // ---- statistics first: every CTA of the cluster waits for every other CTA's push ----
        float gs = 0.f, gq = 0.f;
        if (real_out) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            gs += y[k];
            gq += y[k] * y[k];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs += __shfl_xor_sync(0xffffffffu, gs, o);
          gq += __shfl_xor_sync(0xffffffffu, gq, o);
        }
        // The four warps of a group add up inside the CTA (fixed order), then lane d of the group's first warp pushes
        // the CTA's (sum, sumsq) to CTA d: 4 * n_tiles remote stores per CTA instead of one per warp and value (a
        // remote store costs ~2 cycles of issue whatever its size; 352 of them were ~700 cycles of every epilogue).
        if (layer == 0) TRACE(28);
        if (lane == 0) s_loc[oct_e][wq] = make_float2(gs, gq);
        if (wq != 0) {
          asm volatile("bar.arrive %0, 128;" ::"r"(1 + oct_e) : "memory");
        } else {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + oct_e) : "memory");
          if (lane < p.n_tiles) {
            const float4 u = *reinterpret_cast<const float4*>(&s_loc[oct_e][0]);
            const float4 w = *reinterpret_cast<const float4*>(&s_loc[oct_e][2]);
            push_partial(&s_part[layer][oct_e][rank], &s_sbar[layer],
                         make_float2((u.x + u.z) + (w.x + w.z), (u.y + u.w) + (w.y + w.w)), (uint32_t)lane, p.n_tiles == 1);
          }
        }
        if (layer == 0) TRACE(29);
        // Halo rows of a layer at the receiver: [2 * halo][32] fp32, lower halo then upper halo, 16-byte chunks
        // XOR-swizzled with the row index (the receiver reads one octet of 32 consecutive rows per instruction).
        // The boundary rows are written into a local image of the receiver's rows -- in the staging buffer of the
        // previous hypothesis, which is dead between the gather and the next step's bulk load -- and go out as ONE
        // bulk copy per neighbour (shared::cta -> shared::cluster, bytes counted on the receiver's mbarrier): a remote
        // store costs ~2 cycles of issue whatever its size, and 2 x 344 of them were ~1.2 k cycles of every epilogue.
        // (rows for rank-1 in the window's lower side rows, rows for rank+1 in its upper side rows: dead from the end of
        //  the gather until that neighbour's last epilogue of this step, by which time it has received them)
        float* hout_prev = reinterpret_cast<float*>(s_win) + (size_t)layer * halo * kC;
        float* hout_next = reinterpret_cast<float*>(s_win) + (size_t)(Wm + MTILE) * kC + (size_t)layer * halo * kC;
        if (jl < halo && rank > 0) {  // -> upper halo of rank-1, its row halo + jl
          const int key = (halo + jl) & 7;
          float* row = hout_prev + (size_t)jl * kC;
          *reinterpret_cast<float4*>(row + (((2 * oct_e) ^ key) << 2)) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(row + (((2 * oct_e + 1) ^ key) << 2)) = make_float4(y[4], y[5], y[6], y[7]);
        }
        if (jl >= MTILE - halo && (int)rank + 1 < p.n_tiles) {  // -> lower halo of rank+1, its row idx
          const int idx = jl - (MTILE - halo);
          *swz_ptr(hout_next, idx, 2 * oct_e) = make_float4(y[0], y[1], y[2], y[3]);
          *swz_ptr(hout_next, idx, 2 * oct_e + 1) = make_float4(y[4], y[5], y[6], y[7]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // these rows -> readable by the copy engine
        if (layer == 0) TRACE(26);
        if (wq < 2) {          // positions [0, 64): the writers of the rows for rank-1 (8 warps)
          if (rank > 0) {
            if (warp == 0) {
              asm volatile("bar.sync 5, 256;" ::: "memory");
              if (!(p.debug & 16) && elect_one()) {
                float* hb = s_halo + ((size_t)layer * 2 + 1) * halo * kC;
                dsmem_bulk_copy(map_to_rank(smem_u32(hb), rank - 1), smem_u32(hout_prev), (uint32_t)halo * kC * 4u,
                                map_to_rank(smem_u32(&s_hbar[layer]), rank - 1));
              }
              __syncwarp();
            } else {
              asm volatile("bar.arrive 5, 256;" ::: "memory");
            }
          }
        } else {               // positions [64, 128): the writers of the rows for rank+1
          if ((int)rank + 1 < p.n_tiles) {
            if (warp == 2) {
              asm volatile("bar.sync 6, 256;" ::: "memory");
              if (!(p.debug & 16) && elect_one()) {
                float* hb = s_halo + (size_t)layer * 2 * halo * kC;
                dsmem_bulk_copy(map_to_rank(smem_u32(hb), rank + 1), smem_u32(hout_next),
                                (uint32_t)halo * kC * 4u, map_to_rank(smem_u32(&s_hbar[layer]), rank + 1));
              }
              __syncwarp();
            } else {
              asm volatile("bar.arrive 6, 256;" ::: "memory");
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      PROF_MARK(2 + 4 * layer);
      TRACE(7 + 8 * layer);
      if (active) mbar_wait_cta(&s_sbar[layer], (uint32_t)((step - 1) & 1));
      // (single-tile image: the partials were stored locally; the block barrier adds nothing the mbarrier has not
      //  ordered already, but it is the synchronisation compute-sanitizer's racecheck can see)
      if (p.n_tiles == 1) asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");
      PROF_MARK(3 + 4 * layer);
      TRACE(8 + 8 * layer);

      if (active) {
        // ---- GroupNorm coefficients of the own octet's group: lane r sums CTA r's four warp partials, then a
        //      butterfly over the lanes (every lane ends with the same bits: deterministic) ----
        float ca[8], cb[8];
        {
          // every thread adds the 16 CTA slots of its group in one fixed tree (broadcast loads, unused slots are
          // zero; the same bits in every thread of the cluster)
          float ts, tq;
          {
            const float4* sp = reinterpret_cast<const float4*>(&s_part[layer][oct_e][0]);
            float4 q[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = sp[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              q[i].x += q[i].z;
              q[i].y += q[i].w;
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
#pragma unroll
              for (int i = 0; i < 8; i += 2 * o) {
                q[i].x += q[i + o].x;
                q[i].y += q[i + o].y;
              }
            ts = q[0].x;
            tq = q[0].y;
          }
          const double mean = (double)ts * (double)inv_count;
          const double var = (double)tq * (double)inv_count - mean * mean;  // cancellation in double
          const float rstd = rsqrtf(fmaxf((float)var, 0.f) + kGnEps);
          const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[layer][8 * oct_e]);
          const float4 g1 = *reinterpret_cast<const float4*>(&s_gamma[layer][8 * oct_e + 4]);
          const float4 e0 = *reinterpret_cast<const float4*>(&s_beta[layer][8 * oct_e]);
          const float4 e1 = *reinterpret_cast<const float4*>(&s_beta[layer][8 * oct_e + 4]);
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bt[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            ca[k] = gm[k] * rstd;
            cb[k] = bt[k] - (float)mean * ca[k];
          }
        }
        TRACE(9 + 8 * layer);
        // ---- next operand: x = lrelu(GN(y)) (+ x0 for the residual block) over own + halo positions ----
        {
          float x[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float t = real_out ? lrelu(fmaf(y[k], ca[k], cb[k])) : 0.f;
            x[k] = (layer == 0) ? t : t + x0own[k];
            if (layer == 0) x0own[k] = t;
          }
          uint4 hi, lo;
          split8(x, &hi, &lo);
          *plane_ptr(PLANE_HI + oct_e, own_l) = hi;
          *plane_ptr(PLANE_LO + oct_e, own_l) = lo;
        }
        TRACE(10 + 8 * layer);
        // the neighbours' boundary rows have had the coefficient and own-row work to arrive
        if (hbytes != 0) mbar_wait_cta(&s_hbar[layer], (uint32_t)((step - 1) & 1));
        if (h_in) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          uint4* ph = plane_ptr(PLANE_HI + oct_e, h_l);
          uint4* plo = plane_ptr(PLANE_LO + oct_e, h_l);
          if (h_real) {
            const float* hb = s_halo + (size_t)layer * 2 * halo * kC;
            const float4 a = *swz_ptr(hb, h_idx, 2 * oct_e), b = *swz_ptr(hb, h_idx, 2 * oct_e + 1);
            const float yy[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            float xprev[8];
            if (layer == 1) unsplit8(*ph, *plo, xprev);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              v[e] = lrelu(fmaf(yy[e], ca[e], cb[e]));
              if (layer == 1) v[e] += xprev[e];
            }
          }
          uint4 hi, lo;
          split8(v, &hi, &lo);
          *ph = hi;
          *plo = lo;
        }
      }
      TRACE(11 + 8 * layer);
      if (!(p.debug & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      operands_ready(1 + layer);
      PROF_MARK(4 + 4 * layer);
      TRACE(12 + 8 * layer);
      wait_conv();
      PROF_MARK(5 + 4 * layer);
      TRACE(13 + 8 * layer);
    }

    // ===== E2: features_step = wf + delta -> own rows of the window; boundary rows -> the neighbours' windows
    //           (their gather source of the next step); the tile -> global memory in pixel order =====
    const bool more = step + 1 < p.D;
    if (active) {
      float v[8], c[8];
      tmem_ld8x2(tmem_my, tmem_my + 32u, v, c);
      TRACE(22);
      if (real_out) {
        uint8_t* rowp = s_win + (size_t)(Wm + jl) * 128;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 wv = *swz_ptr(s_wf, jl, 2 * oct_e + q);
          const float4 bv = *reinterpret_cast<const float4*>(&s_bias[1][oct_e * 8 + 4 * q]);
          float4 r;
          r.x = wv.x + ((v[4 * q + 0] + c[4 * q + 0]) + bv.x);
          r.y = wv.y + ((v[4 * q + 1] + c[4 * q + 1]) + bv.y);
          r.z = wv.z + ((v[4 * q + 2] + c[4 * q + 2]) + bv.z);
          r.w = wv.w + ((v[4 * q + 3] + c[4 * q + 3]) + bv.w);
          *reinterpret_cast<float4*>(rowp + (((2 * oct_e + q) ^ (jl & 7)) << 4)) = r;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (more) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // own rows -> readable by the copy engine
    asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");   // the 512 workers
    // The first / last wm own rows are the upper / lower side rows of the neighbours' windows.  They are past this
    // step's gather (they sent the statistics this CTA waited for after it), so their side rows are free.
    if (more && warp == 1) {
      if (has_prev && elect_one())
        dsmem_bulk_copy(map_to_rank(smem_u32(s_win + (size_t)(Wm + MTILE) * 128), rank - 1), smem_u32(s_win + (size_t)Wm * 128),
                        (uint32_t)Wm * 128u, map_to_rank(smem_u32(&s_tbar), rank - 1));
      __syncwarp();
    }
    if (more && warp == 2) {
      if (has_next && elect_one())
        dsmem_bulk_copy(map_to_rank(smem_u32(s_win), rank + 1), smem_u32(s_win + (size_t)MTILE * 128), (uint32_t)Wm * 128u,
                        map_to_rank(smem_u32(&s_tbar), rank + 1));
      __syncwarp();
    }
    TRACE(24);
    {
      float* dst = p.vol + ((size_t)n * p.D + step) * pixels * kC;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (co_src[k] >= 0) __stcg(reinterpret_cast<float4*>(dst + co_dst[k]), *reinterpret_cast<const float4*>(s_win + co_src[k]));
    }
    if (slow && more) {   // out-of-window taps of other CTAs read this tile from global memory
      asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");
      if (active && tid == 32) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(gflags + rank), "r"(step) : "memory");
    }
    // the NEXT step's gather plan, in flight while the neighbours' rows arrive
    if (more) load_plan(step + 1);
    PROF_MARK(10);
    TRACE(23);
    PROF_MARK(11);
    TRACE(25);
  }
  cluster_sync_all();   // no CTA leaves while a peer may still read from or write into its shared memory
  if (PROF && p.prof != nullptr && tid == 0 && blockIdx.y == 0) {
    for (int k = 0; k < 12; ++k) p.prof[rank * 12 + k] = acc_t[PROF ? k : 0];
  }

  asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");   // the workers (the MMA warp has left)
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// Gather plan of the sweep: for every step and every padded input position Lg (image pixel (Lg / PW - 1, Lg % PW - 1);
// CTA r gathers positions [128 r, 128 r + 128 + 2 halo)), where the incremental homography H_{d-1}^-1 H_d
// (multi_view_stereonet.py:280-282) sends the pixel: {output position y0 * PW + x0 of the north-west tap, east weight,
// south weight, flags (1 valid, 2 east tap exists, 4 south tap exists)}; flags[n][16] is raised when a tap lies
// outside the shared-memory window of a CTA that gathers it (the sweep then keeps progress flags in global memory).  It depends on the cameras only, so it is computed for all
// steps next to the sweep's other inputs instead of inside its dependent chain (there it was ~250 instructions per
// thread and step, longer than the MMAs it was hidden under).
__global__ void __launch_bounds__(256) gather_plan_kernel(const float* __restrict__ Hinc, int D, int rows, int cols,
                                                          int plan_stride, int n_tiles, float4* __restrict__ plan,
                                                          int* __restrict__ flags) {
  pdl_wait();
  pdl_launch_dependents();
  const int Lg = blockIdx.x * blockDim.x + threadIdx.x;
  const int step = blockIdx.y + 1, n = blockIdx.z;
  if (Lg >= plan_stride) return;
  float H[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = __ldg(Hinc + ((size_t)n * D + step) * 9 + i);
  const Layout L = make_layout(rows, cols);
  const int PW = L.PW;
  const int gy = Lg / PW - 1, gx = Lg % PW - 1;
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
    const WarpCoord c = homography_coord(H, (float)gx, (float)gy, rows, cols);
    if (!c.invalid) {
      const float fx0 = floorf(c.ix), fy0 = floorf(c.iy);   // bilinear_setup (common.cuh), kept as (tap, weights)
      const int x0 = (int)fx0, y0 = (int)fy0;
      int fl = 1;
      if (x0 + 1 <= cols - 1) fl |= 2;   // otherwise the east tap is clamped onto x0 (and carries weight 0)
      if (y0 + 1 <= rows - 1) fl |= 4;
      const int P00 = y0 * PW + x0;
      out.x = __int_as_float(P00);
      out.y = c.ix - fx0;
      out.z = c.iy - fy0;
      out.w = __int_as_float(fl);
      // every CTA that gathers this position (as an own or a halo position) must find the taps in its window
      int r_lo = (Lg - L.npl + MTILE) / MTILE, r_hi = Lg / MTILE;
      r_lo = r_lo < 0 ? 0 : r_lo;
      r_hi = r_hi > n_tiles - 1 ? n_tiles - 1 : r_hi;
      bool inside = true;
      for (int r = r_lo; r <= r_hi; ++r) inside = inside && taps_in_window(P00, fl, r, L);
      if (!inside) atomicOr(flags + (size_t)n * 17 + 16, 1);
    }
  }
  plan[((size_t)n * D + step) * plan_stride + Lg] = out;
}


// =================================================================================================================
// The level-4 tail of FeatureNetwork -- six residual blocks x_{i+1} = lrelu(GN(conv_i(x_i))) + x_i and conv_final
// (multi_view_stereonet.py:96-105, 119-127; utils/resnet.py:93-109) -- as ONE cluster kernel per image with the
// activations, the GroupNorm statistics and the tile boundaries exchanged on chip, exactly as in the sweep above:
// seven 3x3 32->32 layers on a 32 x 40 image are seven ~3 us layers here against seven ~8 us kernels (each of those
// is statistics round trip + staging + 36 MMAs + epilogue + store flush + launch boundary).  The weights of a layer
// (36 KB, split fp16, the same 2 KB blocks as conv_tc.cu uses) stream through two buffers one layer ahead.
// =================================================================================================================
constexpr int TAIL_LAYERS = 7;
constexpr uint32_t TAIL_W_BYTES = W1_BLOCKS * 2u * 1024u;   // one layer: 18 blocks of [W_hi | W_lo]

struct TailLayout {
  uint32_t off_w, off_planes, off_halo, off_hout, total;
};
__host__ __device__ inline TailLayout make_tail_layout(const Layout& L) {
  TailLayout T;
  uint32_t o = 0;
  T.off_w = o;
  o += 2 * TAIL_W_BYTES;
  T.off_planes = o;
  o += NUM_PLANES * L.plane_bytes;
  T.off_halo = o;          // [layer parity 2][lower, upper][halo][32] fp32, written by the neighbour CTAs
  o += 2 * 2 * (uint32_t)L.halo * kC * 4;
  T.off_hout = o;          // [layer parity 2][to prev, to next][halo][32] fp32, the bulk copies' sources
  o += 2 * 2 * (uint32_t)L.halo * kC * 4;
  T.total = o;
  return T;
}

struct TailParams {
  const float* x0;         // [images][rows*cols][32] output of FeatureNetwork.conv3
  float* out;              // image i at out + i * out_stride, [rows*cols][32]
  long long out_stride;
  const uint8_t* w[TAIL_LAYERS];
  const float* bias[TAIL_LAYERS];
  const float* gamma[TAIL_LAYERS - 1];
  const float* beta[TAIL_LAYERS - 1];
  int rows, cols, n_tiles;
};

__global__ void __launch_bounds__(NT, 1) l4_tail_kernel(const TailParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) float2 s_part[2][kGroups][16];
  __shared__ __align__(16) float2 s_loc[kGroups][4];
  __shared__ __align__(16) float s_bias[TAIL_LAYERS][kC], s_gamma[TAIL_LAYERS - 1][kC], s_beta[TAIL_LAYERS - 1][kC];
  __shared__ __align__(8) uint64_t s_bar, s_sbar[2], s_hbar[2], s_wbar[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t rank = cluster_ctarank();
  const int img = blockIdx.y;
  const Layout L = make_layout(p.rows, p.cols);
  const TailLayout T = make_tail_layout(L);
  const int PW = L.PW, halo = L.halo;
  const int pixels = p.rows * p.cols;
  const bool active = (int)rank < p.n_tiles;
  const int pos0 = (int)rank * MTILE;
  uint8_t* s_w = smem + T.off_w;
  uint8_t* s_planes = smem + T.off_planes;
  float* s_halo = reinterpret_cast<float*>(smem + T.off_halo);
  float* s_hout = reinterpret_cast<float*>(smem + T.off_hout);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint64_t* b : {&s_bar, &s_sbar[0], &s_sbar[1], &s_hbar[0], &s_hbar[1], &s_wbar[0], &s_wbar[1]})
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 2 * kGroups * 16) (&s_part[0][0][0])[tid] = make_float2(0.f, 0.f);
  for (int i = tid; i < TAIL_LAYERS * kC; i += NT) {
    const int l = i / kC, c = i % kC;
    s_bias[l][c] = p.bias[l] != nullptr ? __ldg(p.bias[l] + c) : 0.f;
    if (l < TAIL_LAYERS - 1) {
      s_gamma[l][c] = __ldg(p.gamma[l] + c);
      s_beta[l][c] = __ldg(p.beta[l] + c);
    }
  }
  {
    uint4* pl = reinterpret_cast<uint4*>(s_planes);
    for (int i = tid; i < NUM_PLANES * L.npl_pad; i += NT) pl[i] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  uint32_t bar_phase = 0;
  pdl_launch_dependents();
  // weights are parameters, not outputs of an earlier kernel: layer 0's go in flight before the dependency wait
  auto load_weights = [&](int layer) {
    if (warp == 1) {
      if (active && elect_one()) {
        uint64_t* bar = &s_wbar[layer & 1];
        mbar_arm_tx(bar, TAIL_W_BYTES);
        uint8_t* dst = s_w + (size_t)(layer & 1) * TAIL_W_BYTES;
        tma_load_1d(dst, p.w[layer], 32768u, bar);
        tma_load_1d(dst + 32768u, p.w[layer] + 32768u, TAIL_W_BYTES - 32768u, bar);
      }
      __syncwarp();
    }
  };
  load_weights(0);
  cluster_sync_all();

  const float inv_count = 1.0f / (8.0f * (float)pixels);
  const uint32_t plane_u16 = L.plane_bytes >> 4;
  const uint64_t da_hi0 = umma_desc(smem_u32(s_planes) + (uint32_t)PLANE_HI * L.plane_bytes, L.plane_bytes, 128u);
  const uint64_t da_lo0 = umma_desc(smem_u32(s_planes) + (uint32_t)PLANE_LO * L.plane_bytes, L.plane_bytes, 128u);
  const uint64_t db0 = umma_desc(smem_u32(s_w), 1024u, 128u);
  auto operands_ready = [&](int layer) {   // warp 0 waits for everyone's operand rows and the weights, then issues
    if (warp == 0) {
      asm volatile("bar.sync 7, %0;" ::"n"(NT) : "memory");
      if (active && elect_one()) {
        mbar_wait_cta(&s_wbar[layer & 1], (uint32_t)((layer >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_conv_mmas(da_hi0, da_lo0, db0 + (uint64_t)((layer & 1) * (TAIL_W_BYTES / 16)), plane_u16, (uint32_t)PW,
                        tmem_base);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&s_bar))
                     : "memory");
      }
      __syncwarp();
    } else {
      asm volatile("bar.arrive 7, %0;" ::"n"(NT) : "memory");
    }
  };
  auto wait_conv = [&]() {
    if (active) mbar_wait_cta(&s_bar, bar_phase);
    bar_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  auto plane_ptr = [&](int plane, int l) -> uint4* {
    return reinterpret_cast<uint4*>(s_planes + (size_t)plane * L.plane_bytes + (size_t)l * 16);
  };

  // thread = (output position, channel octet) as in the sweep
  const int wq = warp & 3, oct_e = warp >> 2;
  const int jl = wq * 32 + lane, jg = pos0 + jl;
  const int oy = jg / PW, ox = jg % PW;
  const bool real_out = active && ox < p.cols && oy < p.rows;
  const uint32_t tmem_my = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(oct_e * 8);
  const int own_l = jl + halo;
  const size_t own_pix = real_out ? (size_t)oy * p.cols + ox : 0;
  const int h_idx = tid & 127;
  const bool h_in = active && h_idx < 2 * halo;
  const int h_l = h_idx < halo ? h_idx : h_idx + MTILE;
  bool h_real;
  size_t h_pix = 0;
  {
    const int Lg = pos0 + h_l;
    const int gy = Lg / PW - 1, gx = Lg % PW - 1;
    h_real = h_in && gy >= 0 && gy < p.rows && gx >= 0 && gx < p.cols;
    if (h_real) h_pix = (size_t)gy * p.cols + gx;
  }
  const bool has_prev = active && rank > 0, has_next = active && (int)rank + 1 < p.n_tiles;
  const uint32_t sbytes = (uint32_t)p.n_tiles * (uint32_t)kGroups * 8u;
  const uint32_t hbytes = ((has_prev ? 1u : 0u) + (has_next ? 1u : 0u)) * (uint32_t)halo * kC * 4u;

  pdl_wait();
  // ---- x_0 (raw conv3 output) of the own and halo positions: every position is in global memory ----
  float xres[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) xres[k] = 0.f;
  if (active) {
    const float* xin = p.x0 + (size_t)img * pixels * kC + oct_e * 8;
    if (real_out) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xin + own_pix * kC));
      const float4 b = __ldg(reinterpret_cast<const float4*>(xin + own_pix * kC) + 1);
      xres[0] = a.x; xres[1] = a.y; xres[2] = a.z; xres[3] = a.w;
      xres[4] = b.x; xres[5] = b.y; xres[6] = b.z; xres[7] = b.w;
    }
    uint4 hi, lo;
    split8(xres, &hi, &lo);
    *plane_ptr(PLANE_HI + oct_e, own_l) = hi;
    *plane_ptr(PLANE_LO + oct_e, own_l) = lo;
    if (h_in) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
      if (h_real) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xin + h_pix * kC));
        const float4 b = __ldg(reinterpret_cast<const float4*>(xin + h_pix * kC) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
      split8(v, &hi, &lo);
      *plane_ptr(PLANE_HI + oct_e, h_l) = hi;
      *plane_ptr(PLANE_LO + oct_e, h_l) = lo;
    }
  }

#pragma unroll 1
  for (int layer = 0; layer < TAIL_LAYERS; ++layer) {
    const int pb = layer & 1;
    const uint32_t par = (uint32_t)((layer >> 1) & 1);
    const bool last = layer == TAIL_LAYERS - 1;
    if (active && tid == 0 && !last) {
      mbar_arm_tx(&s_sbar[pb], sbytes);
      if (hbytes != 0) mbar_arm_tx(&s_hbar[pb], hbytes);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    operands_ready(layer);
    // the other weight buffer was last read by layer - 1's MMAs, which everyone has seen complete
    if (!last) load_weights(layer + 1);
    wait_conv();

    float y[8];
    if (active) {
      float c[8];
      tmem_ld8x2(tmem_my, tmem_my + 32u, y, c);
      const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[layer][oct_e * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&s_bias[layer][oct_e * 8 + 4]);
      const float add[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = (y[k] + c[k]) + add[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (last) {   // conv_final: no normalisation, no activation
      if (real_out) {
        float* dst = p.out + (size_t)img * p.out_stride + own_pix * kC + oct_e * 8;
        __stcg(reinterpret_cast<float4*>(dst), make_float4(y[0], y[1], y[2], y[3]));
        __stcg(reinterpret_cast<float4*>(dst) + 1, make_float4(y[4], y[5], y[6], y[7]));
      }
      break;
    }
    if (active) {
      // ---- statistics to every CTA, boundary rows to the neighbours (see the sweep's epilogue) ----
      float gs = 0.f, gq = 0.f;
      if (real_out) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gs += y[k];
          gq += y[k] * y[k];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, o);
        gq += __shfl_xor_sync(0xffffffffu, gq, o);
      }
      if (lane == 0) s_loc[oct_e][wq] = make_float2(gs, gq);
      if (wq != 0) {
        asm volatile("bar.arrive %0, 128;" ::"r"(1 + oct_e) : "memory");
      } else {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + oct_e) : "memory");
        if (lane < p.n_tiles) {
          const float4 u = *reinterpret_cast<const float4*>(&s_loc[oct_e][0]);
          const float4 w = *reinterpret_cast<const float4*>(&s_loc[oct_e][2]);
          push_partial(&s_part[pb][oct_e][rank], &s_sbar[pb],
                       make_float2((u.x + u.z) + (w.x + w.z), (u.y + u.w) + (w.y + w.w)), (uint32_t)lane, p.n_tiles == 1);
        }
      }
      float* hout_prev = s_hout + (size_t)(pb * 2) * halo * kC;
      float* hout_next = s_hout + (size_t)(pb * 2 + 1) * halo * kC;
      if (jl < halo && has_prev) {
        const int key = (halo + jl) & 7;
        float* row = hout_prev + (size_t)jl * kC;
        *reinterpret_cast<float4*>(row + (((2 * oct_e) ^ key) << 2)) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(row + (((2 * oct_e + 1) ^ key) << 2)) = make_float4(y[4], y[5], y[6], y[7]);
      }
      if (jl >= MTILE - halo && has_next) {
        const int idx = jl - (MTILE - halo);
        *swz_ptr(hout_next, idx, 2 * oct_e) = make_float4(y[0], y[1], y[2], y[3]);
        *swz_ptr(hout_next, idx, 2 * oct_e + 1) = make_float4(y[4], y[5], y[6], y[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (wq < 2) {
        if (has_prev) {
          if (warp == 0) {
            asm volatile("bar.sync 5, 256;" ::: "memory");
            if (elect_one())
              dsmem_bulk_copy(map_to_rank(smem_u32(s_halo + (size_t)(pb * 2 + 1) * halo * kC), rank - 1), smem_u32(hout_prev),
                              (uint32_t)halo * kC * 4u, map_to_rank(smem_u32(&s_hbar[pb]), rank - 1));
            __syncwarp();
          } else {
            asm volatile("bar.arrive 5, 256;" ::: "memory");
          }
        }
      } else {
        if (has_next) {
          if (warp == 2) {
            asm volatile("bar.sync 6, 256;" ::: "memory");
            if (elect_one())
              dsmem_bulk_copy(map_to_rank(smem_u32(s_halo + (size_t)(pb * 2) * halo * kC), rank + 1), smem_u32(hout_next),
                              (uint32_t)halo * kC * 4u, map_to_rank(smem_u32(&s_hbar[pb]), rank + 1));
            __syncwarp();
          } else {
            asm volatile("bar.arrive 6, 256;" ::: "memory");
          }
        }
      }
      mbar_wait_cta(&s_sbar[pb], par);
      if (p.n_tiles == 1) asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory");   // (for racecheck, see the sweep)

      // ---- GroupNorm coefficients, then x_{i+1} = lrelu(GN(y)) + x_i over the own and halo positions ----
      float ca[8], cb[8];
      {
        const float4* sp = reinterpret_cast<const float4*>(&s_part[pb][oct_e][0]);
        float4 q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = sp[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          q[i].x += q[i].z;
          q[i].y += q[i].w;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1)
#pragma unroll
          for (int i = 0; i < 8; i += 2 * o) {
            q[i].x += q[i + o].x;
            q[i].y += q[i + o].y;
          }
        const double mean = (double)q[0].x * (double)inv_count;
        const double var = (double)q[0].y * (double)inv_count - mean * mean;
        const float rstd = rsqrtf(fmaxf((float)var, 0.f) + kGnEps);
        const float4 g0 = *reinterpret_cast<const float4*>(&s_gamma[layer][8 * oct_e]);
        const float4 g1 = *reinterpret_cast<const float4*>(&s_gamma[layer][8 * oct_e + 4]);
        const float4 e0 = *reinterpret_cast<const float4*>(&s_beta[layer][8 * oct_e]);
        const float4 e1 = *reinterpret_cast<const float4*>(&s_beta[layer][8 * oct_e + 4]);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bt[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          ca[k] = gm[k] * rstd;
          cb[k] = bt[k] - (float)mean * ca[k];
        }
      }
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) xres[k] = real_out ? lrelu(fmaf(y[k], ca[k], cb[k])) + xres[k] : 0.f;
        uint4 hi, lo;
        split8(xres, &hi, &lo);
        *plane_ptr(PLANE_HI + oct_e, own_l) = hi;
        *plane_ptr(PLANE_LO + oct_e, own_l) = lo;
      }
      if (hbytes != 0) mbar_wait_cta(&s_hbar[pb], par);
      if (h_in) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        uint4* ph = plane_ptr(PLANE_HI + oct_e, h_l);
        uint4* plo = plane_ptr(PLANE_LO + oct_e, h_l);
        if (h_real) {
          const float* hb = s_halo + (size_t)pb * 2 * halo * kC;
          const float4 a = *swz_ptr(hb, h_idx, 2 * oct_e), b = *swz_ptr(hb, h_idx, 2 * oct_e + 1);
          const float yy[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          float xprev[8];
          unsplit8(*ph, *plo, xprev);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = lrelu(fmaf(yy[e], ca[e], cb[e])) + xprev[e];
        }
        uint4 hi, lo;
        split8(v, &hi, &lo);
        *ph = hi;
        *plo = lo;
      }
    }
  }
  cluster_sync_all();   // no CTA leaves while a peer may still read from or write into its shared memory
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace

int recurrence_plan_stride(int rows, int cols) {
  const Layout L = make_layout(rows, cols);
  return cdiv(rows * L.PW, MTILE) * MTILE + 2 * L.halo;
}

int launch_gather_plan(const float* Hinc, int n, int D, int rows, int cols, float* plan, int* flags,
                       cudaStream_t stream) {
  if (D < 2 || n <= 0) return 0;
  const Layout L = make_layout(rows, cols);
  const int stride = recurrence_plan_stride(rows, cols);
  if (cudaMemsetAsync(flags, 0, (size_t)n * 17 * sizeof(int), stream) != cudaSuccess) {
    set_error("launch_gather_plan: memset failed");
    return -1;
  }
  launch_pdl(gather_plan_kernel, dim3(cdiv(stride, 256), D - 1, n), dim3(256), (size_t)0, stream, Hinc, D, rows, cols,
             stride, cdiv(rows * L.PW, MTILE), reinterpret_cast<float4*>(plan), flags);
  B200MVS_LAUNCH_OK("gather_plan_kernel");
  return 0;
}

// Weight blocks of 2 KB = one N=64 B operand [k half (2)][n (64)][8 fp16] (UMMA K-major, no swizzle) whose
// rows 0..31 hold W_hi and rows 32..63 hold W_lo; an N=32 descriptor on the same block reads W_hi only.
// Order: conv0 (feature channels only, reference input channels 3..34) [tap][kstep 2], conv1, conv2.
void pack_recurrence_weights(const float* w0_oihw35, const float* w1_oihw32, const float* w2_oihw32,
                             std::vector<uint8_t>* out) {
  out->assign(W_TOTAL_BYTES, 0);
  __half* h = reinterpret_cast<__half*>(out->data());
  auto put = [&](int block, int k, int nn, float w) {
    const size_t e = (size_t)block * 1024 + (size_t)(k / 8) * 512 + (size_t)nn * 8 + (size_t)(k % 8);  // in halves
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    h[e] = hi;
    h[e + 32 * 8] = lo;
  };
  for (int tap = 0; tap < 9; ++tap)
    for (int nn = 0; nn < 32; ++nn)
      for (int c = 0; c < 32; ++c) {
        // reference conv0 input order: [image 0..2, features 0..31]  (multi_view_stereonet.py:425)
        put(tap * 2 + c / 16, c % 16, nn, w0_oihw35[((size_t)nn * 35 + 3 + c) * 9 + tap]);
        put(W0_BLOCKS + tap * 2 + c / 16, c % 16, nn, w1_oihw32[((size_t)nn * 32 + c) * 9 + tap]);
        put(W0_BLOCKS + W1_BLOCKS + tap * 2 + c / 16, c % 16, nn, w2_oihw32[((size_t)nn * 32 + c) * 9 + tap]);
      }
}

bool recurrence_supported(int rows, int cols, int* n_tiles, size_t* smem_bytes) {
  const Layout L = make_layout(rows, cols);
  const int tiles = cdiv(rows * L.PW, MTILE);
  if (n_tiles != nullptr) *n_tiles = tiles;
  if (smem_bytes != nullptr) *smem_bytes = L.total;
  return tiles <= 16 && L.halo <= MTILE && L.npl * 4 <= MAX_TASKS * NT && 2 * L.halo <= 128 &&
         L.wm <= MTILE && 2 * L.halo <= L.wm && L.total <= kSmemBudget;   // (window side rows also stage 2 halo blocks)
}

int recurrence_max_clusters(int rows, int cols) {
  int n_tiles = 0;
  size_t smem = 0;
  if (!recurrence_supported(rows, cols, &n_tiles, &smem)) return 0;
  const int cached = cached_cluster_size(1000 + n_tiles);   // (device, key) -> value store of api.cu
  if (cached != 0) return cached;
  for (const void* f : {reinterpret_cast<const void*>(&recurrence_kernel<false>)}) {
    if (ensure_func_smem(f, smem) != 0 || ensure_func_nonportable_cluster(f) != 0) return 0;
  }
  int best = 0;
  const int candidates[3] = {n_tiles, (n_tiles + 1) & ~1, 16};
  for (int c = 0; c < 3 && best == 0; ++c) {
    const int cs = candidates[c];
    if (cs < n_tiles || cs > 16) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, 64, 1);
    cfg.blockDim = dim3(NT_ALL, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int mc = 0;
    if (cudaOccupancyMaxActiveClusters(&mc, recurrence_kernel<false>, &cfg) == cudaSuccess && mc >= 1) best = mc;
    cudaGetLastError();
  }
  if (best > 0) remember_cluster_size(1000 + n_tiles, best);
  return best;
}

bool l4_tail_supported(int rows, int cols) {
  int n_tiles = 0;
  if (!recurrence_supported(rows, cols, &n_tiles, nullptr)) return false;
  const Layout L = make_layout(rows, cols);
  return make_tail_layout(L).total <= kSmemBudget;
}

int launch_l4_tail(const L4TailArgs& a, cudaStream_t stream) {
  if (a.n <= 0) return 0;
  int n_tiles = 0;
  if (!l4_tail_supported(a.rows, a.cols) || !recurrence_supported(a.rows, a.cols, &n_tiles, nullptr)) {
    set_error("launch_l4_tail: shape not supported");
    return -1;
  }
  const Layout L = make_layout(a.rows, a.cols);
  const size_t smem = make_tail_layout(L).total;
  const void* f = reinterpret_cast<const void*>(&l4_tail_kernel);
  if (int rc = ensure_func_smem(f, smem)) return rc;
  if (int rc = ensure_func_nonportable_cluster(f)) return rc;
  TailParams p;
  p.x0 = a.x0;
  p.out = a.out;
  p.out_stride = a.out_stride;
  for (int i = 0; i < TAIL_LAYERS; ++i) {
    p.w[i] = a.w[i];
    p.bias[i] = a.bias[i];
    if (i < TAIL_LAYERS - 1) {
      p.gamma[i] = a.gamma[i];
      p.beta[i] = a.beta[i];
    }
  }
  p.rows = a.rows;
  p.cols = a.cols;
  p.n_tiles = n_tiles;
  const int known = cached_cluster_size(2000 + n_tiles);
  int candidates[3] = {n_tiles, (n_tiles + 1) & ~1, 16};
  if (known != 0) candidates[0] = candidates[1] = candidates[2] = known;
  cudaError_t e = cudaErrorUnknown;
  for (int c = 0; c < 3; ++c) {
    const int cs = candidates[c];
    if (cs < n_tiles || cs > 16) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, a.n, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (known == 0) {
      int max_clusters = 0;
      e = cudaOccupancyMaxActiveClusters(&max_clusters, l4_tail_kernel, &cfg);
      if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        e = cudaErrorLaunchOutOfResources;
        continue;
      }
    }
    e = cudaLaunchKernelEx(&cfg, l4_tail_kernel, p);
    if (e == cudaSuccess) {
      if (known == 0) remember_cluster_size(2000 + n_tiles, cs);
      break;
    }
    cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error(std::string("l4_tail_kernel launch: ") + cudaGetErrorString(e));
    return -2;
  }
  note_launch();
  return 0;
}

int launch_recurrence(const RecurrenceArgs& a, cudaStream_t stream) {
  int n_tiles = 0;
  size_t smem = 0;
  if (!recurrence_supported(a.rows, a.cols, &n_tiles, &smem)) {
    set_error("launch_recurrence: shape not supported by the persistent kernel");
    return -1;
  }
  for (const void* f : {reinterpret_cast<const void*>(&recurrence_kernel<false>),
                        reinterpret_cast<const void*>(&recurrence_kernel<true>)}) {
    if (int rc = ensure_func_smem(f, smem)) return rc;
    if (int rc = ensure_func_nonportable_cluster(f)) return rc;
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
  p.imgconv = a.imgconv;
  p.plan = reinterpret_cast<const float4*>(a.plan);
  p.plan_stride = recurrence_plan_stride(a.rows, a.cols);
  p.flags = a.flags;
  p.D = a.D;
  p.rows = a.rows;
  p.cols = a.cols;
  p.n_tiles = n_tiles;
  p.prof = a.prof;
  p.debug = a.debug;

  // Cluster size: one CTA per M-tile; if that size cannot be scheduled, pad with idle CTAs.
  const int known_cluster = cached_cluster_size(n_tiles);
  int candidates[3] = {n_tiles, (n_tiles + 1) & ~1, 16};
  if (known_cluster != 0) candidates[0] = candidates[1] = candidates[2] = known_cluster;
  cudaError_t e = cudaErrorUnknown;
  for (int c = 0; c < 3; ++c) {
    const int cs = candidates[c];
    if (cs < n_tiles || cs > 16) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs, a.n, 1);
    cfg.blockDim = dim3(NT_ALL, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    static const bool occ_dbg = getenv("B200MVS_REC_OCC") != nullptr;
    if (occ_dbg) {
      for (int sz = 8; sz <= 16; ++sz) {
        cudaLaunchConfig_t c2 = cfg;
        cudaLaunchAttribute a2[1];
        a2[0].id = cudaLaunchAttributeClusterDimension;
        a2[0].val.clusterDim.x = sz;
        a2[0].val.clusterDim.y = 1;
        a2[0].val.clusterDim.z = 1;
        c2.attrs = a2;
        c2.numAttrs = 1;
        c2.gridDim = dim3(sz, a.n, 1);
        int mc = -1;
        cudaError_t ee = cudaOccupancyMaxActiveClusters(&mc, recurrence_kernel<false>, &c2);
        fprintf(stderr, "recurrence occupancy: cluster %d smem %zu -> max active clusters %d (%s)\n", sz, smem, mc,
                cudaGetErrorString(ee));
      }
      cudaGetLastError();
    }
    if (known_cluster == 0) {
      int max_clusters = 0;
      e = cudaOccupancyMaxActiveClusters(&max_clusters, recurrence_kernel<false>, &cfg);
      if (e != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        e = cudaErrorLaunchOutOfResources;
        continue;
      }
    }
    e = (a.prof != nullptr) ? cudaLaunchKernelEx(&cfg, recurrence_kernel<true>, p)
                            : cudaLaunchKernelEx(&cfg, recurrence_kernel<false>, p);
    if (e == cudaSuccess) {
      if (known_cluster == 0) remember_cluster_size(n_tiles, cs);
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
