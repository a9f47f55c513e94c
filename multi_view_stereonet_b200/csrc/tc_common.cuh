// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace b200mvs {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully active warp.  The MMA-issuing code must sit in warp-uniform control flow
// (`if (warp == 0)` with a provably uniform warp index, then `if (elect_one())`) with descriptors computed from
// warp-uniform values: ptxas then keeps them in uniform registers and emits back-to-back UTCHMMA.  Issuing from
// `if (tid == 0)` makes it wrap every MMA in an ELECT / R2UR / BRA.U.ANY waterfall loop (~70 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE.  Canonical layout (16-byte units):
// ((8, m), 2) : ((1, SBO), LBO) -- a core matrix is 8 rows x 16 bytes stored contiguously, the next
// 8-row group is SBO bytes further, the second 16-byte K chunk of the MMA is LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = F16 (K-major), M = 128.
constexpr uint32_t idesc_f16(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

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

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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

// One lane polls (with a hardware suspend-time hint), the warp then reconverges: hundreds of threads spinning on
// try_wait flood the shared-memory pipe and starve the warps that do the work.
// One try_wait per lane; true if the phase has completed.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if (__all_sync(0xffffffffu, mbar_test(bar, parity))) return;   // fast path: already complete
  if ((threadIdx.x & 31) == 0) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(a), "r"(parity), "r"(2000u)
          : "memory");
    }
  }
  __syncwarp();
  mbar_wait(bar, parity);   // completes at once; gives every lane its own acquire of the phase
}


// Weights into shared memory with the bulk-copy engine: one thread arms `bar` with the byte count and issues 1-D
// cp.async.bulk copies (chunks of at most 32 KB); whoever consumes the weights waits on `bar` (parity 0).  A
// 100 KB weight set staged with LDG/STS costs every thread ~25 dependent vector loads in the prologue.
__device__ __forceinline__ void bulk_load_weights(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  for (uint32_t off = 0; off < bytes; off += 32768u) {
    const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst) + off),
                 "l"(reinterpret_cast<const uint8_t*>(gsrc) + off), "r"(n), "r"(smem_u32(bar))
                 : "memory");
  }
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane.
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

// 8 consecutive fp32 accumulator columns of this thread's TMEM lane.
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

// Host: one 2 KB B-operand block [k half (2)][n (64)][8 fp16], rows 0..31 = W_hi, rows 32..63 = W_lo.
inline void put_split_weight(__half* blocks, int block, int k, int n, float w) {
  const size_t e = (size_t)block * 1024 + (size_t)(k / 8) * 512 + (size_t)n * 8 + (size_t)(k % 8);
  const __half hi = __float2half_rn(w);
  const __half lo = __float2half_rn(w - __half2float(hi));
  blocks[e] = hi;
  blocks[e + 32 * 8] = lo;
}

}  // namespace tc
}  // namespace b200mvs
