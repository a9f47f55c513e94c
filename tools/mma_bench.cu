// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128) as a function of N and of the A operand's
// shared-memory layout / start alignment (no-swizzle K-major planes as the conv kernels use them).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multi_view_stereonet_b200/csrc tools/mma_bench.cu -o tools/bin/mma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace b200mvs;

namespace b200mvs {
void set_error(const std::string&) {}
void note_launch() {}
bool pdl_enabled() { return false; }
void set_pdl_enabled(bool) {}
}


// PATTERN: 0 = same A every MMA; 1 = nine taps (ky*PW + kx) of a PW=42 plane, two K-steps, as the conv kernels issue
// them; 2 = split-precision pairs (N=64 on A_hi, N=32 on A_lo) over the nine taps x two K-steps.
template <int N, int PATTERN, int COUNT, int BASE, bool INTERLEAVED>
__global__ void __launch_bounds__(128, 1) bench_kernel(long long* out, int plane_bytes, int d_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = tc::uniform_warp_index();
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&s_tmem, 512u);
  if (threadIdx.x == 32) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t plane_u16 = (uint32_t)plane_bytes >> 4;
  const uint64_t da0 = INTERLEAVED ? tc::umma_desc(tc::smem_u32(smem), 128u, 256u)
                                   : tc::umma_desc(tc::smem_u32(smem), (uint32_t)plane_bytes, 128u);
  const uint64_t db0 = tc::umma_desc(tc::smem_u32(smem + 160 * 1024), 1024u, 128u);
  if (warp == 0) {
    if (tc::elect_one()) {
#pragma unroll 1
      for (int rep = 0; rep < 5; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < COUNT; ++i) {
          if (PATTERN == 0) {
            tc::mma_f16(tmem, da0 + BASE, db0, tc::idesc_f16(N), i ? 1u : 0u);
          } else {
            const int tap = (i / 2) % 9, ks = i % 2;
            const uint32_t pos = (uint32_t)((tap / 3) * 42 + (tap % 3));
            const uint64_t a = da0 + (uint64_t)(BASE + 2u * ks * plane_u16 + pos);
            const uint64_t b = db0 + (uint64_t)(((tap * 2 + ks) % 18) * 128);
            if (PATTERN == 1) {
              tc::mma_f16(tmem + d_off, a, b, tc::idesc_f16(N), i ? 1u : 0u);
            } else if (PATTERN == 3) {
              // sliding-window Conv3d batch (cvf_tc.cu): A_hi W_hi, A_lo W_hi into X, A_hi W_lo into Y = X + 256 columns,
              // weight blocks of 96 rows (LBO 1536)
              const uint64_t b3 = tc::umma_desc(tc::smem_u32(smem + 64 * 1024), 1536u, 128u) + (uint64_t)((tap * 2 + ks) * 384);
              tc::mma_f16(tmem + d_off, a, b3, tc::idesc_f16(N), i ? 1u : 0u);
              tc::mma_f16(tmem + d_off, a + 4u * plane_u16, b3, tc::idesc_f16(N), 1u);
              tc::mma_f16(tmem + 256 + d_off, a, b3 + 192u, tc::idesc_f16(N), i ? 1u : 0u);
            } else {
              tc::mma_f16(tmem, a, b, tc::idesc_f16(64), i ? 1u : 0u);
              tc::mma_f16(tmem, a + 4u * plane_u16, b, tc::idesc_f16(32), 1u);
            }
          }
        }
        const long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, (uint32_t)(rep & 1));
        const long long t2 = clock64();
        out[rep * 2 + 0] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512u);
}

// ---- M = 64: cost of the split pairs, and where the 64 accumulator rows live in TMEM ----
constexpr uint32_t idesc_m(uint32_t n, uint32_t m) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

template <int M>
__global__ void __launch_bounds__(128, 1) bench_m_kernel(long long* out, float* rows_out, int plane_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = tc::uniform_warp_index();
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // A: row r has the value r + 1 at k = 0 (K-major, no swizzle, SBO = 128, LBO = plane_bytes);  B: every row n has 1 at k = 0
  for (int r = threadIdx.x; r < 128; r += 128)
    *reinterpret_cast<__half*>(smem + (r / 8) * 128 + (r % 8) * 16) = __float2half((float)(r + 1));
  for (int n = threadIdx.x; n < 64; n += 128)
    *reinterpret_cast<__half*>(smem + 160 * 1024 + (n / 8) * 128 + (n % 8) * 16) = __float2half(1.0f);
  if (warp == 0) tc::tmem_alloc(&s_tmem, 512u);
  if (threadIdx.x == 32) {
    tc::mbar_init(&bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t plane_u16 = (uint32_t)plane_bytes >> 4;
  const uint64_t da0 = tc::umma_desc(tc::smem_u32(smem), (uint32_t)plane_bytes, 128u);
  const uint64_t db0 = tc::umma_desc(tc::smem_u32(smem + 160 * 1024), 1024u, 128u);
  if (warp == 0) {
    if (tc::elect_one()) {
      // layout probe: one MMA, rows -> TMEM lanes
      tc::mma_f16(tmem, da0, db0, idesc_m(32, M), 0u);
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, 0u);
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    rows_out[threadIdx.x] = v[0];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    if (tc::elect_one()) {
#pragma unroll 1
      for (int rep = 0; rep < 5; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < 18; ++i) {
          const int tap = (i / 2) % 9, ks = i % 2;
          const uint32_t pos = (uint32_t)((tap / 3) * 42 + (tap % 3));
          const uint64_t a = da0 + (uint64_t)(2u * ks * plane_u16 + pos);
          const uint64_t b = db0 + (uint64_t)(((tap * 2 + ks) % 18) * 128);
          tc::mma_f16(tmem, a, b, idesc_m(64, M), i ? 1u : 0u);
          tc::mma_f16(tmem, a + 4u * plane_u16, b, idesc_m(32, M), 1u);
        }
        const long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, (uint32_t)((rep + 1) & 1));
        const long long t2 = clock64();
        out[rep * 2 + 0] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512u);
}

template <int M>
void run_m(long long* d) {
  auto k = bench_m_kernel<M>;
  float* rows;
  cudaMalloc(&rows, 128 * sizeof(float));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 128, 200 * 1024>>>(d, rows, 218 * 16);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[10];
  float r[128];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(r, rows, sizeof(r), cudaMemcpyDeviceToHost);
  printf("M=%d split pairs N=64+N=32 (36 MMAs): issue %5lld total %6lld cycles = %6.1f / MMA (%s)\n", M, h[8], h[9],
         (double)h[9] / 36, cudaGetErrorString(e));
  printf("M=%d accumulator row (+1) held by TMEM lane 0..127, column 0:\n", M);
  for (int i = 0; i < 128; ++i) printf("%s%3.0f", i % 32 == 0 ? "\n  " : " ", r[i]);
  printf("\n");
  cudaFree(rows);
}

int g_grid = 1;
template <int N, int PATTERN, int COUNT, int BASE, bool INTERLEAVED>
void run(const char* name, long long* d, int d_off = 0) {
  auto k = bench_kernel<N, PATTERN, COUNT, BASE, INTERLEAVED>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<g_grid, 128, 200 * 1024>>>(d, 218 * 16, d_off);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[10];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int mmas = COUNT * (PATTERN == 2 ? 2 : PATTERN == 3 ? 3 : 1);
  printf("%-44s issue %5lld  total %6lld cycles = %6.1f / MMA  (%s)\n", name, h[8], h[9], (double)h[9] / mmas,
         cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 10 * sizeof(long long));
  run_m<128>(d);
  run_m<64>(d);
  run<32, 0, 36, 0, false>("N=32 same A aligned", d);
  run<32, 0, 36, 1, false>("N=32 same A +16 B", d);
  run<32, 0, 36, 4, false>("N=32 same A +64 B", d);
  run<64, 0, 36, 0, false>("N=64 same A aligned", d);
  run<64, 0, 36, 1, false>("N=64 same A +16 B", d);
  run<128, 0, 36, 0, false>("N=128 same A aligned", d);
  run<128, 0, 36, 1, false>("N=128 same A +16 B", d);
  run<256, 0, 36, 0, false>("N=256 same A aligned", d);
  run<256, 0, 36, 1, false>("N=256 same A +16 B", d);
  run<8, 0, 36, 0, false>("N=8 same A aligned", d);
  run<16, 0, 36, 0, false>("N=16 same A aligned", d);
  run<32, 1, 18, 0, false>("N=32 conv taps x2 ksteps (18)", d);
  run<64, 1, 18, 0, false>("N=64 conv taps x2 ksteps (18)", d);
  run<96, 1, 18, 0, false>("N=96 conv taps x2 ksteps (18)", d);
  run<128, 1, 18, 0, false>("N=128 conv taps x2 ksteps (18)", d);
  run<192, 1, 18, 0, false>("N=192 conv taps x2 ksteps (18)", d);
  run<256, 1, 18, 0, false>("N=256 conv taps x2 ksteps (18)", d);
  run<96, 1, 54, 0, false>("N=96 conv taps x2 ksteps (54)", d);
  run<96, 1, 54, 0, false>("N=96 conv taps (54), D at column 32", d, 32);
  run<96, 1, 54, 0, false>("N=96 conv taps (54), D at column 64", d, 64);
  run<96, 1, 54, 0, false>("N=96 conv taps (54), D at column 96", d, 96);
  run<96, 1, 54, 0, false>("N=96 conv taps (54), D at column 128", d, 128);
  run<96, 3, 18, 0, false>("sliding batch N=96 x3 (54), D at 0", d, 0);
  run<96, 3, 18, 0, false>("sliding batch N=96 x3 (54), D at 32", d, 32);
  run<96, 3, 18, 0, false>("sliding batch N=96 x3 (54), D at 64", d, 64);
  run<96, 3, 18, 0, false>("sliding batch N=96 x3 (54), D at 128", d, 128);
  run<64, 3, 18, 0, false>("sliding batch N=64 x3 (54), D at 32", d, 32);
  run<128, 3, 18, 0, false>("sliding batch N=128 x3 (54), D at 0", d, 0);
  run<128, 3, 18, 0, false>("sliding batch N=128 x3 (54), D at 32", d, 32);
  g_grid = 148;
  run<96, 3, 18, 0, false>("148 CTAs: sliding batch N=96 x3 (54)", d, 32);
  run<64, 3, 18, 0, false>("148 CTAs: sliding batch N=64 x3 (54)", d, 32);
  run<32, 2, 54, 0, false>("148 CTAs: split pairs N=64+N=32 (108)", d);
  run<256, 0, 36, 0, false>("148 CTAs: N=256 same A", d);
  g_grid = 1;
  run<32, 2, 18, 0, false>("split pairs N=64+N=32 (36), one conv", d);
  run<32, 2, 54, 0, false>("split pairs N=64+N=32 (108), conv3d slice", d);
  run<32, 0, 36, 0, true>("N=32 interleaved K chunks aligned", d);
  run<32, 0, 36, 1, true>("N=32 interleaved K chunks +16 B", d);
  run<32, 0, 36, 2, true>("N=32 interleaved K chunks +32 B", d);
  run<64, 0, 36, 2, true>("N=64 interleaved K chunks +32 B", d);
  return 0;
}
