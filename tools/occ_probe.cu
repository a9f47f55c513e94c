// Which feature of a kernel makes cudaOccupancyMaxActiveBlocksPerMultiprocessor report one CTA per SM on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/occ_probe tools/occ_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void kA(float* o) { o[threadIdx.x] = 1.f; }
template <int COLS>
__global__ void kB(float* o) {
  __shared__ uint32_t s;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&s)), "r"((uint32_t)COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  o[threadIdx.x] = (float)s;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s), "r"((uint32_t)COLS) : "memory");
}
__global__ void kD(float* o, const float* g) {
  __shared__ __align__(128) float buf[1024];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(4096u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(buf)), "l"(g), "r"(4096u), "r"(s32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
  o[threadIdx.x] = buf[threadIdx.x];
}
__global__ void kE(float* o, unsigned* c) {
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(c, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory"); } while (v < gridDim.x);
  }
  __syncthreads();
  o[threadIdx.x] = 1.f;
}
__global__ void __maxnreg__(120) kF(float* o) { o[threadIdx.x] = 1.f; }
template <class K>
void report(const char* name, K k, int threads, size_t dyn) {
  int n = -1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, dyn);
  cudaFuncAttributes fa{};
  cudaFuncGetAttributes(&fa, k);
  printf("%-28s threads %d dyn %6zu -> %d CTAs/SM (%s) regs %d static %zu\n", name, threads, dyn, n, cudaGetErrorString(e), fa.numRegs, fa.sharedSizeBytes);
}
int main() {
  report("plain", kA, 256, 0);
  report("tcgen05.alloc 64", kB<64>, 256, 0);
  report("tcgen05.alloc 128", kB<128>, 256, 0);
  report("tcgen05.alloc 256", kB<256>, 256, 0);
  report("tcgen05.alloc 512", kB<512>, 256, 0);
  report("mbarrier + bulk copy", kD, 256, 0);
  report("acquire spin", kE, 256, 0);
  report("maxnreg 120", kF, 256, 0);
  cudaFuncSetAttribute(kB<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  report("tcgen05.alloc 128 + 100 KB", kB<128>, 256, 100 * 1024);
  cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  report("plain + 100 KB", kA, 256, 100 * 1024);
  return 0;
}
