// Post-processing after the hot path: inverse depth -> metric depth with the baseline
// un-normalisation of test.py:211-214, the ground-truth range mask of test.py:167-186, 218-236 and the KITTI-style
// depth metrics of test.py:41-71, as ONE pass over the estimate and the ground truth.
//
// The reference does this per image on the host: D2H of the estimate, numpy boolean indexing (two compacted
// copies), seven full-array numpy reductions.  Here every pixel is read once; each CTA keeps eight float64
// partial sums, the last CTA of an image adds the partials in CTA order (deterministic) and finalises.
#include "evalpost.cuh"

namespace b200mvs {
namespace {

constexpr int kThreads = 256;
constexpr int kSums = 8;  // abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, count

struct PixelTerms {
  float v[kSums];
};

// One pixel, float32 operation by operation as numpy evaluates test.py:48-61 on float32 arrays.
__device__ __forceinline__ void accumulate(float est_in, float gt_in, float baseline, bool est_is_depth,
                                           float min_depth, float max_depth, float* idepth_out, float* depth_out,
                                           double* acc) {
  float idepth = est_in, depth = est_in;
  if (!est_is_depth) {
    // test.py:211-213: idepth / baseline, then 1 / x where x > 0 (zeros stay zero)
    idepth = __fdiv_rn(est_in, baseline);
    depth = idepth > 0.0f ? __fdiv_rn(1.0f, idepth) : idepth;
  }
  if (idepth_out != nullptr) *idepth_out = idepth;
  if (depth_out != nullptr) *depth_out = depth;
  if (acc == nullptr) return;
  // test.py:170-173: the unpacked ground truth is multiplied back by the baseline
  const float gt = __fmul_rn(gt_in, baseline);
  const bool valid = gt > min_depth && gt < max_depth && depth > min_depth && depth < max_depth;  // test.py:222, 236
  if (!valid) return;
  const float thresh = fmaxf(__fdiv_rn(gt, depth), __fdiv_rn(depth, gt));
  const float diff = __fsub_rn(gt, depth);
  const float sq = __fmul_rn(diff, diff);
  // float32 logarithms as numpy takes them (both implementations are accurate to 1 ulp, neither is correctly
  // rounded; the float64 route costs 4x the kernel time for no gain in agreement)
  const float dl = __fsub_rn(logf(gt), logf(depth));
  acc[0] += (double)__fdiv_rn(fabsf(diff), gt);
  acc[1] += (double)__fdiv_rn(sq, gt);
  acc[2] += (double)sq;
  acc[3] += (double)__fmul_rn(dl, dl);
  acc[4] += thresh < 1.25f ? 1.0 : 0.0;
  acc[5] += thresh < 1.5625f ? 1.0 : 0.0;     // 1.25 ** 2
  acc[6] += thresh < 1.953125f ? 1.0 : 0.0;   // 1.25 ** 3
  acc[7] += 1.0;
}

__global__ void __launch_bounds__(kThreads)
depth_metrics_kernel(const float* __restrict__ est, const float* __restrict__ baseline,
                     const float* __restrict__ depth_true, int est_is_depth, float min_depth, float max_depth,
                     long long pixels, float* __restrict__ idepth_out, float* __restrict__ depth_out,
                     double* __restrict__ metrics, double* __restrict__ partials, unsigned int* __restrict__ counters) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const float bl = baseline != nullptr ? __ldg(baseline + b) : 1.0f;
  const float* e = est + (size_t)b * pixels;
  const float* g = depth_true != nullptr ? depth_true + (size_t)b * pixels : nullptr;
  float* io = idepth_out != nullptr ? idepth_out + (size_t)b * pixels : nullptr;
  float* dp = depth_out != nullptr ? depth_out + (size_t)b * pixels : nullptr;
  double acc[kSums];
#pragma unroll
  for (int i = 0; i < kSums; ++i) acc[i] = 0.0;
  double* accp = g != nullptr ? acc : nullptr;

  // 16-byte path when every per-image plane starts 16-byte aligned (pixels % 4 == 0 and aligned bases)
  const bool vec = (pixels % 4 == 0) && ((reinterpret_cast<uintptr_t>(est) & 15) == 0) &&
                   (depth_true == nullptr || (reinterpret_cast<uintptr_t>(depth_true) & 15) == 0) &&
                   (idepth_out == nullptr || (reinterpret_cast<uintptr_t>(idepth_out) & 15) == 0) &&
                   (depth_out == nullptr || (reinterpret_cast<uintptr_t>(depth_out) & 15) == 0);
  if (vec) {
    const long long quads = pixels / 4;
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < quads; q += (long long)gridDim.x * kThreads) {
      const float4 ev = __ldg(reinterpret_cast<const float4*>(e) + q);
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g != nullptr) gv = __ldg(reinterpret_cast<const float4*>(g) + q);
      float4 iv, dv;
      accumulate(ev.x, gv.x, bl, est_is_depth != 0, min_depth, max_depth, &iv.x, &dv.x, accp);
      accumulate(ev.y, gv.y, bl, est_is_depth != 0, min_depth, max_depth, &iv.y, &dv.y, accp);
      accumulate(ev.z, gv.z, bl, est_is_depth != 0, min_depth, max_depth, &iv.z, &dv.z, accp);
      accumulate(ev.w, gv.w, bl, est_is_depth != 0, min_depth, max_depth, &iv.w, &dv.w, accp);
      if (io != nullptr) reinterpret_cast<float4*>(io)[q] = iv;
      if (dp != nullptr) reinterpret_cast<float4*>(dp)[q] = dv;
    }
  } else {
    for (long long p = (long long)blockIdx.x * kThreads + threadIdx.x; p < pixels; p += (long long)gridDim.x * kThreads)
      accumulate(__ldg(e + p), g != nullptr ? __ldg(g + p) : 0.0f, bl, est_is_depth != 0, min_depth, max_depth,
                 io != nullptr ? io + p : nullptr, dp != nullptr ? dp + p : nullptr, accp);
  }
  if (metrics == nullptr || g == nullptr) return;

  // CTA reduction: warp shuffles, then one partial row per CTA
  __shared__ double sh[kThreads / 32][kSums];
  __shared__ bool is_last;
#pragma unroll
  for (int i = 0; i < kSums; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kSums) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += sh[w][threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * kSums + threadIdx.x] = v;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counters + b, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) counters[b] = 0;   // ready for the next call
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < kSums) {
    double v = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) v += __ldcg(partials + ((size_t)b * gridDim.x + k) * kSums + threadIdx.x);
    sh[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < kSums) {
    const double n = sh[0][7];
    double v = sh[0][threadIdx.x];
    if (threadIdx.x < 7) {
      v = v / n;                                               // mean over the masked pixels (NaN when none)
      if (threadIdx.x == 2 || threadIdx.x == 3) v = sqrt(v);   // rmse, rmse_log
    }
    metrics[(size_t)b * kSums + threadIdx.x] = v;
  }
}

}  // namespace

int depth_metrics_ctas(long long pixels) {
  const long long per_cta = (long long)kThreads * 4 * 4;   // four 16-byte loads per thread
  long long n = (pixels + per_cta - 1) / per_cta;
  if (n < 1) n = 1;
  if (n > 296) n = 296;                                   // two CTAs per SM
  return (int)n;
}

// Per-device scratch for the per-CTA partial sums and the per-image arrival counters, grown on demand and kept
// (a stream-ordered cudaMallocAsync / cudaFreeAsync pair per call cost ~160 us against a ~15 us kernel).  The
// counters are zeroed when the buffer is created; the kernel leaves them at zero.  Calls on one device must not
// overlap on different streams (one process per GPU, one evaluation loop).
struct EvalScratch {
  char* base = nullptr;
  size_t capacity = 0;
};
static EvalScratch g_scratch[64];

int launch_depth_metrics(const float* est, const float* baseline, const float* depth_true, bool est_is_depth,
                         float min_depth, float max_depth, int batch, long long pixels, float* idepth_out,
                         float* depth_out, double* metrics, cudaStream_t stream) {
  if (batch <= 0 || pixels <= 0) return 0;
  const int ctas = depth_metrics_ctas(pixels);
  double* partials = nullptr;
  unsigned int* counters = nullptr;
  const bool reduce = metrics != nullptr && depth_true != nullptr;
  if (reduce) {
    int dev = 0;
    B200MVS_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
      set_error("b200mvs_depth_metrics: device index out of range");
      return -1;
    }
    EvalScratch& sc = g_scratch[dev];
    // a fixed 64 KB of counters first (always zero between calls), the partial sums after them
    constexpr size_t poff = 65536;
    if ((size_t)batch * sizeof(unsigned int) > poff) {
      set_error("b200mvs_depth_metrics: at most 16384 images per call");
      return -1;
    }
    const size_t need = poff + (size_t)batch * ctas * kSums * sizeof(double);
    if (need > sc.capacity) {
      if (sc.base != nullptr) B200MVS_CUDA_OK(cudaFree(sc.base));   // waits for kernels still using it
      sc.base = nullptr;
      sc.capacity = 0;
      const size_t cap = need < (1u << 20) ? (1u << 20) : need;
      B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&sc.base), cap));
      B200MVS_CUDA_OK(cudaMemset(sc.base, 0, cap));
      sc.capacity = cap;
    }
    counters = reinterpret_cast<unsigned int*>(sc.base);
    partials = reinterpret_cast<double*>(sc.base + poff);
  }
  launch_pdl(depth_metrics_kernel, dim3(ctas, batch), dim3(kThreads), (size_t)0, stream, est, baseline, depth_true,
             est_is_depth ? 1 : 0, min_depth, max_depth, pixels, idepth_out, depth_out, reduce ? metrics : (double*)nullptr,
             partials, counters);
  B200MVS_LAUNCH_OK("depth_metrics_kernel");
  return 0;
}

}  // namespace b200mvs
