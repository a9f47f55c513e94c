// Shared definitions for the b200mvs kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>

namespace b200mvs {

constexpr int kC = 32;            // feature channels everywhere (multi_view_stereonet.py:87)
constexpr int kGroups = 4;        // GroupNorm(32 // 8, 32)  (multi_view_stereonet.py:25-31)
constexpr float kGnEps = 1e-5f;
constexpr float kLreluSlope = 0.2f;  // multi_view_stereonet.py:64

void set_error(const std::string& msg);
// Counts kernel launches of the current forward (api.cu).
void note_launch();

#define B200MVS_CUDA_OK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::b200mvs::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));            \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

#define B200MVS_LAUNCH_OK(what)                                                            \
  do {                                                                                     \
    ::b200mvs::note_launch();                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess) {                                                               \
      ::b200mvs::set_error(std::string(what) + ": " + cudaGetErrorString(_e));             \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  The forward is a chain of ~70 short dependent kernels; each one is
// launched with the programmatic-stream-serialisation attribute so that its CTAs become resident and
// run their prologue (TMEM allocation, barrier init, weight staging) while the previous kernel
// drains.  Contract for every kernel of this library:
//   * acquire on-chip resources (TMEM) first, then pdl_launch_dependents();
//   * do not read anything a previous kernel wrote before pdl_wait();
//   * every thread executes pdl_wait() before the kernel exits, so "kernel N complete" implies
//     "kernels < N complete" (a kernel may read buffers written several launches earlier).
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();
void set_pdl_enabled(bool on);

// cudaFuncSetAttribute applies to the device that is current when it is called, and one process may hold handles
// on several GPUs (b200mvs_create takes a device index): the opt-ins for > 48 KB of dynamic shared memory and for
// non-portable cluster sizes are therefore cached per (kernel, device), under a mutex (api.cu).
int ensure_func_smem(const void* func, size_t bytes);      // raises the limit if it is below `bytes`
int ensure_func_nonportable_cluster(const void* func);
int current_device_sm_count(int* sms);
// Remembers the cluster size that could be scheduled for (device, tiles); 0 = unknown.
int cached_cluster_size(int tiles);
void remember_cluster_size(int tiles, int cluster);

#ifdef __CUDACC__
// L2 residency hints.  The refiner activations ping-pong between four buffers that together fit in the 126 MB L2:
// stores of tensors the next layer reads are marked evict_last, loads of tensors that are dead after this layer
// evict_first, so the live set stays on chip instead of round-tripping through HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ldg_hint(const void* ptr, uint64_t policy) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void stg_hint(void* ptr, const uint4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w), "l"(policy)
               : "memory");
}

// 256-bit global accesses (sm_100: LDG / STG .256): a thread's 8 channels of one position in ONE instruction -- the
// L1 processes a warp instruction once per 128-byte line it touches, so two 16-byte halves cost twice the wavefronts.
// ld8: weak load that may be served by L1 (never the non-coherent path); _cg: L2 only; _nc: read-only data.
#define B200MVS_LD8(name, op)                                                                                        \
  __device__ __forceinline__ void name(const float* p, float* v) {                                                  \
    asm volatile(op " {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                                       \
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])    \
                 : "l"(p)                                                                                            \
                 : "memory");                                                                                        \
  }
B200MVS_LD8(ld8, "ld.global.v8.f32")
B200MVS_LD8(ld8_cg, "ld.global.cg.v8.f32")
B200MVS_LD8(ld8_nc, "ld.global.nc.v8.f32")
#undef B200MVS_LD8
#define B200MVS_ST8(name, op)                                                                                        \
  __device__ __forceinline__ void name(float* p, const float* v) {                                                  \
    asm volatile(op " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), \
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])                                                          \
                 : "memory");                                                                                        \
  }
B200MVS_ST8(st8, "st.global.v8.f32")
B200MVS_ST8(st8_cg, "st.global.cg.v8.f32")
#undef B200MVS_ST8

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// Homography pixel transfer with the reference's float32 operation order.
//
// stereo/image_predictor.py:493-516 computes  p = H @ (x, y, 1)  with a float32 GEMM (each output
// is the FMA chain  fma(h2, 1, fma(h1, y, h0 * x)) -- verified bit-exact against torch.matmul on
// the CPU), divides by p.z with no epsilon, maps to grid_sample's normalised coordinate
// u = ((px + 0.5) * 2) / size - 1  and masks where |u| > 1.  The mask is a hard threshold, so the
// same roundings are reproduced here with explicit non-contracted intrinsics; out-of-image
// decisions then agree with the reference wherever they are not decided by the last bit of H.
// ---------------------------------------------------------------------------------------------
struct WarpCoord {
  float ix, iy;   // source coordinates in pixels, border-clamped to [0, size-1]
  bool invalid;   // True = outside the image (reference mask convention)
};

__device__ __forceinline__ WarpCoord homography_coord(const float* __restrict__ H, float x, float y,
                                                      int rows, int cols) {
  const float X = __fadd_rn(__fmaf_rn(H[1], y, __fmul_rn(H[0], x)), H[2]);
  const float Y = __fadd_rn(__fmaf_rn(H[4], y, __fmul_rn(H[3], x)), H[5]);
  const float Z = __fadd_rn(__fmaf_rn(H[7], y, __fmul_rn(H[6], x)), H[8]);
  const float px = __fdiv_rn(X, Z);
  const float py = __fdiv_rn(Y, Z);
  const float u = __fsub_rn(__fdiv_rn(__fmul_rn(__fadd_rn(px, 0.5f), 2.0f), (float)cols), 1.0f);
  const float v = __fsub_rn(__fdiv_rn(__fmul_rn(__fadd_rn(py, 0.5f), 2.0f), (float)rows), 1.0f);
  WarpCoord r;
  r.invalid = (fabsf(u) > 1.0f) || (fabsf(v) > 1.0f);
  // grid_sample(align_corners=False) un-normalisation as ATen's CPU kernel does it:
  // (u + 1) * (size / 2) - 0.5, then border clipping to [0, size - 1].
  float ix = __fsub_rn(__fmul_rn(__fadd_rn(u, 1.0f), 0.5f * (float)cols), 0.5f);
  float iy = __fsub_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f * (float)rows), 0.5f);
  r.ix = fminf(fmaxf(ix, 0.0f), (float)(cols - 1));
  r.iy = fminf(fmaxf(iy, 0.0f), (float)(rows - 1));
  return r;
}

struct Bilinear {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

__device__ __forceinline__ Bilinear bilinear_setup(const WarpCoord& c, int rows, int cols) {
  Bilinear b;
  const float fx0 = floorf(c.ix), fy0 = floorf(c.iy);
  const float we = c.ix - fx0, ws = c.iy - fy0;   // weight of the east / south neighbour
  const float ww = 1.0f - we, wn = 1.0f - ws;
  b.x0 = (int)fx0;
  b.y0 = (int)fy0;
  b.x1 = min(b.x0 + 1, cols - 1);   // the out-of-range neighbour always carries weight 0
  b.y1 = min(b.y0 + 1, rows - 1);
  b.w00 = wn * ww;
  b.w01 = wn * we;
  b.w10 = ws * ww;
  b.w11 = ws * we;
  return b;
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize, align_corners=False (ATen upsample_bilinear2d): src = scale*(dst+0.5)-0.5
// clamped at 0, i1 = i0 + (i0 < in-1), l1 = src - i0.
// ---------------------------------------------------------------------------------------------
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_setup(int dst, float scale, int in_size) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = src < 0.f ? 0.f : src;
  Lerp r;
  r.i0 = (int)src;
  if (r.i0 > in_size - 1) r.i0 = in_size - 1;
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = __fsub_rn(src, (float)r.i0);
  r.l0 = __fsub_rn(1.0f, r.l1);
  return r;
}
// One output pixel of F.interpolate(size=(H, W), mode="bilinear", align_corners=False) of a (h, w) map
// (Upsampler.forward, multi_view_stereonet.py:372-380); shared by the stand-alone upsampling kernel and the
// refiner head that upsamples its prior on the fly, so both produce the same bits.
__device__ __forceinline__ float upsample_bilinear_at(const float* __restrict__ src, int h, int w, int H, int W, int y,
                                                      int x) {
  const Lerp ly = lerp_setup(y, (float)h / (float)H, h);
  const Lerp lx = lerp_setup(x, (float)w / (float)W, w);
  const float v00 = __ldg(src + ly.i0 * w + lx.i0), v01 = __ldg(src + ly.i0 * w + lx.i1);
  const float v10 = __ldg(src + ly.i1 * w + lx.i0), v11 = __ldg(src + ly.i1 * w + lx.i1);
  return ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11);
}

__device__ __forceinline__ float lrelu(float v) { return v > 0.0f ? v : kLreluSlope * v; }

// 1 / sqrt(var + eps) of a GroupNorm from float64 statistics.  Every conv kernel computes this on its critical path
// (32 threads, then a block barrier): the float32 MUFU estimate refined by one Newton step in float64 (relative
// error ~1e-13, i.e. the same float32 coefficients after rounding) replaces the float64 library rsqrt, whose
// software iteration cost ~1 us per layer (measured with B200MVS_TC_PROFILE, "coeffs").
__device__ __forceinline__ double gn_rstd(double var) {
  const double x = (var > 0.0 ? var : 0.0) + (double)kGnEps;
  const double r = (double)rsqrtf((float)x);
  return r * (1.5 - 0.5 * x * r * r);
}

}  // namespace b200mvs
