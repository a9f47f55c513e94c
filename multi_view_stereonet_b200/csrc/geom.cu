// Per-(image group, view) geometry: baseline normalisation, idepth sample range, plane-sweep
// homographies and their increments.  One CTA per (b, v).
//
// The reference does this with ~60 tiny torch ops per view (multi_view_stereonet.py:566-576,
// 131-194, 280-282; stereo/image_predictor.py:120-209, 446-459).  The small matrix algebra is done
// here in float64 by one thread (it is a few hundred flops) and rounded to float32 once; the
// per-pixel least-squares idepth over the 1/16-scale grid is spread over the CTA.
#include "kernels.cuh"

namespace b200mvs {
namespace {

__device__ void inv4(const double* a, double* out) {
  // Gauss-Jordan with partial pivoting on [a | I].
  double m[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      m[i][j] = a[i * 4 + j];
      m[i][4 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    double best = fabs(m[c][c]);
    for (int r = c + 1; r < 4; ++r)
      if (fabs(m[r][c]) > best) {
        best = fabs(m[r][c]);
        piv = r;
      }
    if (piv != c)
      for (int j = 0; j < 8; ++j) {
        const double tmp = m[c][j];
        m[c][j] = m[piv][j];
        m[piv][j] = tmp;
      }
    const double d = 1.0 / m[c][c];
    for (int j = 0; j < 8; ++j) m[c][j] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        const double f = m[r][c];
        for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][4 + j];
}

__device__ void inv3(const double* a, double* o) {
  const double c00 = a[4] * a[8] - a[5] * a[7];
  const double c01 = a[5] * a[6] - a[3] * a[8];
  const double c02 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c01 * id;
  o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c02 * id;
  o[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

__device__ void mul3(const double* a, const double* b, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

struct Shared {
  double R[9], t[3];       // inverse(T_right_in_left normalised)
  double K4[9], K4inv[9];  // level-4 intrinsics 3x3 and inverse
  double KRK[9], Kt[3];    // for disparity_to_idepth (uses the 4x4 K inverse / product)
  float tz;
  double red_sum[32];
  int red_cnt[32];
  float delta;
};

__global__ void __launch_bounds__(640) geometry_kernel(ViewPtrs T, const float* __restrict__ K0p,
                                                       const float* __restrict__ K4p, int D, int rows4, int cols4,
                                                       GeomOut out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ Shared s;
  const int n = blockIdx.x;
  const int b = n / T.views, v = n % T.views;
  const int tid = threadIdx.x;
  const float* Tin = T.p[v] + (size_t)b * 16;
  const float* K0 = K0p + (size_t)b * 16;
  const float* K4 = K4p + (size_t)b * 16;

  if (tid == 0) {
    // Baseline normalisation in float32 as the reference does it (multi_view_stereonet.py:568-571).
    const float t0 = Tin[3], t1 = Tin[7], t2 = Tin[11];
    const float baseline = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(t0, t0), __fmul_rn(t1, t1)), __fmul_rn(t2, t2)));
    out.baseline[n] = baseline;
    double Tn[16], Tinv[16];
    for (int i = 0; i < 16; ++i) Tn[i] = (double)Tin[i];
    Tn[3] = (double)__fdiv_rn(t0, baseline);
    Tn[7] = (double)__fdiv_rn(t1, baseline);
    Tn[11] = (double)__fdiv_rn(t2, baseline);
    s.tz = __fdiv_rn(t2, baseline);
    inv4(Tn, Tinv);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) s.R[i * 3 + j] = Tinv[i * 4 + j];
      s.t[i] = Tinv[i * 4 + 3];
    }
    // Level-4 intrinsics.
    double K[16], Kinv[16], K3[9], K3inv[9], tmp[9];
    for (int i = 0; i < 16; ++i) K[i] = (double)K4[i];
    inv4(K, Kinv);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        K3[i * 3 + j] = K[i * 4 + j];
        tmp[i * 3 + j] = Kinv[i * 4 + j];
      }
    // KRKinv = K[:3,:3] R Kinv[:3,:3]   (image_predictor.py:153)
    double RK[9];
    mul3(s.R, tmp, RK);
    mul3(K3, RK, s.KRK);
    // Kt = (K @ T_left_in_right)[:3, 3]   (image_predictor.py:158-159)
    for (int i = 0; i < 3; ++i) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc += K[i * 4 + k] * Tinv[k * 4 + 3];
      s.Kt[i] = acc;
    }
    inv3(K3, K3inv);
    for (int i = 0; i < 9; ++i) {
      s.K4[i] = K3[i];
      s.K4inv[i] = K3inv[i];
    }
    // Level-0 homography of hypothesis 0 (idepth 0): K0 R K0^-1  (multi_view_stereonet.py:254-255).
    double K03[9], K03inv[9], H0[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) K03[i * 3 + j] = (double)K0[i * 4 + j];
    inv3(K03, K03inv);
    mul3(s.R, K03inv, tmp);
    mul3(K03, tmp, H0);
    for (int i = 0; i < 9; ++i) out.H0[(size_t)n * 9 + i] = (float)H0[i];
  }
  __syncthreads();

  // disparity_to_idepth for a constant disparity of D-1 px at level 4 (multi_view_stereonet.py:139-141).
  const double disp = (double)(D - 1);
  double sum = 0.0;
  int cnt = 0;
  for (int p = tid; p < rows4 * cols4; p += blockDim.x) {
    const double x = (double)(p % cols4), y = (double)(p / cols4);
    const double pz = s.KRK[6] * x + s.KRK[7] * y + s.KRK[8];
    const double ix = (s.KRK[0] * x + s.KRK[1] * y + s.KRK[2]) / pz;
    const double iy = (s.KRK[3] * x + s.KRK[4] * y + s.KRK[5]) / pz;
    const double fzz = 100.0 * pz + s.Kt[2];
    const double fx = (100.0 * (s.KRK[0] * x + s.KRK[1] * y + s.KRK[2]) + s.Kt[0]) / fzz;
    const double fy = (100.0 * (s.KRK[3] * x + s.KRK[4] * y + s.KRK[5]) + s.Kt[1]) / fzz;
    double ex = fx - ix, ey = fy - iy;
    const double nrm = sqrt(ex * ex + ey * ey);
    const bool invalid = nrm < 1e-6;
    ex /= (nrm + 1e-6);
    ey /= (nrm + 1e-6);
    const double A0 = s.Kt[0] - s.Kt[2] * (ix + disp * ex);
    const double A1 = s.Kt[1] - s.Kt[2] * (iy + disp * ey);
    const double b0 = pz * disp * ex, b1 = pz * disp * ey;
    double idepth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
    if (invalid) idepth = 0.0;
    if (idepth > 0.0) {
      sum += idepth;
      ++cnt;
    } else if (idepth != idepth) {
      sum += idepth;  // NaN propagates as in the reference
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((tid & 31) == 0) {
    s.red_sum[tid >> 5] = sum;
    s.red_cnt[tid >> 5] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    int c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      tot += s.red_sum[w];
      c += s.red_cnt[w];
    }
    float mx = (float)(tot / (double)c);                       // mean over positive idepths (:143-148)
    if (mx > 2.0f) mx = 2.0f;                                   // :149
    if (__fdiv_rn(1.0f, mx) < s.tz) mx = __fdiv_rn(1.0f, s.tz);  // :152-154
    s.delta = __fdiv_rn(mx, (float)(D - 1));                    // :158
  }
  __syncthreads();

  // Samples and homographies, one hypothesis per thread.
  for (int d = tid; d < D; d += blockDim.x) {
    out.samples[(size_t)n * D + d] = __fmul_rn((float)d, s.delta);
  }
  for (int d = tid; d < D; d += blockDim.x) {
    double Hc[9], Hp[9], M[9], tmp[9];
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: H_{d-1} (only needed for the increment), pass 1: H_d
      const int dd = d - 1 + pass;
      if (dd < 0) continue;
      const double sd = (double)__fmul_rn((float)dd, s.delta);
      for (int i = 0; i < 9; ++i) M[i] = s.R[i];
      M[2] += s.t[0] * sd;
      M[5] += s.t[1] * sd;
      M[8] += s.t[2] * sd;
      mul3(M, s.K4inv, tmp);
      mul3(s.K4, tmp, pass == 0 ? Hp : Hc);
    }
    float* Ho = out.H + ((size_t)n * D + d) * 9;
    float* Io = out.Hinc + ((size_t)n * D + d) * 9;
    for (int i = 0; i < 9; ++i) Ho[i] = (float)Hc[i];
    if (d == 0) {
      for (int i = 0; i < 9; ++i) Io[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    } else {
      // H_inc = inverse(H_{d-1}) @ H_d  (multi_view_stereonet.py:280-282).  The reference inverts
      // the float32 H; the float32-rounded matrices are used here too so that the increment is the
      // one between the homographies actually applied.
      double Hpf[9], Hcf[9], Hpi[9];
      for (int i = 0; i < 9; ++i) {
        Hpf[i] = (double)(float)Hp[i];
        Hcf[i] = (double)(float)Hc[i];
      }
      inv3(Hpf, Hpi);
      mul3(Hpi, Hcf, tmp);
      for (int i = 0; i < 9; ++i) Io[i] = (float)tmp[i];
    }
  }
}

}  // namespace

int launch_geometry(const ViewPtrs& T, const float* K0, const float* K4, int batch, int D, int rows4, int cols4,
                    const GeomOut& out, cudaStream_t stream) {
  launch_pdl(geometry_kernel, dim3(batch * T.views), dim3(640), (size_t)0, stream, T, K0, K4, D, rows4, cols4, out);
  B200MVS_LAUNCH_OK("geometry_kernel");
  return 0;
}

}  // namespace b200mvs
