// The reprojection layers of the reference's stereo/image_predictor.py that sit next to the hot path:
//   DisparityToIDepth (:120-209), IDepthToDisparity (:211-273), IDepthmapProjector (:525-576),
//   IDepthImagePredictor (:347-398), ImagePredictor (:578-601), RectifiedImagePredictor (:275-345)
// as ONE per-pixel kernel: disparity -> idepth (least squares along the epipolar line) -> back-project -> transform ->
// project -> normalised grid coordinate + out-of-image mask -> bilinear border-clamped sample.  The reference runs
// ~40 torch ops over (B, 3, rows*cols) temporaries for the same thing; here each pixel is read once and every
// requested output is written once.  The 4x4 inverses and products are done in float64 by one thread per image.
#include "kernels.cuh"

namespace b200mvs {
namespace {

__device__ void inv4d(const double* a, double* out) {
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

struct ReprojParams {
  const float* K;        // (n, 4, 4)
  const float* T;        // (n, 4, 4) T_right_in_left
  const float* map;      // (n, 1, rows, cols): idepth (kind 0), disparity (1), rectified disparity (2)
  int kind;
  const float* right;    // (n, C, rows, cols) or null
  int channels, rows, cols;
  float* pred;           // (n, C, rows, cols) or null
  uint8_t* mask;         // (n, 1, rows, cols) or null
  float* right_pixels;   // (n, rows, cols, 2) normalised grid coordinates, or null
  float* right_idepths;  // (n, 1, rows, cols) or null
  float* idepth_out;     // (n, 1, rows, cols) or null (DisparityToIDepth)
  float* disparity_out;  // (n, 1, rows, cols) or null (IDepthToDisparity)
};

struct Mats {
  float Kinv3[9];   // inverse(K)[:3, :3]
  float Tl[12];     // inverse(T_right_in_left)[:3, :]
  float P[12];      // (K @ inverse(T))[:3, :]
  float K3[9];      // K[:3, :3]
  float KRK[9];     // K3 R Kinv3
  float Kt[3];      // (K @ inverse(T))[:3, 3]
  float sign;       // sign(T_right_in_left[0, 3])
};

__global__ void __launch_bounds__(256) reproject_kernel(const ReprojParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ Mats s;
  const int n = blockIdx.y;
  if (threadIdx.x == 0) {
    double K[16], T[16], Kinv[16], Tinv[16], KT[16];
    for (int i = 0; i < 16; ++i) {
      K[i] = (double)p.K[(size_t)n * 16 + i];
      T[i] = (double)p.T[(size_t)n * 16 + i];
    }
    inv4d(K, Kinv);
    inv4d(T, Tinv);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double acc = 0.0;
        for (int k = 0; k < 4; ++k) acc += K[i * 4 + k] * Tinv[k * 4 + j];
        KT[i * 4 + j] = acc;
      }
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        s.Kinv3[i * 3 + j] = (float)Kinv[i * 4 + j];
        s.K3[i * 3 + j] = (float)K[i * 4 + j];
      }
      for (int j = 0; j < 4; ++j) {
        s.Tl[i * 4 + j] = (float)Tinv[i * 4 + j];
        s.P[i * 4 + j] = (float)KT[i * 4 + j];
      }
      s.Kt[i] = (float)KT[i * 4 + 3];
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double acc = 0.0;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) acc += K[i * 4 + a] * Tinv[a * 4 + b] * Kinv[b * 4 + j];
        s.KRK[i * 3 + j] = (float)acc;
      }
    const float tx = p.T[(size_t)n * 16 + 3];
    s.sign = tx > 0.f ? 1.f : (tx < 0.f ? -1.f : 0.f);
  }
  __syncthreads();
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = p.rows * p.cols;
  if (pix >= plane) return;
  const float x = (float)(pix % p.cols), y = (float)(pix / p.cols);
  const float m = __ldg(p.map + (size_t)n * plane + pix);
  float u, v;   // normalised grid coordinates in the right image
  if (p.kind == 2) {
    // RectifiedImagePredictor: shift along x by the signed disparity
    u = ((x - s.sign * m) + 0.5f) * 2.0f / (float)p.cols - 1.0f;
    v = (y + 0.5f) * 2.0f / (float)p.rows - 1.0f;
  } else {
    float idepth = m;
    // pixel at infinite depth (shared by disparity -> idepth and idepth -> disparity)
    const float pz = s.KRK[6] * x + s.KRK[7] * y + s.KRK[8];
    const float infx = (s.KRK[0] * x + s.KRK[1] * y + s.KRK[2]) / pz;
    const float infy = (s.KRK[3] * x + s.KRK[4] * y + s.KRK[5]) / pz;
    if (p.kind == 1) {
      const float fz = 100.0f * pz + s.Kt[2];
      const float farx = (100.0f * (s.KRK[0] * x + s.KRK[1] * y + s.KRK[2]) + s.Kt[0]) / fz;
      const float fary = (100.0f * (s.KRK[3] * x + s.KRK[4] * y + s.KRK[5]) + s.Kt[1]) / fz;
      float ex = farx - infx, ey = fary - infy;
      const float nrm = sqrtf(ex * ex + ey * ey);
      const bool bad = nrm < 1e-6f;
      ex /= (nrm + 1e-6f);
      ey /= (nrm + 1e-6f);
      const float A0 = s.Kt[0] - s.Kt[2] * (infx + m * ex);
      const float A1 = s.Kt[1] - s.Kt[2] * (infy + m * ey);
      const float b0 = pz * m * ex, b1 = pz * m * ey;
      idepth = (A0 * b0 + A1 * b1) / (A0 * A0 + A1 * A1);
      if (bad) idepth = 0.0f * idepth;   // (~mask).float() * idepth: NaN stays NaN as in the reference
    }
    if (p.idepth_out != nullptr) p.idepth_out[(size_t)n * plane + pix] = idepth;
    const float depth = 1.0f / (idepth + 1e-6f);
    const float cx = depth * (s.Kinv3[0] * x + s.Kinv3[1] * y + s.Kinv3[2]);
    const float cy = depth * (s.Kinv3[3] * x + s.Kinv3[4] * y + s.Kinv3[5]);
    const float cz = depth * (s.Kinv3[6] * x + s.Kinv3[7] * y + s.Kinv3[8]);
    const float rx = s.Tl[0] * cx + s.Tl[1] * cy + s.Tl[2] * cz + s.Tl[3];
    const float ry = s.Tl[4] * cx + s.Tl[5] * cy + s.Tl[6] * cz + s.Tl[7];
    const float rz = s.Tl[8] * cx + s.Tl[9] * cy + s.Tl[10] * cz + s.Tl[11];
    if (p.right_idepths != nullptr) p.right_idepths[(size_t)n * plane + pix] = 1.0f / (rz + 1e-6f);
    if (p.disparity_out != nullptr) {
      const float qz = s.K3[6] * rx + s.K3[7] * ry + s.K3[8] * rz;
      const float qx = (s.K3[0] * rx + s.K3[1] * ry + s.K3[2] * rz) / qz;
      const float qy = (s.K3[3] * rx + s.K3[4] * ry + s.K3[5] * rz) / qz;
      const float dx = qx - infx, dy = qy - infy;
      p.disparity_out[(size_t)n * plane + pix] = sqrtf(dx * dx + dy * dy);
    }
    const float wx = s.P[0] * cx + s.P[1] * cy + s.P[2] * cz + s.P[3];
    const float wy = s.P[4] * cx + s.P[5] * cy + s.P[6] * cz + s.P[7];
    const float wz = s.P[8] * cx + s.P[9] * cy + s.P[10] * cz + s.P[11];
    u = (wx / (wz + 1e-7f) + 0.5f) * 2.0f / (float)p.cols - 1.0f;
    v = (wy / (wz + 1e-7f) + 0.5f) * 2.0f / (float)p.rows - 1.0f;
  }
  if (p.right_pixels != nullptr) {
    p.right_pixels[((size_t)n * plane + pix) * 2 + 0] = u;
    p.right_pixels[((size_t)n * plane + pix) * 2 + 1] = v;
  }
  if (p.mask != nullptr) p.mask[(size_t)n * plane + pix] = (fabsf(u) > 1.0f || fabsf(v) > 1.0f) ? 1 : 0;
  if (p.pred != nullptr && p.right != nullptr) {
    // grid_sample(bilinear, padding_mode="border", align_corners=False)
    float ix = ((u + 1.0f) * (float)p.cols - 1.0f) * 0.5f;
    float iy = ((v + 1.0f) * (float)p.rows - 1.0f) * 0.5f;
    ix = fminf(fmaxf(ix, 0.0f), (float)(p.cols - 1));
    iy = fminf(fmaxf(iy, 0.0f), (float)(p.rows - 1));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float we = ix - fx0, ws = iy - fy0;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const int x1 = min(x0 + 1, p.cols - 1), y1 = min(y0 + 1, p.rows - 1);
    const float w00 = (1.0f - ws) * (1.0f - we), w01 = (1.0f - ws) * we, w10 = ws * (1.0f - we), w11 = ws * we;
    for (int c = 0; c < p.channels; ++c) {
      const float* pl = p.right + ((size_t)n * p.channels + c) * plane;
      p.pred[((size_t)n * p.channels + c) * plane + pix] =
          __ldg(pl + y0 * p.cols + x0) * w00 + __ldg(pl + y0 * p.cols + x1) * w01 + __ldg(pl + y1 * p.cols + x0) * w10 +
          __ldg(pl + y1 * p.cols + x1) * w11;
    }
  }
}

}  // namespace

int launch_reproject(const float* K, const float* T, const float* map, int kind, const float* right, int n,
                     int channels, int rows, int cols, float* pred, uint8_t* mask, float* right_pixels,
                     float* right_idepths, float* idepth_out, float* disparity_out, cudaStream_t stream) {
  if (n <= 0) return 0;
  ReprojParams p;
  p.K = K;
  p.T = T;
  p.map = map;
  p.kind = kind;
  p.right = right;
  p.channels = channels;
  p.rows = rows;
  p.cols = cols;
  p.pred = pred;
  p.mask = mask;
  p.right_pixels = right_pixels;
  p.right_idepths = right_idepths;
  p.idepth_out = idepth_out;
  p.disparity_out = disparity_out;
  dim3 grid(cdiv(rows * cols, 256), n);
  launch_pdl(reproject_kernel, grid, dim3(256), (size_t)0, stream, p);
  B200MVS_LAUNCH_OK("reproject_kernel");
  return 0;
}

}  // namespace b200mvs
