// geometry.cu — library globals, batched camera-model kernels (K7) and the fused two-view
// triangulation kernel (K6).
//
// Replaces (mavmap/mavmap):
//   camera_model_world2image / image2world      src/base3d/camera_models.h:375-423
//   camera_model_image2world (vector overload)  src/base3d/camera_models.cc:24-44
//   camera_model_name_to_code                   src/base3d/camera_models.cc:12-21
//   camera_model_image2world_threshold          src/base3d/camera_models.cc:47-52
//   triangulate_point(s)                        src/base3d/triangulation.cc:12-74
//   calc_tri_angles                             src/base3d/triangulation.cc:101-147
//   calc_reproj_errors / calc_depth             src/base3d/projection.cc:107-149
// These are latency-bound (<= a few thousand items per call): one thread per item, grid sized
// to the SM count, P matrices and intrinsics passed by value in the launch parameters.
#include <stdarg.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "camera.cuh"

namespace mm {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}

int ensure_device() {
  static std::once_flag once; static int status = MM_ERR_NO_DEVICE;
  std::call_once(once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return; }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (prop.major != 10) { set_error("device %s is sm_%d%d; this library carries sm_100a code only", prop.name, prop.major, prop.minor); return; }
    status = MM_OK;
  });
  if (status != MM_OK && !g_err[0]) set_error("no usable CUDA device (sm_100a required; there is no CPU fallback)");
  return status;
}

struct Params9 { double p[9]; };
struct Proj { double P[12]; };

__global__ void k_world2image(int model, Params9 prm, int64_t n, const double* __restrict__ xyz, double* __restrict__ uv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double u, v;
    world2image<false>(model, prm.p, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], u, v, nullptr, nullptr);
    uv[2 * i] = u; uv[2 * i + 1] = v;
  }
}

__global__ void k_image2world(int model, Params9 prm, int64_t n, const double* __restrict__ uv, double* __restrict__ out, int normalized) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double x, y, z;
    image2world(model, prm.p, uv[2 * i], uv[2 * i + 1], x, y, z);
    if (normalized) { out[2 * i] = x / z; out[2 * i + 1] = y / z; }       // camera_models.cc:41-42
    else { out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z; }
  }
}

// One-sided Jacobi SVD of the 6x4 DLT matrix, all in registers; the null vector is the column
// of V that belongs to the smallest singular value (triangulation.cc:40-41).
__device__ __forceinline__ void null_vector_6x4(double (&A)[6][4], double* h) {
  double V[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) V[r][c] = r == c ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double app = 0, aqq = 0, apq = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) { app += A[r][p] * A[r][p]; aqq += A[r][q] * A[r][q]; apq += A[r][p] * A[r][q]; }
        const double denom = sqrt(app * aqq);
        if (apq != 0.0 && denom != 0.0 && fabs(apq) > 1e-300) {
          off = fmax(off, fabs(apq) / denom);
          const double zeta = (aqq - app) / (2.0 * apq);
          const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
#pragma unroll
          for (int r = 0; r < 6; ++r) { const double x = A[r][p], y = A[r][q]; A[r][p] = c * x - s * y; A[r][q] = s * x + c * y; }
#pragma unroll
          for (int r = 0; r < 4; ++r) { const double x = V[r][p], y = V[r][q]; V[r][p] = c * x - s * y; V[r][q] = s * x + c * y; }
        }
      }
    if (off < 1e-17) break;
  }
  double bn = INFINITY; int best = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double nn = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) nn += A[r][c] * A[r][c];
    if (nn < bn) { bn = nn; best = c; }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) h[r] = best == 0 ? V[r][0] : (best == 1 ? V[r][1] : (best == 2 ? V[r][2] : V[r][3]));
}

__device__ __forceinline__ double reproj_err(const double* P, const double* X, double x, double y) {
  const double q0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
  const double q1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
  const double q2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
  const double dx = q0 / q2 - x, dy = q1 / q2 - y;
  return sqrt(dx * dx + dy * dy);
}

__global__ void k_triangulate(Proj p1, Proj p2, double c1x, double c1y, double c1z, double c2x, double c2y, double c2z,
                              double baseline2, int64_t n, const double* __restrict__ x1, const double* __restrict__ x2,
                              double* __restrict__ X, double* __restrict__ rp1, double* __restrict__ rp2,
                              double* __restrict__ dp1, double* __restrict__ dp2, double* __restrict__ ang) {
  const double* P1 = p1.P; const double* P2 = p2.P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xa = x1[2 * i], ya = x1[2 * i + 1], xb = x2[2 * i], yb = x2[2 * i + 1];
    double A[6][4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                       // triangulation.cc:26-35
      A[0][k] = xa * P1[8 + k] - P1[k];
      A[1][k] = ya * P1[8 + k] - P1[4 + k];
      A[2][k] = xa * P1[4 + k] - ya * P1[k];
      A[3][k] = xb * P2[8 + k] - P2[k];
      A[4][k] = yb * P2[8 + k] - P2[4 + k];
      A[5][k] = xb * P2[4 + k] - yb * P2[k];
    }
    null_vector_6x4(A, h);
    double Xi[3] = { h[0] / h[3], h[1] / h[3], h[2] / h[3] };
    X[3 * i] = Xi[0]; X[3 * i + 1] = Xi[1]; X[3 * i + 2] = Xi[2];
    if (rp1) rp1[i] = reproj_err(P1, Xi, xa, ya);
    if (rp2) rp2[i] = reproj_err(P2, Xi, xb, yb);
    if (dp1) dp1[i] = (P1[8] * Xi[0] + P1[9] * Xi[1] + P1[10] * Xi[2] + P1[11]) * sqrt(P1[2] * P1[2] + P1[6] * P1[6] + P1[10] * P1[10]);
    if (dp2) dp2[i] = (P2[8] * Xi[0] + P2[9] * Xi[1] + P2[10] * Xi[2] + P2[11]) * sqrt(P2[2] * P2[2] + P2[6] * P2[6] + P2[10] * P2[10]);
    if (ang) {
      const double r1 = sqrt((Xi[0] - c1x) * (Xi[0] - c1x) + (Xi[1] - c1y) * (Xi[1] - c1y) + (Xi[2] - c1z) * (Xi[2] - c1z));
      const double r2 = sqrt((Xi[0] - c2x) * (Xi[0] - c2x) + (Xi[1] - c2y) * (Xi[1] - c2y) + (Xi[2] - c2z) * (Xi[2] - c2z));
      const double a = acos((r1 * r1 + r2 * r2 - baseline2) / (2.0 * r1 * r2));
      ang[i] = isnan(a) ? 0.0 : a;                       // triangulation.cc:134-140
    }
  }
}

__global__ void k_reproj(Proj p, int64_t n, const double* __restrict__ x, const double* __restrict__ X, double* __restrict__ err, double* __restrict__ depth) {
  const double* P = p.P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double Xi[3] = { X[3 * i], X[3 * i + 1], X[3 * i + 2] };
    if (err) err[i] = reproj_err(P, Xi, x[2 * i], x[2 * i + 1]);
    if (depth) depth[i] = (P[8] * Xi[0] + P[9] * Xi[1] + P[10] * Xi[2] + P[11]) * sqrt(P[2] * P[2] + P[6] * P[6] + P[10] * P[10]);
  }
}

// calc_tri_angles (triangulation.cc:101-147) for GIVEN 3-D points: law of cosines in the triangle (centre 1, centre 2, point)
__global__ void k_tri_angles(double c1x, double c1y, double c1z, double c2x, double c2y, double c2z, double baseline2, int64_t n,
                             const double* __restrict__ X, double* __restrict__ ang) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
    const double r1 = sqrt((x - c1x) * (x - c1x) + (y - c1y) * (y - c1y) + (z - c1z) * (z - c1z));
    const double r2 = sqrt((x - c2x) * (x - c2x) + (y - c2y) * (y - c2y) + (z - c2z) * (z - c2z));
    const double a = acos((r1 * r1 + r2 * r2 - baseline2) / (2.0 * r1 * r2));
    ang[i] = isnan(a) ? 0.0 : a;                         // triangulation.cc:134-140
  }
}

static void camera_center(const double* P, double* C) {       // projection.cc:82-88 translation column
  const double a = P[0], b = P[1], c = P[2], d = P[4], e = P[5], f = P[6], g = P[8], h = P[9], i = P[10];
  const double A = e * i - f * h, B = -(d * i - f * g), Cc = d * h - e * g;
  const double det = a * A + b * B + c * Cc;
  const double inv[9] = { A / det, (c * h - b * i) / det, (b * f - c * e) / det,
                          B / det, (a * i - c * g) / det, (c * d - a * f) / det,
                          Cc / det, (b * g - a * h) / det, (a * e - b * d) / det };
  const double t[3] = { P[3], P[7], P[11] };
  for (int r = 0; r < 3; ++r) C[r] = -(inv[3 * r] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]);
}

static int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int cap = num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int camera_batch(int which, int model, const double* params, int64_t n, const double* in, double* out) {
  if (n < 0 || !params || model_num_params(model) < 0 || (n > 0 && (!in || !out))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc != MM_OK) return rc;
  if (n == 0) return MM_OK;
  const int in_w = which == 0 ? 3 : 2, out_w = which == 0 ? 2 : (which == 1 ? 3 : 2);
  DevBuf<double> din, dout;
  MM_CUDA(din.alloc((size_t)n * in_w)); MM_CUDA(dout.alloc((size_t)n * out_w));
  MM_CUDA(cudaMemcpy(din.p, in, sizeof(double) * (size_t)n * in_w, cudaMemcpyHostToDevice));
  Params9 prm; memset(&prm, 0, sizeof prm);
  memcpy(prm.p, params, sizeof(double) * (size_t)model_num_params(model));
  const int block = 128, grid = grid_for(n, block);
  if (which == 0) k_world2image<<<grid, block>>>(model, prm, n, din.p, dout.p);
  else k_image2world<<<grid, block>>>(model, prm, n, din.p, dout.p, which == 2);
  MM_LAUNCH_CHECK();
  MM_CUDA(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n * out_w, cudaMemcpyDeviceToHost));
  return MM_OK;
}

}  // namespace mm

using namespace mm;


// ---- RANSAC hypothesis scoring ------------------------------------------------------------------
// The reference scores every hypothesis with an OpenMP loop over all correspondences (util/estimation.cc:83-116 calling
// P3PEstimator::residuals p3p.cc:172-199, ProjectiveTransformEstimator::residuals projective_transform.cc:48-74,
// EssentialMatrixEstimator::residuals essential_matrix.cc:131-162).  Here: one CTA per hypothesis, threads stride over the
// correspondences, inlier test |r| <= threshold, block reduction of (inlier count, sum of |r| over inliers).
// Arithmetic is written without FMA contraction and in the reference's operation order, so residuals - and with them the
// inlier masks - are bit-identical to the CPU oracle.
#define MM_RANSAC_P3P 0
#define MM_RANSAC_HOMOGRAPHY 1
#define MM_RANSAC_ESSENTIAL 2
__device__ __forceinline__ double dot3_rn(double a0, double a1, double a2, double b0, double b1, double b2) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}
__device__ __forceinline__ double ransac_residual(int kind, const double* m, const double* x, const double* y, int64_t i) {
  if (kind == MM_RANSAC_P3P) {            // x = points2D [n,2], y = points3D [n,3], m = [R | t] 3x4 row-major
    const double X0 = y[3 * i], X1 = y[3 * i + 1], X2 = y[3 * i + 2];
    double p0 = __dadd_rn(dot3_rn(m[0], m[1], m[2], X0, X1, X2), m[3]);
    double p1 = __dadd_rn(dot3_rn(m[4], m[5], m[6], X0, X1, X2), m[7]);
    const double p2 = __dadd_rn(dot3_rn(m[8], m[9], m[10], X0, X1, X2), m[11]);
    p0 = __ddiv_rn(p0, p2); p1 = __ddiv_rn(p1, p2);
    const double dx = __dadd_rn(p0, -x[2 * i]), dy = __dadd_rn(p1, -x[2 * i + 1]);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  } else if (kind == MM_RANSAC_HOMOGRAPHY) {   // x = src [n,2], y = dst [n,2], m = H 3x3 row-major
    const double s0 = x[2 * i], s1 = x[2 * i + 1];
    const double t0 = dot3_rn(m[0], m[1], m[2], s0, s1, 1.0), t1 = dot3_rn(m[3], m[4], m[5], s0, s1, 1.0), t2 = dot3_rn(m[6], m[7], m[8], s0, s1, 1.0);
    const double dx = __dadd_rn(__ddiv_rn(t0, t2), -y[2 * i]), dy = __dadd_rn(__ddiv_rn(t1, t2), -y[2 * i + 1]);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  } else {                                      // Sampson distance; x = points1, y = points2, m = E 3x3 row-major
    const double a0 = x[2 * i], a1 = x[2 * i + 1], b0 = y[2 * i], b1 = y[2 * i + 1];
    const double e0 = dot3_rn(m[0], m[1], m[2], a0, a1, 1.0), e1 = dot3_rn(m[3], m[4], m[5], a0, a1, 1.0), e2 = dot3_rn(m[6], m[7], m[8], a0, a1, 1.0);
    const double f0 = dot3_rn(m[0], m[3], m[6], b0, b1, 1.0), f1 = dot3_rn(m[1], m[4], m[7], b0, b1, 1.0);
    const double num = dot3_rn(b0, b1, 1.0, e0, e1, e2);
    const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1)), __dmul_rn(f0, f0)), __dmul_rn(f1, f1));
    return __ddiv_rn(num, __dsqrt_rn(den));
  }
}
__global__ void __launch_bounds__(256) k_ransac_score(int kind, int msize, const double* __restrict__ models, int64_t n,
                                                      const double* __restrict__ x, const double* __restrict__ y, double threshold,
                                                      int* __restrict__ num_inliers, double* __restrict__ residual_sum) {
  __shared__ double sm_m[12]; __shared__ double sred[32]; __shared__ int cred[32];
  if (threadIdx.x < msize) sm_m[threadIdx.x] = models[(size_t)blockIdx.x * msize + threadIdx.x];
  __syncthreads();
  int cnt = 0; double sum = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fabs(ransac_residual(kind, sm_m, x, y, i));
    if (a <= threshold) { ++cnt; sum += a; }
  }
  sum = mm::warp_sum(sum);
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sred[w] = sum; cred[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0; int c = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s += sred[k]; c += cred[k]; }
    num_inliers[blockIdx.x] = c; residual_sum[blockIdx.x] = s;
  }
}
__global__ void k_ransac_residuals(int kind, int msize, const double* __restrict__ model, int64_t n, const double* __restrict__ x,
                                   const double* __restrict__ y, double threshold, double* __restrict__ res, unsigned char* __restrict__ mask) {
  __shared__ double sm_m[12];
  if (threadIdx.x < msize) sm_m[threadIdx.x] = model[threadIdx.x];
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = ransac_residual(kind, sm_m, x, y, i);
  if (res) res[i] = r;
  if (mask) mask[i] = fabs(r) <= threshold ? 1 : 0;
}

extern "C" {

int mm_abi_version(void) { return MM_ABI_VERSION; }
const char* mm_last_error(void) { return g_err; }
uint64_t mm_kernel_launch_count(void) { return g_launches.load(); }

int mm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mm_camera_model_name_to_code(const char* name) {
  if (!name) return -1;
  if (!strcmp(name, "PINHOLE")) return MM_MODEL_PINHOLE;
  if (!strcmp(name, "OPENCV")) return MM_MODEL_OPENCV;
  if (!strcmp(name, "CATA")) return MM_MODEL_CATA;
  return -1;
}

int mm_camera_model_num_params(int model_code) { return model_num_params(model_code); }

double mm_camera_image2world_threshold(double threshold, int model_code, const double* params) {
  (void)model_code;
  return threshold / ((params[0] + params[1]) / 2);
}

int mm_camera_world2image(int model, const double* params, int64_t n, const double* xyz, double* uv) {
  return camera_batch(0, model, params, n, xyz, uv);
}
int mm_camera_image2world(int model, const double* params, int64_t n, const double* uv, double* xyz) {
  return camera_batch(1, model, params, n, uv, xyz);
}
int mm_camera_image2world_normalized(int model, const double* params, int64_t n, const double* uv, double* xy) {
  return camera_batch(2, model, params, n, uv, xy);
}

int mm_triangulate_two_view(const double* P1, const double* P2, int64_t n, const double* x1, const double* x2,
                            double* X, double* reproj1, double* reproj2, double* depth1, double* depth2, double* angle) {
  if (n < 0 || !P1 || !P2 || (n > 0 && (!x1 || !x2 || !X))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc != MM_OK) return rc;
  if (n == 0) return MM_OK;
  Proj p1, p2; memcpy(p1.P, P1, sizeof p1.P); memcpy(p2.P, P2, sizeof p2.P);
  double C1[3], C2[3]; camera_center(P1, C1); camera_center(P2, C2);
  const double bl = sqrt((C1[0] - C2[0]) * (C1[0] - C2[0]) + (C1[1] - C2[1]) * (C1[1] - C2[1]) + (C1[2] - C2[2]) * (C1[2] - C2[2]));
  // one staging buffer: [x1 | x2 | X | 5 optional outputs]
  DevBuf<double> buf;
  MM_CUDA(buf.alloc((size_t)n * (2 + 2 + 3 + 5)));
  double* dx1 = buf.p; double* dx2 = dx1 + 2 * n; double* dX = dx2 + 2 * n; double* dout = dX + 3 * n;
  MM_CUDA(cudaMemcpy(dx1, x1, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice));
  MM_CUDA(cudaMemcpy(dx2, x2, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice));
  double* host_out[5] = { reproj1, reproj2, depth1, depth2, angle };
  double* dev_out[5];
  for (int k = 0; k < 5; ++k) dev_out[k] = host_out[k] ? dout + (size_t)k * n : nullptr;
  const int block = 128, grid = grid_for(n, block);
  k_triangulate<<<grid, block>>>(p1, p2, C1[0], C1[1], C1[2], C2[0], C2[1], C2[2], bl * bl, n, dx1, dx2, dX,
                                 dev_out[0], dev_out[1], dev_out[2], dev_out[3], dev_out[4]);
  MM_LAUNCH_CHECK();
  MM_CUDA(cudaMemcpy(X, dX, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 5; ++k) if (host_out[k]) MM_CUDA(cudaMemcpy(host_out[k], dev_out[k], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return MM_OK;
}

int mm_tri_angles(const double* P1, const double* P2, int64_t n, const double* X, double* angle) {
  if (n < 0 || !P1 || !P2 || (n > 0 && (!X || !angle))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc != MM_OK) return rc;
  if (n == 0) return MM_OK;
  double C1[3], C2[3];
  camera_center(P1, C1); camera_center(P2, C2);
  const double bl = sqrt((C1[0] - C2[0]) * (C1[0] - C2[0]) + (C1[1] - C2[1]) * (C1[1] - C2[1]) + (C1[2] - C2[2]) * (C1[2] - C2[2]));
  DevBuf<double> buf; MM_CUDA(buf.alloc((size_t)n * 4));
  MM_CUDA(cudaMemcpy(buf.p, X, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  k_tri_angles<<<grid_for(n, 128), 128>>>(C1[0], C1[1], C1[2], C2[0], C2[1], C2[2], bl * bl, n, buf.p, buf.p + 3 * n);
  MM_LAUNCH_CHECK();
  MM_CUDA(cudaMemcpy(angle, buf.p + 3 * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return MM_OK;
}

int mm_reproj_errors(const double* P, int64_t n, const double* x2d, const double* X, double* err, double* depth) {
  if (n < 0 || !P || (n > 0 && (!X || (err && !x2d)))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc != MM_OK) return rc;
  if (n == 0 || (!err && !depth)) return MM_OK;
  Proj p; memcpy(p.P, P, sizeof p.P);
  DevBuf<double> buf; MM_CUDA(buf.alloc((size_t)n * 7));
  double* dx = buf.p; double* dX = dx + 2 * n; double* de = dX + 3 * n; double* dd = de + n;
  if (err) MM_CUDA(cudaMemcpy(dx, x2d, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice));
  MM_CUDA(cudaMemcpy(dX, X, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  k_reproj<<<grid_for(n, 128), 128>>>(p, n, dx, dX, err ? de : nullptr, depth ? dd : nullptr);
  MM_LAUNCH_CHECK();
  if (err) MM_CUDA(cudaMemcpy(err, de, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  if (depth) MM_CUDA(cudaMemcpy(depth, dd, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return MM_OK;
}

int mm_ransac_score(int32_t kind, const double* models, int32_t n_models, int64_t n, const double* x, const double* y, double threshold,
                    int32_t* num_inliers, double* residual_sum, int32_t* best, double* best_residuals, uint8_t* best_mask) {
  if (kind < 0 || kind > 2 || n_models < 0 || n < 0 || (n_models > 0 && !models) || (n > 0 && (!x || !y))) { set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  int rc = ensure_device(); if (rc != MM_OK) return rc;
  if (best) *best = -1;
  if (n_models == 0) return MM_OK;
  const int msize = kind == MM_RANSAC_P3P ? 12 : 9, xd = 2, yd = kind == MM_RANSAC_P3P ? 3 : 2;
  DevBuf<double> d_m, d_x, d_y, d_sum, d_res; DevBuf<int> d_cnt; DevBuf<unsigned char> d_mask;
  MM_CUDA(d_m.alloc((size_t)n_models * msize)); MM_CUDA(d_x.alloc((size_t)n * xd)); MM_CUDA(d_y.alloc((size_t)n * yd));
  MM_CUDA(d_sum.alloc((size_t)n_models)); MM_CUDA(d_cnt.alloc((size_t)n_models));
  MM_CUDA(cudaMemcpyAsync(d_m.p, models, sizeof(double) * (size_t)n_models * msize, cudaMemcpyHostToDevice, nullptr));
  if (n) { MM_CUDA(cudaMemcpyAsync(d_x.p, x, sizeof(double) * (size_t)n * xd, cudaMemcpyHostToDevice, nullptr));
           MM_CUDA(cudaMemcpyAsync(d_y.p, y, sizeof(double) * (size_t)n * yd, cudaMemcpyHostToDevice, nullptr)); }
  k_ransac_score<<<n_models, 256>>>(kind, msize, d_m.p, n, d_x.p, d_y.p, threshold, d_cnt.p, d_sum.p);
  MM_LAUNCH_CHECK();
  std::vector<int> h_cnt((size_t)n_models); std::vector<double> h_sum((size_t)n_models);
  MM_CUDA(cudaMemcpy(h_cnt.data(), d_cnt.p, sizeof(int) * (size_t)n_models, cudaMemcpyDeviceToHost));
  MM_CUDA(cudaMemcpy(h_sum.data(), d_sum.p, sizeof(double) * (size_t)n_models, cudaMemcpyDeviceToHost));
  // util/estimation.cc:118-126: more inliers wins, ties go to the smaller residual sum (first hypothesis wins exact ties)
  int b = -1;
  for (int h = 0; h < n_models; ++h) {
    if (num_inliers) num_inliers[h] = h_cnt[h];
    if (residual_sum) residual_sum[h] = h_sum[h];
    if (b < 0 || h_cnt[h] > h_cnt[b] || (h_cnt[h] == h_cnt[b] && h_sum[h] < h_sum[b])) b = h;
  }
  if (best) *best = b;
  if (b >= 0 && n > 0 && (best_residuals || best_mask)) {
    MM_CUDA(d_res.alloc((size_t)n)); MM_CUDA(d_mask.alloc((size_t)n));
    k_ransac_residuals<<<grid_for(n, 256), 256>>>(kind, msize, d_m.p + (size_t)b * msize, n, d_x.p, d_y.p, threshold, d_res.p, d_mask.p);
    MM_LAUNCH_CHECK();
    if (best_residuals) MM_CUDA(cudaMemcpy(best_residuals, d_res.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    if (best_mask) MM_CUDA(cudaMemcpy(best_mask, d_mask.p, (size_t)n, cudaMemcpyDeviceToHost));
  }
  return MM_OK;
}

}  // extern "C"
