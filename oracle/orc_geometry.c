/*
 * oracle/orc_geometry.c — CPU restatement of two-view triangulation and its filters,
 * TEST INFRASTRUCTURE ONLY (see orc.h).
 *
 * Follows (mavmap/mavmap):
 *   triangulate_point      src/base3d/triangulation.cc:12-50  (6x4 DLT, right singular
 *                          vector of the smallest singular value, dehomogenise)
 *   triangulate_points     src/base3d/triangulation.cc:53-74
 *   calc_tri_angles        src/base3d/triangulation.cc:101-147 (law of cosines, NaN -> 0)
 *   calc_reproj_errors     src/base3d/projection.cc:107-130
 *   calc_depth             src/base3d/projection.cc:133-149
 *   invert_proj_matrix     src/base3d/projection.cc:82-88
 * The SVD lives in the absent dependency Eigen 3.2.0 (JacobiSVD, triangulation.cc:40);
 * restated here as a one-sided (Hestenes) Jacobi SVD, which like Eigen's two-sided
 * Jacobi resolves the smallest singular vector to full relative accuracy.
 */
#include <math.h>
#include <string.h>
#include "orc.h"

/* right singular vector of the smallest singular value of a 6x4 matrix (row-major) */
static void null_vector_6x4(const double* A_in, double* h) {
  double A[6][4], V[4][4];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 4; ++c) A[r][c] = A_in[4*r + c];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) V[r][c] = r == c ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 4; ++q) {
      double app = 0, aqq = 0, apq = 0;
      for (int r = 0; r < 6; ++r) { app += A[r][p]*A[r][p]; aqq += A[r][q]*A[r][q]; apq += A[r][p]*A[r][q]; }
      if (apq == 0.0) continue;
      const double denom = sqrt(app * aqq);
      if (denom == 0.0 || fabs(apq) <= 1e-300) continue;
      if (fabs(apq) / denom > off) off = fabs(apq) / denom;
      const double zeta = (aqq - app) / (2.0 * apq);
      const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
      for (int r = 0; r < 6; ++r) { const double x = A[r][p], y = A[r][q]; A[r][p] = c*x - s*y; A[r][q] = s*x + c*y; }
      for (int r = 0; r < 4; ++r) { const double x = V[r][p], y = V[r][q]; V[r][p] = c*x - s*y; V[r][q] = s*x + c*y; }
    }
    if (off < 1e-17) break;
  }
  int best = 0; double bn = INFINITY;
  for (int c = 0; c < 4; ++c) { double n = 0; for (int r = 0; r < 6; ++r) n += A[r][c]*A[r][c]; if (n < bn) { bn = n; best = c; } }
  for (int r = 0; r < 4; ++r) h[r] = V[r][best];
}

static void camera_center(const double* P, double* C) {
  /* -M^-1 p4, i.e. the translation column of invert_proj_matrix (projection.cc:82-88) */
  const double a = P[0], b = P[1], c = P[2], d = P[4], e = P[5], f = P[6], g = P[8], h = P[9], i = P[10];
  const double A = e*i - f*h, B = -(d*i - f*g), Cc = d*h - e*g;
  const double det = a*A + b*B + c*Cc;
  const double inv[9] = { A/det, (c*h - b*i)/det, (b*f - c*e)/det,
                          B/det, (a*i - c*g)/det, (c*d - a*f)/det,
                          Cc/det, (b*g - a*h)/det, (a*e - b*d)/det };
  const double t[3] = { P[3], P[7], P[11] };
  for (int r = 0; r < 3; ++r) C[r] = -(inv[3*r]*t[0] + inv[3*r+1]*t[1] + inv[3*r+2]*t[2]);
}

static double reproj_err(const double* P, const double* X, const double* x) {
  double q[3];
  for (int r = 0; r < 3; ++r) q[r] = P[4*r]*X[0] + P[4*r+1]*X[1] + P[4*r+2]*X[2] + P[4*r+3];
  const double dx = q[0] / q[2] - x[0], dy = q[1] / q[2] - x[1];
  return sqrt(dx*dx + dy*dy);
}

static double depth_of(const double* P, const double* X) {
  const double w = P[8]*X[0] + P[9]*X[1] + P[10]*X[2] + P[11];
  const double mx = P[2], my = P[6], mz = P[10];      /* proj_matrix(0..2, 2) — projection.cc:143-145 */
  return w * sqrt(mx*mx + my*my + mz*mz);
}

int orc_triangulate_two_view(const double* P1, const double* P2, int64_t n,
                             const double* x1, const double* x2, double* X,
                             double* reproj1, double* reproj2,
                             double* depth1, double* depth2, double* angle) {
  if (n < 0 || !P1 || !P2 || (n > 0 && (!x1 || !x2 || !X))) return MM_ERR_INVALID_ARG;
  double C1[3], C2[3];
  camera_center(P1, C1); camera_center(P2, C2);
  const double bl = sqrt((C1[0]-C2[0])*(C1[0]-C2[0]) + (C1[1]-C2[1])*(C1[1]-C2[1]) + (C1[2]-C2[2])*(C1[2]-C2[2]));
  const double baseline2 = bl * bl;
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double A[24], h[4];
    const double xa = x1[2*i], ya = x1[2*i+1], xb = x2[2*i], yb = x2[2*i+1];
    for (int k = 0; k < 4; ++k) {       /* triangulation.cc:26-35 */
      A[0*4+k] = xa * P1[8+k] - P1[k];
      A[1*4+k] = ya * P1[8+k] - P1[4+k];
      A[2*4+k] = xa * P1[4+k] - ya * P1[k];
      A[3*4+k] = xb * P2[8+k] - P2[k];
      A[4*4+k] = yb * P2[8+k] - P2[4+k];
      A[5*4+k] = xb * P2[4+k] - yb * P2[k];
    }
    null_vector_6x4(A, h);
    double* Xi = X + 3*i;
    Xi[0] = h[0] / h[3]; Xi[1] = h[1] / h[3]; Xi[2] = h[2] / h[3];
    if (reproj1) reproj1[i] = reproj_err(P1, Xi, x1 + 2*i);
    if (reproj2) reproj2[i] = reproj_err(P2, Xi, x2 + 2*i);
    if (depth1) depth1[i] = depth_of(P1, Xi);
    if (depth2) depth2[i] = depth_of(P2, Xi);
    if (angle) {
      const double r1 = sqrt((Xi[0]-C1[0])*(Xi[0]-C1[0]) + (Xi[1]-C1[1])*(Xi[1]-C1[1]) + (Xi[2]-C1[2])*(Xi[2]-C1[2]));
      const double r2 = sqrt((Xi[0]-C2[0])*(Xi[0]-C2[0]) + (Xi[1]-C2[1])*(Xi[1]-C2[1]) + (Xi[2]-C2[2])*(Xi[2]-C2[2]));
      const double a = acos((r1*r1 + r2*r2 - baseline2) / (2.0 * r1 * r2));
      angle[i] = isnan(a) ? 0.0 : a;
    }
  }
  return MM_OK;
}


/* calc_tri_angles, triangulation.cc:101-147, for given 3-D points */
int orc_tri_angles(const double* P1, const double* P2, int64_t n, const double* X, double* angle) {
  if (n < 0 || !P1 || !P2 || (n > 0 && (!X || !angle))) return MM_ERR_INVALID_ARG;
  double C1[3], C2[3];
  camera_center(P1, C1); camera_center(P2, C2);                 /* :106-113 translation column of the inverted projection matrices */
  const double bl = sqrt((C1[0]-C2[0])*(C1[0]-C2[0]) + (C1[1]-C2[1])*(C1[1]-C2[1]) + (C1[2]-C2[2])*(C1[2]-C2[2]));   /* :116 */
  const double baseline2 = bl * bl;
  for (int64_t i = 0; i < n; ++i) {
    const double* Xi = X + 3*i;
    const double r1 = sqrt((Xi[0]-C1[0])*(Xi[0]-C1[0]) + (Xi[1]-C1[1])*(Xi[1]-C1[1]) + (Xi[2]-C1[2])*(Xi[2]-C1[2]));      /* :127-128 */
    const double r2 = sqrt((Xi[0]-C2[0])*(Xi[0]-C2[0]) + (Xi[1]-C2[1])*(Xi[1]-C2[1]) + (Xi[2]-C2[2])*(Xi[2]-C2[2]));
    const double a = acos((r1*r1 + r2*r2 - baseline2) / (2.0 * r1 * r2));                                          /* :132-133 */
    angle[i] = isnan(a) ? 0.0 : a;                                                                                  /* :134-140 */
  }
  return MM_OK;
}

/* ---- RANSAC hypothesis scoring: util/estimation.cc:83-126 with the residuals of p3p.cc:172-199 (kind 0),
 * projective_transform.cc:48-74 (kind 1) and essential_matrix.cc:131-162 (kind 2).  Row-major models. */
static double ransac_residual(int kind, const double* m, const double* x, const double* y, int64_t i) {
  if (kind == 0) {
    const double X0 = y[3*i], X1 = y[3*i+1], X2 = y[3*i+2];
    double p0 = (m[0]*X0 + m[1]*X1 + m[2]*X2) + m[3];
    double p1 = (m[4]*X0 + m[5]*X1 + m[6]*X2) + m[7];
    const double p2 = (m[8]*X0 + m[9]*X1 + m[10]*X2) + m[11];
    p0 /= p2; p1 /= p2;
    const double dx = p0 - x[2*i], dy = p1 - x[2*i+1];
    return sqrt(dx*dx + dy*dy);
  } else if (kind == 1) {
    const double s0 = x[2*i], s1 = x[2*i+1];
    const double t0 = m[0]*s0 + m[1]*s1 + m[2]*1.0, t1 = m[3]*s0 + m[4]*s1 + m[5]*1.0, t2 = m[6]*s0 + m[7]*s1 + m[8]*1.0;
    const double dx = t0 / t2 - y[2*i], dy = t1 / t2 - y[2*i+1];
    return sqrt(dx*dx + dy*dy);
  } else {
    const double a0 = x[2*i], a1 = x[2*i+1], b0 = y[2*i], b1 = y[2*i+1];
    const double e0 = m[0]*a0 + m[1]*a1 + m[2]*1.0, e1 = m[3]*a0 + m[4]*a1 + m[5]*1.0, e2 = m[6]*a0 + m[7]*a1 + m[8]*1.0;
    const double f0 = m[0]*b0 + m[3]*b1 + m[6]*1.0, f1 = m[1]*b0 + m[4]*b1 + m[7]*1.0;
    const double num = b0*e0 + b1*e1 + 1.0*e2;
    return num / sqrt(e0*e0 + e1*e1 + f0*f0 + f1*f1);
  }
}
int orc_ransac_score(int kind, const double* models, int n_models, int64_t n, const double* x, const double* y, double threshold,
                     int32_t* num_inliers, double* residual_sum, int32_t* best, double* best_residuals, uint8_t* best_mask) {
  const int msize = kind == 0 ? 12 : 9; int b = -1; int bc = 0; double bs = 0.0;
  for (int h = 0; h < n_models; ++h) {
    int c = 0; double s = 0.0;
    for (int64_t i = 0; i < n; ++i) { const double a = fabs(ransac_residual(kind, models + (size_t)h*msize, x, y, i)); if (a <= threshold) { ++c; s += a; } }
    if (num_inliers) num_inliers[h] = c;
    if (residual_sum) residual_sum[h] = s;
    if (b < 0 || c > bc || (c == bc && s < bs)) { b = h; bc = c; bs = s; }
  }
  if (best) *best = b;
  if (b >= 0) for (int64_t i = 0; i < n; ++i) {
    const double r = ransac_residual(kind, models + (size_t)b*msize, x, y, i);
    if (best_residuals) best_residuals[i] = r;
    if (best_mask) best_mask[i] = fabs(r) <= threshold;
  }
  return 0;
}
