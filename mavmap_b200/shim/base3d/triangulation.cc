// Drop-in bodies for src/base3d/triangulation.cc (:12-147) and the two per-point filters of
// src/base3d/projection.cc (:107-149) over the C ABI.  Include the reference's own
// base3d/triangulation.h / base3d/projection.h in front of this file (signatures unchanged).
#include <cmath>
#include <vector>
#include <stdexcept>
#include <string>
#include <Eigen/Core>
#include "mavmap_b200.h"

namespace {
void flat_proj(const Eigen::Matrix<double, 3, 4>& P, double* out) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out[4 * r + c] = P(r, c); }
void check(int rc) { if (rc != MM_OK) throw std::runtime_error(std::string("mavmap_b200: ") + mm_last_error()); }
}

std::vector<Eigen::Vector3d> triangulate_points(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                                const std::vector<Eigen::Vector2d>& points1, const std::vector<Eigen::Vector2d>& points2) {
  const size_t n = points1.size();
  double P1[12], P2[12]; flat_proj(proj_matrix1, P1); flat_proj(proj_matrix2, P2);
  std::vector<double> x1(2 * n), x2(2 * n), X(3 * n);
  for (size_t i = 0; i < n; ++i) { x1[2 * i] = points1[i](0); x1[2 * i + 1] = points1[i](1); x2[2 * i] = points2[i](0); x2[2 * i + 1] = points2[i](1); }
  check(mm_triangulate_two_view(P1, P2, (int64_t)n, x1.data(), x2.data(), X.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
  std::vector<Eigen::Vector3d> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = Eigen::Vector3d(X[3 * i], X[3 * i + 1], X[3 * i + 2]);
  return out;
}

// Single correspondence (mapper.cc:499 triangulates one ground-control point at a time): a device round trip would cost far
// more than the 6 x 4 null-vector problem itself, so this one runs on the host - the same one-sided Jacobi sweep as the
// batched kernel (csrc/geometry.cu: null_vector_6x4), on the DLT rows of triangulation.cc:26-35.
Eigen::Vector3d triangulate_point(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                  const Eigen::Vector2d& point1, const Eigen::Vector2d& point2) {
  double P1[12], P2[12]; flat_proj(proj_matrix1, P1); flat_proj(proj_matrix2, P2);
  const double xa = point1(0), ya = point1(1), xb = point2(0), yb = point2(1);
  double A[6][4], V[4][4];
  for (int k = 0; k < 4; ++k) {
    A[0][k] = xa * P1[8 + k] - P1[k];     A[1][k] = ya * P1[8 + k] - P1[4 + k];  A[2][k] = xa * P1[4 + k] - ya * P1[k];
    A[3][k] = xb * P2[8 + k] - P2[k];     A[4][k] = yb * P2[8 + k] - P2[4 + k];  A[5][k] = xb * P2[4 + k] - yb * P2[k];
  }
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) V[r][c] = r == c ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 4; ++q) {
      double app = 0, aqq = 0, apq = 0;
      for (int r = 0; r < 6; ++r) { app += A[r][p] * A[r][p]; aqq += A[r][q] * A[r][q]; apq += A[r][p] * A[r][q]; }
      const double denom = std::sqrt(app * aqq);
      if (apq != 0.0 && denom != 0.0 && std::fabs(apq) > 1e-300) {
        off = std::fmax(off, std::fabs(apq) / denom);
        const double zeta = (aqq - app) / (2.0 * apq);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + tt * tt), s = c * tt;
        for (int r = 0; r < 6; ++r) { const double x = A[r][p], y = A[r][q]; A[r][p] = c * x - s * y; A[r][q] = s * x + c * y; }
        for (int r = 0; r < 4; ++r) { const double x = V[r][p], y = V[r][q]; V[r][p] = c * x - s * y; V[r][q] = s * x + c * y; }
      }
    }
    if (off < 1e-17) break;
  }
  double bn = INFINITY; int best = 0;
  for (int c = 0; c < 4; ++c) { double nn = 0; for (int r = 0; r < 6; ++r) nn += A[r][c] * A[r][c]; if (nn < bn) { bn = nn; best = c; } }
  return Eigen::Vector3d(V[0][best] / V[3][best], V[1][best] / V[3][best], V[2][best] / V[3][best]);      // triangulation.cc:44-46
}

// Nx2 / Nx3 matrix overload (triangulation.cc:77-98); forwards to the vector version
Eigen::Matrix<double, Eigen::Dynamic, 3> triangulate_points(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                                            const Eigen::Matrix<double, Eigen::Dynamic, 2>& points1,
                                                            const Eigen::Matrix<double, Eigen::Dynamic, 2>& points2) {
  const size_t n = (size_t)points1.rows();
  std::vector<Eigen::Vector2d> p1(n), p2(n);
  for (size_t i = 0; i < n; ++i) { p1[i] = Eigen::Vector2d(points1(i, 0), points1(i, 1)); p2[i] = Eigen::Vector2d(points2(i, 0), points2(i, 1)); }
  const std::vector<Eigen::Vector3d> X = triangulate_points(proj_matrix1, proj_matrix2, p1, p2);
  Eigen::Matrix<double, Eigen::Dynamic, 3> out(n, 3);
  for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) out(i, k) = X[i](k);
  return out;
}

// calc_tri_angles (triangulation.cc:101-147) uses the GIVEN 3-D points: batched on the device
std::vector<double> calc_tri_angles(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                    const std::vector<Eigen::Vector3d>& points3D) {
  const size_t n = points3D.size();
  double P1[12], P2[12]; flat_proj(proj_matrix1, P1); flat_proj(proj_matrix2, P2);
  std::vector<double> X(3 * n), ang(n);
  for (size_t i = 0; i < n; ++i) { X[3 * i] = points3D[i](0); X[3 * i + 1] = points3D[i](1); X[3 * i + 2] = points3D[i](2); }
  check(mm_tri_angles(P1, P2, (int64_t)n, X.data(), ang.data()));
  return ang;
}

std::vector<double> calc_reproj_errors(const std::vector<Eigen::Vector2d>& points2D, const std::vector<Eigen::Vector3d>& points3D,
                                       const Eigen::Matrix<double, 3, 4>& proj_matrix) {
  const size_t n = points3D.size();
  double P[12]; flat_proj(proj_matrix, P);
  std::vector<double> x(2 * n), X(3 * n), err(n);
  for (size_t i = 0; i < n; ++i) { x[2 * i] = points2D[i](0); x[2 * i + 1] = points2D[i](1); X[3 * i] = points3D[i](0); X[3 * i + 1] = points3D[i](1); X[3 * i + 2] = points3D[i](2); }
  check(mm_reproj_errors(P, (int64_t)n, x.data(), X.data(), err.data(), nullptr));
  return err;
}

// calc_depth (projection.cc:133-149) is called once per POINT inside the mapper's loops (sequential_mapper.cc:341, :801,
// essential_matrix.cc:250-252): eleven flops, evaluated in place.  (The batched form is mm_reproj_errors.)
double calc_depth(const Eigen::Matrix<double, 3, 4>& proj_matrix, const Eigen::Vector3d& point3D) {
  const double w = proj_matrix(2, 0) * point3D(0) + proj_matrix(2, 1) * point3D(1) + proj_matrix(2, 2) * point3D(2) + proj_matrix(2, 3);
  const double mx = proj_matrix(0, 2), my = proj_matrix(1, 2), mz = proj_matrix(2, 2);
  return w * std::sqrt(mx * mx + my * my + mz * mz);
}
