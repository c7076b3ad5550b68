// Drop-in bodies for src/base3d/triangulation.cc (:12-147) and the two per-point filters of
// src/base3d/projection.cc (:107-149) over the C ABI.  Include the reference's own
// base3d/triangulation.h / base3d/projection.h in front of this file (signatures unchanged).
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

Eigen::Vector3d triangulate_point(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                  const Eigen::Vector2d& point1, const Eigen::Vector2d& point2) {
  return triangulate_points(proj_matrix1, proj_matrix2, std::vector<Eigen::Vector2d>(1, point1), std::vector<Eigen::Vector2d>(1, point2))[0];
}

std::vector<double> calc_tri_angles(const Eigen::Matrix<double, 3, 4>& proj_matrix1, const Eigen::Matrix<double, 3, 4>& proj_matrix2,
                                    const std::vector<Eigen::Vector3d>& points3D) {
  // the angle only needs the camera centres and the points: reuse the fused kernel through its reproj/depth-free form
  const size_t n = points3D.size();
  double P1[12], P2[12]; flat_proj(proj_matrix1, P1); flat_proj(proj_matrix2, P2);
  // project the given points into both views so that the DLT reproduces them, then read the angle output
  std::vector<double> x1(2 * n), x2(2 * n), X(3 * n), ang(n);
  for (size_t i = 0; i < n; ++i) {
    double q1[3], q2[3];
    for (int r = 0; r < 3; ++r) {
      q1[r] = P1[4 * r] * points3D[i](0) + P1[4 * r + 1] * points3D[i](1) + P1[4 * r + 2] * points3D[i](2) + P1[4 * r + 3];
      q2[r] = P2[4 * r] * points3D[i](0) + P2[4 * r + 1] * points3D[i](1) + P2[4 * r + 2] * points3D[i](2) + P2[4 * r + 3];
    }
    x1[2 * i] = q1[0] / q1[2]; x1[2 * i + 1] = q1[1] / q1[2]; x2[2 * i] = q2[0] / q2[2]; x2[2 * i + 1] = q2[1] / q2[2];
  }
  check(mm_triangulate_two_view(P1, P2, (int64_t)n, x1.data(), x2.data(), X.data(), nullptr, nullptr, nullptr, nullptr, ang.data()));
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

double calc_depth(const Eigen::Matrix<double, 3, 4>& proj_matrix, const Eigen::Vector3d& point3D) {
  double P[12]; flat_proj(proj_matrix, P);
  const double X[3] = { point3D(0), point3D(1), point3D(2) }; double d = 0.0;
  check(mm_reproj_errors(P, 1, nullptr, X, nullptr, &d));
  return d;
}
