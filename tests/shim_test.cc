// Exercises the C++ drop-in shim (mavmap_b200/shim) exactly as sequential_mapper.cc calls the reference:
// bundle_adjustment(), pose_refinement(), match_brute_force(), triangulate_points(), calc_*.
// Prints one line per check; exit code 0 iff all pass.  Built by tests/test_shim.py.
#include <cmath>
#include <cstdio>
#include <random>
#include <set>
#include <stdexcept>
#include "base3d/bundle_adjustment.h"
#include <opencv2/core/core.hpp>

void match_brute_force(const std::vector<cv::KeyPoint>&, const cv::Mat&, const std::vector<cv::KeyPoint>&, const cv::Mat&,
                       std::vector<cv::DMatch>&, const bool ratio_test = true, const double max_ratio = 0.6,
                       const double max_distance = -1, const int norm_type = cv::NORM_L2);
std::vector<Eigen::Vector3d> triangulate_points(const Eigen::Matrix<double, 3, 4>&, const Eigen::Matrix<double, 3, 4>&,
                                                const std::vector<Eigen::Vector2d>&, const std::vector<Eigen::Vector2d>&);
std::vector<double> calc_reproj_errors(const std::vector<Eigen::Vector2d>&, const std::vector<Eigen::Vector3d>&, const Eigen::Matrix<double, 3, 4>&);
double calc_depth(const Eigen::Matrix<double, 3, 4>&, const Eigen::Vector3d&);
Eigen::Vector3d triangulate_point(const Eigen::Matrix<double, 3, 4>&, const Eigen::Matrix<double, 3, 4>&, const Eigen::Vector2d&, const Eigen::Vector2d&);
Eigen::Matrix<double, Eigen::Dynamic, 3> triangulate_points(const Eigen::Matrix<double, 3, 4>&, const Eigen::Matrix<double, 3, 4>&,
                                                            const Eigen::Matrix<double, Eigen::Dynamic, 2>&, const Eigen::Matrix<double, Eigen::Dynamic, 2>&);
std::vector<double> calc_tri_angles(const Eigen::Matrix<double, 3, 4>&, const Eigen::Matrix<double, 3, 4>&, const std::vector<Eigen::Vector3d>&);

static int fails = 0;
#define CHECK(cond, what) do { const bool ok_ = (cond); std::printf("%s %s\n", ok_ ? "ok  " : "FAIL", what); if (!ok_) ++fails; } while (0)

int main() {
  std::mt19937 rng(7); std::normal_distribution<double> N(0.0, 1.0); std::uniform_real_distribution<double> U(-1.0, 1.0);
  // ---- a 4-image scene, pinhole, exact observations + 0.3 px noise
  FeatureManager fm;
  const size_t cam = fm.add_camera({1000.0, 1000.0, 640.0, 480.0, 1});
  const int n_pt = 80, n_img = 4;
  std::vector<Eigen::Vector3d> X(n_pt);
  for (auto& x : X) x = Eigen::Vector3d(2 * U(rng), 2 * U(rng), 10 + 2 * U(rng));
  std::vector<size_t> ids;
  for (int i = 0; i < n_img; ++i) {
    const double tx = 0.5 * i;
    std::vector<Eigen::Vector2d> uv(n_pt);
    for (int p = 0; p < n_pt; ++p) uv[p] = Eigen::Vector2d(1000 * (X[p](0) + tx) / X[p](2) + 640 + 0.3 * N(rng), 1000 * X[p](1) / X[p](2) + 480 + 0.3 * N(rng));
    const size_t id = fm.add_image(cam, uv);
    fm.rvecs[id] = Eigen::Vector3d(0.003 * N(rng), 0.003 * N(rng), 0.003 * N(rng));
    fm.tvecs[id] = Eigen::Vector3d(tx + 0.02 * N(rng), 0.02 * N(rng), 0.02 * N(rng));
    ids.push_back(id);
  }
  for (int p = 0; p < n_pt; ++p) {
    const size_t pid = fm.add_point3D();
    fm.points3D[pid] = Eigen::Vector3d(X[p](0) + 0.05 * N(rng), X[p](1) + 0.05 * N(rng), X[p](2) + 0.05 * N(rng));
    for (size_t id : ids) fm.point2D_to_point3D[fm.image_to_points2D[id][p]] = pid;
  }
  BundleAdjustmentOptions opt; opt.print_summary = false; opt.update_point3D_errors = true; opt.max_num_iterations = 30;
  std::unordered_map<size_t, double> errs;
  const Eigen::Vector3d r0 = fm.rvecs[ids[0]];
  const double cost = bundle_adjustment(fm, {ids[2], ids[3]}, {ids[0]}, {ids[1]}, opt, errs);
  CHECK(cost > 0.05 && cost < 0.6, "bundle_adjustment returns sqrt(final_cost/num_residuals) ~ noise level");
  CHECK(fm.rvecs[ids[0]](0) == r0(0) && fm.rvecs[ids[0]](1) == r0(1), "FIXED pose untouched");
  CHECK(errs.size() == (size_t)n_pt, "point3D_errors filled for every point in the problem");
  bool threw = false;
  try { bundle_adjustment(fm, {ids[1], ids[2], ids[3]}, {ids[0]}, {}, opt, errs); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw, "std::invalid_argument for < 7 fixed parameters (bundle_adjustment.cc:459-466)");
  opt.min_track_len = 1; threw = false;
  try { bundle_adjustment(fm, {ids[2], ids[3]}, {ids[0]}, {ids[1]}, opt, errs); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw, "std::invalid_argument for min_track_len < 2 (bundle_adjustment.cc:468-471)");
  // ---- constrain_rotation (bundle_adjustment.cc:390-446): the shim rotates the whole feature manager by M = R_FM' R_C, which turns
  // the first fixed image's rotation into R_FM R_C' R_FM, keeps every camera-frame point R X + t, and adds one residual per free image
  {
    opt.min_track_len = 2; opt.constrain_rotation = true; opt.constrain_rotation_weight = 5.0; opt.max_num_iterations = 8;
    std::unordered_map<size_t, Eigen::Vector3d> cons;
    for (size_t id : ids) cons[id] = Eigen::Vector3d(-fm.rvecs[id](0) + 0.001, -fm.rvecs[id](1), -fm.rvecs[id](2));   // ~ camera-to-world rotations
    cons[ids[0]] = Eigen::Vector3d(0.0, 0.0, 0.3);                                                                       // a real change of frame
    auto rot = [](const Eigen::Vector3d& r, const Eigen::Vector3d& x) {      // Rodrigues: R(r) x
      const double t = std::sqrt(r(0) * r(0) + r(1) * r(1) + r(2) * r(2));
      if (t < 1e-14) return Eigen::Vector3d(x(0), x(1), x(2));
      const double k0 = r(0) / t, k1 = r(1) / t, k2 = r(2) / t, c = std::cos(t), s_ = std::sin(t), d = (k0 * x(0) + k1 * x(1) + k2 * x(2)) * (1 - c);
      return Eigen::Vector3d(x(0) * c + (k1 * x(2) - k2 * x(1)) * s_ + k0 * d, x(1) * c + (k2 * x(0) - k0 * x(2)) * s_ + k1 * d, x(2) * c + (k0 * x(1) - k1 * x(0)) * s_ + k2 * d);
    };
    const size_t pid0 = fm.point2D_to_point3D[fm.image_to_points2D[ids[0]][0]];
    const Eigen::Vector3d before = rot(fm.rvecs[ids[0]], fm.points3D[pid0]);      // camera-frame direction of a point in the FIXED image
    const double c2 = bundle_adjustment(fm, {ids[2], ids[3]}, {ids[0]}, {ids[1]}, opt, errs, cons);
    const Eigen::Vector3d ez(0, 0, 1), rz = rot(fm.rvecs[ids[0]], ez);
    // R_FM ~ I here, so the fixed image's new rotation is ~ R_C' = rotation by -0.3 about z: it leaves e_z in place
    CHECK(std::isfinite(c2) && c2 > 0.0, "bundle_adjustment with constrain_rotation runs");
    CHECK(std::fabs(rz(2) - 1.0) < 1e-3 && std::fabs(std::fabs(fm.rvecs[ids[0]](2)) - 0.3) < 2e-2, "scene rotated into the frame of the constraints (R' = R M')");
    // the point moved with the scene (X' = M X) and the fixed pose is not optimised: same camera-frame coordinates up to the BA update of X
    const Eigen::Vector3d after = rot(fm.rvecs[ids[0]], fm.points3D[pid0]);
    CHECK(std::fabs(after(0) - before(0)) < 0.1 && std::fabs(after(1) - before(1)) < 0.1 && std::fabs(after(2) - before(2)) < 0.3, "camera-frame geometry kept by the pre-rotation");
    opt.constrain_rotation = false;
  }
  // ---- pose refinement
  std::vector<Eigen::Vector2d> p2; std::vector<Eigen::Vector3d> p3; std::vector<bool> mask;
  for (int p = 0; p < n_pt; ++p) { p2.push_back(Eigen::Vector2d(1000 * (X[p](0) + 0.7) / X[p](2) + 640, 1000 * (X[p](1) - 0.2) / X[p](2) + 480)); p3.push_back(X[p]); mask.push_back(p % 9 != 0); }
  Eigen::Vector3d rv(0.01, -0.01, 0.005), tv(0.65, -0.15, 0.05);
  std::vector<double> cp = {1000.0, 1000.0, 640.0, 480.0, 1};
  BundleAdjustmentOptions po; po.print_summary = false; po.function_tolerance = 1e-12;
  const double pc = pose_refinement(rv, tv, cp, p2, p3, mask, po);
  CHECK(pc < 1e-6 && std::fabs(tv(0) - 0.7) < 1e-6 && std::fabs(tv(1) + 0.2) < 1e-6 && std::fabs(rv(0)) < 1e-7, "pose_refinement recovers the exact pose");
  // ---- matching
  const int nd = 300, kd = 64;
  cv::Mat d1(nd, kd), d2(nd, kd);
  for (int i = 0; i < nd; ++i) for (int k = 0; k < kd; ++k) { d1.ptr<float>(i)[k] = (float)N(rng); d2.ptr<float>((i * 7) % nd)[k] = d1.ptr<float>(i)[k] + 0.05f * (float)N(rng); }
  std::vector<cv::KeyPoint> kp1(nd), kp2(nd); std::vector<cv::DMatch> matches;
  match_brute_force(kp1, d1, kp2, d2, matches, true, 0.9, -1, cv::NORM_L2);
  bool all_ok = matches.size() == (size_t)nd;
  for (const auto& m : matches) all_ok = all_ok && m.trainIdx == (m.queryIdx * 7) % nd && m.imgIdx == 0;
  CHECK(all_ok, "match_brute_force finds the planted permutation, ascending queryIdx");
  // ---- triangulation + filters
  Eigen::Matrix<double, 3, 4> P1, P2; for (int i = 0; i < 3; ++i) { P1(i, i) = 1; P2(i, i) = 1; } P2(0, 3) = -1.0;
  std::vector<Eigen::Vector2d> x1, x2;
  for (int p = 0; p < n_pt; ++p) { x1.push_back(Eigen::Vector2d(X[p](0) / X[p](2), X[p](1) / X[p](2))); x2.push_back(Eigen::Vector2d((X[p](0) - 1) / X[p](2), X[p](1) / X[p](2))); }
  const std::vector<Eigen::Vector3d> T = triangulate_points(P1, P2, x1, x2);
  double worst = 0; for (int p = 0; p < n_pt; ++p) for (int c = 0; c < 3; ++c) worst = std::fmax(worst, std::fabs(T[p](c) - X[p](c)));
  CHECK(worst < 1e-9, "triangulate_points recovers the points (triangulation_test.cc tolerance class)");
  const std::vector<double> re = calc_reproj_errors(x2, T, P2);
  double wre = 0; for (double e : re) wre = std::fmax(wre, e);
  CHECK(wre < 1e-10 && std::fabs(calc_depth(P2, T[0]) - X[0](2)) < 1e-9, "calc_reproj_errors / calc_depth");
  {
    // single-point triangulation runs on the host (mapper.cc:499), same algorithm as the batched kernel
    double w1 = 0; for (int p = 0; p < n_pt; ++p) { const Eigen::Vector3d t1 = triangulate_point(P1, P2, x1[p], x2[p]); for (int c = 0; c < 3; ++c) w1 = std::fmax(w1, std::fabs(t1(c) - T[p](c))); }
    CHECK(w1 < 1e-12, "triangulate_point (host) agrees with the batched kernel");
    Eigen::Matrix<double, Eigen::Dynamic, 2> m1(n_pt, 2), m2(n_pt, 2);
    for (int p = 0; p < n_pt; ++p) { m1(p, 0) = x1[p](0); m1(p, 1) = x1[p](1); m2(p, 0) = x2[p](0); m2(p, 1) = x2[p](1); }
    const Eigen::Matrix<double, Eigen::Dynamic, 3> TM = triangulate_points(P1, P2, m1, m2);
    double w2 = 0; for (int p = 0; p < n_pt; ++p) for (int c = 0; c < 3; ++c) w2 = std::fmax(w2, std::fabs(TM(p, c) - T[p](c)));
    CHECK(TM.rows() == (size_t)n_pt && w2 == 0.0, "triangulate_points, Nx2 matrix overload (triangulation.cc:77-98)");
    // calc_tri_angles on GIVEN points (triangulation.cc:101-147): camera centres (0,0,0) and (1,0,0); incl. a point on the
    // principal plane of view 1 (z = 0), which has no projection but a perfectly good ray angle, and a far point
    std::vector<Eigen::Vector3d> G = { Eigen::Vector3d(0.5, 0.0, 10.0), Eigen::Vector3d(0.5, 2.0, 0.0), Eigen::Vector3d(0.5, 0.0, 1e9), Eigen::Vector3d(3.0, -1.0, 4.0) };
    const std::vector<double> ang = calc_tri_angles(P1, P2, G);
    bool aok = ang.size() == G.size();
    for (size_t i = 0; i < G.size() && aok; ++i) {
      const double ax = G[i](0), ay = G[i](1), az = G[i](2), bx = ax - 1.0;
      const double ra = std::sqrt(ax * ax + ay * ay + az * az), rb = std::sqrt(bx * bx + ay * ay + az * az);
      double ref = std::acos((ra * ra + rb * rb - 1.0) / (2 * ra * rb)); if (std::isnan(ref)) ref = 0.0;
      aok = std::fabs(ang[i] - ref) < 1e-12;
    }
    CHECK(aok && ang[1] > 0.4, "calc_tri_angles uses the given points (valid on the principal plane, ~0 at infinity)");
    Eigen::Matrix<double, 3, 4> Pd; Pd(0, 0) = 0.6; Pd(0, 2) = 0.8; Pd(1, 1) = 1; Pd(2, 0) = -0.8; Pd(2, 2) = 0.6; Pd(2, 3) = 2.0;
    const Eigen::Vector3d q(1.0, 2.0, 3.0);
    CHECK(std::fabs(calc_depth(Pd, q) - (-0.8 * 1.0 + 0.6 * 3.0 + 2.0) * std::sqrt(0.8 * 0.8 + 0.6 * 0.6)) < 1e-15, "calc_depth: closed form of projection.cc:133-149 (host, no device round trip)");
  }
  // ---- a larger feature manager: what the marshalling costs next to the device call (printed with MM_SHIM_TIMING=1)
  {
    FeatureManager big;
    const size_t cam2 = big.add_camera({1000.0, 1000.0, 640.0, 480.0, 1});
    const int NI = 60, NP = 6000;
    std::vector<Eigen::Vector3d> XB(NP);
    for (auto& x : XB) x = Eigen::Vector3d(30 * U(rng) + 30, 4 * U(rng), 40 + 4 * U(rng));
    std::vector<size_t> bid;
    for (int i = 0; i < NI; ++i) {
      const double tx = -1.0 * i;
      std::vector<Eigen::Vector2d> uv(NP);
      for (int p = 0; p < NP; ++p) uv[p] = Eigen::Vector2d(1000 * (XB[p](0) + tx) / XB[p](2) + 640 + 0.3 * N(rng), 1000 * XB[p](1) / XB[p](2) + 480 + 0.3 * N(rng));
      const size_t id = big.add_image(cam2, uv);
      big.rvecs[id] = Eigen::Vector3d(0.002 * N(rng), 0.002 * N(rng), 0.002 * N(rng));
      big.tvecs[id] = Eigen::Vector3d(tx + 0.02 * N(rng), 0.02 * N(rng), 0.02 * N(rng));
      bid.push_back(id);
    }
    for (int p = 0; p < NP; ++p) {
      const size_t pid = big.add_point3D();
      big.points3D[pid] = Eigen::Vector3d(XB[p](0) + 0.05 * N(rng), XB[p](1) + 0.05 * N(rng), XB[p](2) + 0.05 * N(rng));
      for (int i = p % 7; i < NI; i += 7) big.point2D_to_point3D[big.image_to_points2D[bid[i]][p]] = pid;      // ~8 observations per point
    }
    BundleAdjustmentOptions bo; bo.print_summary = false; bo.max_num_iterations = 5;
    std::unordered_map<size_t, double> e2;
    std::vector<size_t> free_ids(bid.begin() + 2, bid.end());
    const double cb = bundle_adjustment(big, free_ids, {bid[0]}, {bid[1]}, bo, e2);
    CHECK(std::isfinite(cb) && cb < 1.0, "bundle_adjustment on a 60-image feature manager (timing: MM_SHIM_TIMING=1)");
  }
  std::printf("%d failure(s)\n", fails);
  return fails ? 1 : 0;
}
