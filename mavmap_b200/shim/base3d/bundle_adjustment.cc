// Drop-in body of bundle_adjustment() / pose_refinement() (reference: src/base3d/bundle_adjustment.cc).
// Host work only: which observations enter, in which order, and which parameter blocks are constant
// follow bundle_adjustment.cc:228-387, :459-471, :545-549 exactly; all arithmetic happens behind
// mm_ba_solve / mm_pose_refine (include/mavmap_b200.h).
#include "base3d/bundle_adjustment.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>

#include "mavmap_b200.h"

namespace {

int model_num_params(int code) { return mm_camera_model_num_params(code); }

void throw_for(int rc) {
  if (rc == MM_OK) return;
  const std::string msg = mm_last_error();
  if (rc == MM_ERR_INVALID_ARG || rc == MM_ERR_DATUM || rc == MM_ERR_MIN_TRACK_LEN) throw std::invalid_argument(msg);
  throw std::runtime_error("mavmap_b200: " + msg);
}

void print_report(const char* title, const mm_ba_summary& s, long long num_parameters) {   // _print_report (.cc:114-136)
  std::cout << title << std::endl << std::string(std::string(title).size(), '-') << std::endl;
  std::cout << std::right << std::setw(18) << "Residuals : " << std::left << s.num_residuals << std::endl;
  std::cout << std::right << std::setw(18) << "Parameters : " << std::left << num_parameters << std::endl;
  std::cout << std::right << std::setw(18) << "Iterations : " << std::left << s.num_successful_steps + s.num_unsuccessful_steps << std::endl;
  std::cout << std::right << std::setw(18) << "Initial cost : " << std::right << std::setprecision(6)
            << std::sqrt(s.initial_cost / s.num_residuals) << " [px]" << std::endl;
  std::cout << std::right << std::setw(18) << "Final cost : " << std::right << std::setprecision(6)
            << std::sqrt(s.final_cost / s.num_residuals) << " [px]" << std::endl << std::endl;
}

mm_ba_options to_options(const BundleAdjustmentOptions& o) {
  mm_ba_options c; mm_ba_options_default(&c);
  c.max_num_iterations = (int32_t)o.max_num_iterations;
  c.function_tolerance = o.function_tolerance;
  c.gradient_tolerance = o.gradient_tolerance;
  c.loss_scale = o.loss_scale_factor;
  c.print_progress = o.print_progress ? 1 : 0;
  return c;
}

}  // namespace

// angle-axis <-> rotation matrix (row-major), the conventions of Eigen::AngleAxisd used by the reference (util/math.h)
static void rvec_to_matrix(const Eigen::Vector3d& r, double* R) {
  const double x = r(0), y = r(1), z = r(2), t2 = x * x + y * y + z * z;
  if (!(t2 > 0.0)) { for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0; return; }
  const double t = std::sqrt(t2), s = std::sin(t), c = std::cos(t), kx = x / t, ky = y / t, kz = z / t, v = 1.0 - c;
  R[0] = c + kx * kx * v;      R[1] = kx * ky * v - kz * s; R[2] = kx * kz * v + ky * s;
  R[3] = ky * kx * v + kz * s; R[4] = c + ky * ky * v;      R[5] = ky * kz * v - kx * s;
  R[6] = kz * kx * v - ky * s; R[7] = kz * ky * v + kx * s; R[8] = c + kz * kz * v;
}
static void matrix_to_rvec(const double* R, Eigen::Vector3d& r) {
  // via the unit quaternion (as Eigen::AngleAxis(Matrix3) does): angle in [0, pi]
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) { const double s = std::sqrt(tr + 1.0) * 2.0; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  if (q[0] < 0.0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
  const double n = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-300) { r(0) = r(1) = r(2) = 0.0; return; }
  const double angle = 2.0 * std::atan2(n, q[0]);
  r(0) = angle * q[1] / n; r(1) = angle * q[2] / n; r(2) = angle * q[3] / n;
}


double pose_refinement(Eigen::Vector3d& rvec, Eigen::Vector3d& tvec, std::vector<double>& camera_params,
                       const std::vector<Eigen::Vector2d>& points2D, std::vector<Eigen::Vector3d>& points3D,
                       const std::vector<bool>& inlier_mask, const BundleAdjustmentOptions& options) {
  const int code = (int)camera_params.back();                               // .cc:172
  if (model_num_params(code) < 0) throw std::invalid_argument("unknown camera model code");
  const size_t n = points2D.size();
  std::vector<double> p2(2 * n), p3(3 * n); std::vector<uint8_t> mask(n);
  for (size_t i = 0; i < n; ++i) {
    p2[2 * i] = points2D[i](0); p2[2 * i + 1] = points2D[i](1);
    p3[3 * i] = points3D[i](0); p3[3 * i + 1] = points3D[i](1); p3[3 * i + 2] = points3D[i](2);
    mask[i] = inlier_mask[i] ? 1 : 0;
  }
  double rv[3] = { rvec(0), rvec(1), rvec(2) }, tv[3] = { tvec(0), tvec(1), tvec(2) }, ret = 0.0;
  mm_ba_options c = to_options(options); mm_ba_summary s;
  throw_for(mm_pose_refine(rv, tv, code, camera_params.data(), (int64_t)n, p2.data(), p3.data(), mask.data(), &c, &s, &ret));
  for (int k = 0; k < 3; ++k) { rvec(k) = rv[k]; tvec(k) = tv[k]; }
  if (options.print_progress) std::cout << std::endl;
  if (options.print_summary) print_report("Pose Refinement Report", s, 6);
  return ret;
}

double bundle_adjustment(FeatureManager& fm, const std::vector<size_t>& free_image_ids,
                         const std::vector<size_t>& fixed_image_ids, const std::vector<size_t>& fixed_x_image_ids,
                         const BundleAdjustmentOptions& options, std::unordered_map<size_t, double>& point3D_errors,
                         const std::unordered_map<size_t, Eigen::Vector3d>& rotation_constraints,
                         const std::set<size_t>& gcp_ids) {
  const auto t_enter = std::chrono::steady_clock::now();
  const size_t num_fixed_params = fixed_image_ids.size() * 6 + fixed_x_image_ids.size() + gcp_ids.size() * 3;
  if (num_fixed_params < 7)                                                   // .cc:459-466
    throw std::invalid_argument("At least 7 parameters should be set as fixed to avoid datum defects resulting in a singular Jacobian.");
  if (options.min_track_len < 2)                                              // .cc:468-471
    throw std::invalid_argument("Minimum track length must be >= 2 in order build valid bundle adjustment problem.");
  if (options.constrain_rotation) {
    // _bundle_adjustment_add_pose_constraints (.cc:390-446), first half: rotate EVERY pose and point of the feature manager
    // into the frame of the constraints, M = R_FM' R_C taken at the first fixed image (.cc:402-425).
    // SimilarityTransform3D(M|0): transform_point X' = M X; transform_pose [R|t] -> [R M' | t].
    if (fixed_image_ids.empty()) throw std::invalid_argument("constrain_rotation needs a fixed image");
    double Rf[9], Rc[9], M[9];
    rvec_to_matrix(fm.rvecs[fixed_image_ids[0]], Rf); rvec_to_matrix(rotation_constraints.at(fixed_image_ids[0]), Rc);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[3 * i + j] = Rf[i] * Rc[j] + Rf[3 + i] * Rc[3 + j] + Rf[6 + i] * Rc[6 + j];
    for (auto& kv : fm.rvecs) {
      double R[9], R2[9];
      rvec_to_matrix(kv.second, R);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R2[3 * i + j] = R[3 * i] * M[3 * j] + R[3 * i + 1] * M[3 * j + 1] + R[3 * i + 2] * M[3 * j + 2];
      matrix_to_rvec(R2, kv.second);
    }
    for (auto& kv : fm.points3D) {
      const double x = kv.second(0), y = kv.second(1), z = kv.second(2);
      kv.second(0) = M[0] * x + M[1] * y + M[2] * z; kv.second(1) = M[3] * x + M[4] * y + M[5] * z; kv.second(2) = M[6] * x + M[7] * y + M[8] * z;
    }
  }

  // _bundle_adjustment_extract_data (.cc:228-286): free, fixed_x, fixed
  struct Obs { size_t p2, p3; };
  std::unordered_map<size_t, std::vector<Obs> > per_image;
  std::unordered_map<size_t, size_t> count;
  const std::vector<size_t>* extract_order[3] = { &free_image_ids, &fixed_x_image_ids, &fixed_image_ids };
  for (int g = 0; g < 3; ++g)
    for (size_t image_id : *extract_order[g]) {
      std::vector<Obs>& obs = per_image[image_id];
      obs.clear();
      for (size_t p2 : fm.image_to_points2D[image_id]) {
        auto it = fm.point2D_to_point3D.find(p2);
        if (it == fm.point2D_to_point3D.end()) continue;
        obs.push_back(Obs{p2, it->second});
        count[it->second] += 1;
      }
    }
  // index spaces in residual-block order: free, fixed, fixed_x (.cc:511-533)
  const std::vector<size_t>* fill_order[3] = { &free_image_ids, &fixed_image_ids, &fixed_x_image_ids };
  const int fill_state[3] = { BA_POSE_FREE, BA_POSE_FIXED, BA_POSE_FIXED_X };
  std::vector<size_t> image_ids, camera_ids, point_ids;
  std::unordered_map<size_t, int32_t> img_index, cam_index, pt_index;
  for (int g = 0; g < 3; ++g)
    for (size_t image_id : *fill_order[g])
      if (!img_index.count(image_id)) { img_index[image_id] = (int32_t)image_ids.size(); image_ids.push_back(image_id); }
  for (size_t image_id : image_ids) {
    const size_t cid = fm.image_to_camera[image_id];
    if (!cam_index.count(cid)) { cam_index[cid] = (int32_t)camera_ids.size(); camera_ids.push_back(cid); }
  }
  const size_t n_img = image_ids.size(), n_cam = camera_ids.size();
  std::vector<uint8_t> pose_const(4 * n_img, 0), intr_const(n_cam, 0);
  std::vector<double> obs_xy; std::vector<int32_t> obs_img, obs_pt;
  for (int g = 0; g < 3; ++g)
    for (size_t image_id : *fill_order[g]) {
      const int32_t ii = img_index[image_id];
      size_t num_residuals = 0;
      for (const Obs& o : per_image[image_id]) {
        if (count[o.p3] < options.min_track_len) continue;                    // .cc:326-332
        auto it = pt_index.find(o.p3);
        if (it == pt_index.end()) { it = pt_index.emplace(o.p3, (int32_t)point_ids.size()).first; point_ids.push_back(o.p3); }
        const Eigen::Vector2d& xy = fm.points2D[o.p2];
        obs_xy.push_back(xy(0)); obs_xy.push_back(xy(1));
        obs_img.push_back(ii); obs_pt.push_back(it->second);
        ++num_residuals;
      }
      if (num_residuals > 1) {                                                // .cc:361
        if (fill_state[g] == BA_POSE_FIXED) for (int k = 0; k < 4; ++k) pose_const[4 * ii + k] = 1;
        else if (fill_state[g] == BA_POSE_FIXED_X) pose_const[4 * ii + 1] = 1;
        if (!options.refine_camera_params) intr_const[cam_index[fm.image_to_camera[image_id]]] = 1;
      }
    }
  std::vector<double> poses(6 * n_img), intr(MM_INTR_STRIDE * n_cam, 0.0), pts(3 * point_ids.size());
  std::vector<int32_t> img_cam(n_img), cam_model(n_cam);
  for (size_t k = 0; k < n_img; ++k) {
    const Eigen::Vector3d& r = fm.rvecs[image_ids[k]]; const Eigen::Vector3d& t = fm.tvecs[image_ids[k]];
    for (int c = 0; c < 3; ++c) { poses[6 * k + c] = r(c); poses[6 * k + 3 + c] = t(c); }
    img_cam[k] = cam_index[fm.image_to_camera[image_ids[k]]];
  }
  for (size_t k = 0; k < n_cam; ++k) {
    const std::vector<double>& p = fm.camera_params[camera_ids[k]];
    const int code = (int)p.back();                                           // .cc:339
    if (model_num_params(code) < 0) throw std::invalid_argument("unknown camera model code");
    cam_model[k] = code;
    for (int c = 0; c < model_num_params(code); ++c) intr[MM_INTR_STRIDE * k + c] = p[c];
  }
  std::vector<uint8_t> pt_const(point_ids.size(), 0);
  for (size_t k = 0; k < point_ids.size(); ++k) {
    const Eigen::Vector3d& X = fm.points3D[point_ids[k]];
    pts[3 * k] = X(0); pts[3 * k + 1] = X(1); pts[3 * k + 2] = X(2);
    if (gcp_ids.count(point_ids[k])) pt_const[k] = 1;                         // .cc:545-549
  }
  std::vector<double> pt_err(options.update_point3D_errors ? point_ids.size() : 0, 0.0);

  mm_ba_problem P; memset(&P, 0, sizeof P);
  static double dummy_d[9]; static int32_t dummy_i[1]; static uint8_t dummy_u[4];
  P.n_img = (int32_t)n_img; P.n_cam = (int32_t)n_cam; P.n_pt = (int32_t)point_ids.size(); P.n_obs = (int64_t)obs_img.size();
  P.poses = n_img ? poses.data() : dummy_d; P.pose_const = n_img ? pose_const.data() : dummy_u; P.img_cam = n_img ? img_cam.data() : dummy_i;
  P.intr = n_cam ? intr.data() : dummy_d; P.cam_model = n_cam ? cam_model.data() : dummy_i; P.intr_const = n_cam ? intr_const.data() : dummy_u;
  P.pts = pts.empty() ? dummy_d : pts.data(); P.pt_const = pt_const.empty() ? dummy_u : pt_const.data();
  P.obs_xy = obs_xy.empty() ? dummy_d : obs_xy.data(); P.obs_img = obs_img.empty() ? dummy_i : obs_img.data(); P.obs_pt = obs_pt.empty() ? dummy_i : obs_pt.data();
  P.pt_err = options.update_point3D_errors && !pt_err.empty() ? pt_err.data() : nullptr;
  std::vector<double> rot_prior(3 * n_img, 0.0), rot_prior_w(n_img, 0.0);
  if (options.constrain_rotation) {                                           // .cc:427-443: one residual per FREE image
    for (size_t image_id : free_image_ids) {
      const Eigen::Vector3d& r0 = rotation_constraints.at(image_id);
      const int32_t ii = img_index[image_id];
      for (int cidx = 0; cidx < 3; ++cidx) rot_prior[3 * ii + cidx] = r0(cidx);
      rot_prior_w[ii] = options.constrain_rotation_weight;
    }
    P.rot_prior = rot_prior.data(); P.rot_prior_w = rot_prior_w.data();
  }
  if (P.n_obs == 0) std::cout << "No observations in bundle adjustment. Consider relaxing the constraints." << std::endl;   // .cc:571-573

  mm_ba_options c = to_options(options); mm_ba_summary s;
  const auto t_flat = std::chrono::steady_clock::now();
  throw_for(mm_ba_solve(&P, &c, &s));
  const auto t_solved = std::chrono::steady_clock::now();

  // the reference optimises the FeatureManager storage in place (.cc:243-247, 269-270)
  for (size_t k = 0; k < n_img; ++k) {
    Eigen::Vector3d& r = fm.rvecs[image_ids[k]]; Eigen::Vector3d& t = fm.tvecs[image_ids[k]];
    for (int cidx = 0; cidx < 3; ++cidx) { r(cidx) = poses[6 * k + cidx]; t(cidx) = poses[6 * k + 3 + cidx]; }
  }
  for (size_t k = 0; k < n_cam; ++k) {
    std::vector<double>& p = fm.camera_params[camera_ids[k]];
    for (int cidx = 0; cidx < model_num_params(cam_model[k]); ++cidx) p[cidx] = intr[MM_INTR_STRIDE * k + cidx];
  }
  for (size_t k = 0; k < point_ids.size(); ++k) {
    Eigen::Vector3d& X = fm.points3D[point_ids[k]];
    X(0) = pts[3 * k]; X(1) = pts[3 * k + 1]; X(2) = pts[3 * k + 2];
  }
  if (options.update_point3D_errors)                                         // .cc:575-598
    for (size_t k = 0; k < point_ids.size(); ++k) point3D_errors[point_ids[k]] = pt_err[k];
  if (std::getenv("MM_SHIM_TIMING")) {       // FeatureManager walk (hash maps) vs the device call, SURVEY 8a-a9
    const auto t_done = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::printf("[mavmap_b200 shim] %zu images, %zu points, %zu observations: flatten %.3f ms, mm_ba_solve %.3f ms (device %.3f ms), write-back %.3f ms\n",
                n_img, point_ids.size(), obs_img.size(), ms(t_enter, t_flat), ms(t_flat, t_solved), s.ms_total, ms(t_solved, t_done));
  }
  if (options.print_progress) std::cout << std::endl;
  if (options.print_summary) {
    long long npar = 0;
    for (size_t k = 0; k < n_img; ++k) npar += 3 * !pose_const[4 * k] + !pose_const[4 * k + 1] + !pose_const[4 * k + 2] + !pose_const[4 * k + 3];
    for (size_t k = 0; k < n_cam; ++k) if (!intr_const[k]) npar += model_num_params(cam_model[k]);
    for (size_t k = 0; k < pt_const.size(); ++k) npar += 3 * !pt_const[k];
    print_report("Bundle Adjustment Report", s, npar);
  }
  return s.return_value;                                                      // .cc:610
}
