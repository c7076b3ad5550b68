// Drives the reference's OWN SequentialMapper::adjust_global_bundle / adjust_bundle (src/sfm/sequential_mapper.cc:1040-1160,
// unmodified, compiled from /root/reference) on a synthetic 20-image sequence (BASELINE cfg1 size).  The mapper state a real
// run would have built by processing images is filled in directly; the bundle adjustment itself goes
// SequentialMapper -> bundle_adjustment() of the shim -> mm_ba_solve -> the CUDA engine.
// Prints one line per check; exit code 0 iff all pass.  Built by build_mapper.sh, run by tests/test_mapper_harness.py (GPU).
#include <cmath>
#include <cstdio>
#include <random>
// everything sequential_mapper.h includes comes first, so that only SequentialMapper itself is opened up below
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <unordered_map>
#include <vector>
#include <boost/lexical_cast.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/features2d/features2d.hpp>
#include "base2d/feature.h"
#include "base2d/feature_cache.h"
#include "base2d/image.h"
#include "base3d/bundle_adjustment.h"
#include "base3d/camera_models.h"
#include "base3d/essential_matrix.h"
#include "base3d/p3p.h"
#include "base3d/projection.h"
#include "base3d/projective_transform.h"
#include "base3d/similarity_transform.h"
#include "base3d/triangulation.h"
#include "fm/feature_management.h"
#include "loop/detection.h"
#include "util/estimation.h"
#include "util/opencv.h"
#include "util/path.h"
#include "util/timer.h"
#define private public            // test-only: the driver sets the bookkeeping members that process() maintains
#include "sfm/sequential_mapper.h"
#undef private

static int fails = 0;
#define CHECK(cond, what) do { const bool ok_ = (cond); std::printf("%s %s\n", ok_ ? "ok  " : "FAIL", what); if (!ok_) ++fails; } while (0)

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  std::mt19937 rng(20); std::normal_distribution<double> N(0.0, 1.0); std::uniform_real_distribution<double> U(-1.0, 1.0);
  const int n_img = 20, n_pt = 1000;
  std::vector<Image> images(n_img);
  for (int i = 0; i < n_img; ++i) { images[i].name = "img" + std::to_string(i); images[i].camera_idx = 0; images[i].rows = 960; images[i].cols = 1280; images[i].roll = images[i].pitch = images[i].yaw = 0; }
  SequentialMapper mapper(images, "/tmp/mavmap_b200_harness_cache", "", SURFOptions(), /*loop_detection=*/false, /*debug=*/false, "");
  FeatureManager& fm = mapper.feature_manager;
  const size_t cam = fm.add_camera({1000.0, 1000.0, 640.0, 480.0, 1});      // PINHOLE, model code as the last entry (sequential_mapper.cc:960-965)
  // scene: cameras along x looking down +z at points 8..12 away; truth poses t_i = (-0.4 i, 0, 0)
  std::vector<Eigen::Vector3d> X(n_pt);
  for (auto& x : X) x = Eigen::Vector3d(4.0 * U(rng) + 3.8, 3.0 * U(rng), 10.0 + 2.0 * U(rng));
  std::vector<Eigen::Vector3d> t_true(n_img);
  std::vector<size_t> ids(n_img);
  std::vector<std::vector<int>> seen(n_img);
  for (int i = 0; i < n_img; ++i) {
    t_true[i] = Eigen::Vector3d(-0.4 * i, 0.0, 0.0);
    std::vector<Eigen::Vector2d> uv;
    for (int p = 0; p < n_pt; ++p) {
      const double xc = X[p](0) + t_true[i](0), yc = X[p](1), zc = X[p](2);
      const double u = 1000.0 * xc / zc + 640.0, v = 1000.0 * yc / zc + 480.0;
      if (u < 0 || u >= 1280 || v < 0 || v >= 960) continue;
      uv.push_back(Eigen::Vector2d(u + 0.3 * N(rng), v + 0.3 * N(rng))); seen[i].push_back(p);
    }
    ids[i] = fm.add_image(cam, uv);
    fm.rvecs[ids[i]] = Eigen::Vector3d(0.002 * N(rng), 0.002 * N(rng), 0.002 * N(rng));
    fm.tvecs[ids[i]] = Eigen::Vector3d(t_true[i](0) + 0.02 * N(rng), 0.02 * N(rng), 0.02 * N(rng));
    mapper.image_idx_to_id_[i] = ids[i]; mapper.image_id_to_idx_[ids[i]] = i;
  }
  fm.rvecs[ids[0]] = Eigen::Vector3d(0, 0, 0); fm.tvecs[ids[0]] = t_true[0];       // the datum: image 0 FIXED, image 1 FIXED_X
  fm.tvecs[ids[1]](0) = t_true[1](0);
  std::vector<size_t> pid(n_pt, 0);
  for (int p = 0; p < n_pt; ++p) { pid[p] = fm.add_point3D(); fm.points3D[pid[p]] = Eigen::Vector3d(X[p](0) + 0.05 * N(rng), X[p](1) + 0.05 * N(rng), X[p](2) + 0.05 * N(rng)); }
  size_t n_obs = 0;
  for (int i = 0; i < n_img; ++i)
    for (size_t k = 0; k < seen[i].size(); ++k) { const size_t p2 = fm.image_to_points2D[ids[i]][k]; fm.point2D_to_point3D[p2] = pid[seen[i][k]]; fm.point3D_to_points2D[pid[seen[i][k]]].push_back(p2); ++n_obs; }
  mapper.num_proc_images_ = n_img; mapper.first_image_idx_ = 0; mapper.second_image_idx_ = 1; mapper.min_image_idx_ = 0; mapper.max_image_idx_ = n_img - 1;
  mapper.prev_image_idx_ = n_img - 1; mapper.prev_prev_image_idx_ = n_img - 2;
  std::printf("scene: %d images, %d points, %zu observations\n", n_img, n_pt, n_obs);

  auto pose_err = [&]() { double e = 0; for (int i = 0; i < n_img; ++i) for (int k = 0; k < 3; ++k) e = std::max(e, std::fabs(fm.tvecs[ids[i]](k) - t_true[i](k))); return e; };
  const double err_before = pose_err();
  // local BA as mapper.cc:1120-1135 runs it after every image: the last 5 images free, the two before fixed
  BundleAdjustmentOptions local; local.max_num_iterations = 10; local.print_summary = false; local.refine_camera_params = false; local.min_track_len = 2;
  const double c_local = mapper.adjust_bundle({15, 16, 17, 18, 19}, {13}, {14}, local);
  CHECK(c_local > 0.05 && c_local < 1.0, "SequentialMapper::adjust_bundle (local window) returns a cost at the noise level");
  // global BA as mapper.cc:170-174: image 0 FIXED, image 1 FIXED_X (sequential_mapper.cc:1095-1097), everything else free
  BundleAdjustmentOptions global; global.max_num_iterations = 30; global.print_summary = false; global.refine_camera_params = false;
  const double c_global = mapper.adjust_global_bundle(global);
  const double err_after = pose_err();
  std::printf("global BA: cost %.4f px, max |t - t_true| %.4f -> %.4f\n", c_global, err_before, err_after);
  CHECK(c_global > 0.05 && c_global < 0.5, "SequentialMapper::adjust_global_bundle returns sqrt(final_cost/num_residuals) ~ 0.3 px noise");
  CHECK(err_after < 0.04 && err_after < 0.5 * err_before, "camera positions recovered (to the noise / gauge level)");
  CHECK(fm.tvecs[ids[0]](0) == t_true[0](0) && fm.rvecs[ids[0]](1) == 0.0, "FIXED image untouched");
  CHECK(fm.tvecs[ids[1]](0) == t_true[1](0), "FIXED_X image keeps its x translation");
  // refined intrinsics (the mapper's default, mapper.cc:878-886): start 1 % off in the focal length
  fm.camera_params[cam][0] *= 1.01; fm.camera_params[cam][1] *= 1.01;
  global.refine_camera_params = true; global.update_point3D_errors = true;
  const double c_ref = mapper.adjust_global_bundle(global);
  std::printf("refine: cost %.4f px, fx %.3f fy %.3f\n", c_ref, fm.camera_params[cam][0], fm.camera_params[cam][1]);
  CHECK(c_ref < 0.5 && c_ref <= c_global * 1.001 && std::fabs(fm.camera_params[cam][0] - 1000.0) < 9.0 && std::fabs(fm.camera_params[cam][1] - 1000.0) < 9.0, "refine_camera_params moves the focal length back towards the truth at no higher cost");
  size_t seen_pid = 0; for (int p = 0; p < n_pt; ++p) if (fm.point3D_to_points2D[pid[p]].size() >= 2) { seen_pid = pid[p]; break; }
  bool have_err = false; double e0 = -1.0;
  try { e0 = mapper.get_point3D_error(seen_pid); have_err = true; } catch (const std::range_error&) {}
  CHECK(have_err && e0 >= 0.0 && e0 < 5.0, "point3D errors available from the mapper after BA (update_point3D_errors)");
  std::printf(fails ? "FAILED (%d)\n" : "all ok\n", fails);
  return fails ? 1 : 0;
}
