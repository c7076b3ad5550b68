// Drop-in replacement for mavmap/mavmap src/base3d/bundle_adjustment.h (Ceres-free).
//
// Same macros, option struct and free-function signatures as the reference header
// (bundle_adjustment.h:33-35, :38-114, :212-230), so src/sfm/sequential_mapper.{h,cc} and
// src/mapper.cc compile against it unchanged.  The bodies (bundle_adjustment.cc next to this
// file) flatten the FeatureManager subset into SoA arrays and call the C ABI of
// libmavmap_b200.so (include/mavmap_b200.h); the Ceres cost-functor classes of the reference
// header (:117-209) have no equivalent here because no caller outside bundle_adjustment.cc uses them.
#ifndef MAVMAP_SRC_BASE3D_BUNDLE_ADJUSTMENT_H_
#define MAVMAP_SRC_BASE3D_BUNDLE_ADJUSTMENT_H_

#include <cstddef>
#include <set>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include <Eigen/Core>

#include "fm/feature_management.h"

#define BA_POSE_FREE       0
#define BA_POSE_FIXED      1
#define BA_POSE_FIXED_X    2

struct BundleAdjustmentOptions {
  BundleAdjustmentOptions() : max_num_iterations(100), function_tolerance(1e-4), gradient_tolerance(1e-8),
                              update_point3D_errors(false), min_track_len(2), loss_scale_factor(1),
                              constrain_rotation(false), constrain_rotation_weight(0), refine_camera_params(false),
                              print_progress(false), print_summary(true) {}
  size_t max_num_iterations;
  double function_tolerance;
  double gradient_tolerance;
  bool update_point3D_errors;
  size_t min_track_len;
  double loss_scale_factor;
  bool constrain_rotation;
  double constrain_rotation_weight;
  bool refine_camera_params;
  bool print_progress;
  bool print_summary;
};

double pose_refinement(Eigen::Vector3d& rvec, Eigen::Vector3d& tvec, std::vector<double>& camera_params,
                       const std::vector<Eigen::Vector2d>& points2D, std::vector<Eigen::Vector3d>& points3D,
                       const std::vector<bool>& inlier_mask, const BundleAdjustmentOptions& options);

double bundle_adjustment(FeatureManager& feature_manager, const std::vector<size_t>& free_image_ids,
                         const std::vector<size_t>& fixed_image_ids, const std::vector<size_t>& fixed_x_image_ids,
                         const BundleAdjustmentOptions& options, std::unordered_map<size_t, double>& point3D_errors,
                         const std::unordered_map<size_t, Eigen::Vector3d>& rotation_constraints
                           = std::unordered_map<size_t, Eigen::Vector3d>(),
                         const std::set<size_t>& gcp_ids = std::set<size_t>());

#endif  // MAVMAP_SRC_BASE3D_BUNDLE_ADJUSTMENT_H_
