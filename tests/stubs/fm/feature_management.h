// Test stand-in for the reference's src/fm/feature_management.h: the public data members the BA shim
// reads and writes (feature_management.h:189-230) plus the three builders the test needs.
#ifndef TEST_FM_STUB_
#define TEST_FM_STUB_
#include <unordered_map>
#include <vector>
#include <Eigen/Core>
class FeatureManager {
 public:
  FeatureManager() : num_cameras_(0), num_images_(0), num_points2D_(0), num_points3D_(0) {}
  size_t add_camera(const std::vector<double>& params) { camera_params[++num_cameras_] = params; return num_cameras_; }
  size_t add_image(const size_t camera_id, const std::vector<Eigen::Vector2d>& pts) {
    const size_t id = ++num_images_; image_to_camera[id] = camera_id; rvecs[id] = Eigen::Vector3d(); tvecs[id] = Eigen::Vector3d();
    image_to_points2D[id] = std::vector<size_t>();
    for (size_t i = 0; i < pts.size(); ++i) { const size_t p = ++num_points2D_; points2D[p] = pts[i]; image_to_points2D[id].push_back(p); point2D_to_image[p] = id; }
    return id;
  }
  size_t add_point3D() { const size_t id = ++num_points3D_; points3D[id] = Eigen::Vector3d(); return id; }
  std::unordered_map<size_t, Eigen::Vector3d> points3D;
  std::unordered_map<size_t, Eigen::Vector2d> points2D;
  std::unordered_map<size_t, size_t> point2D_to_point3D;
  std::unordered_map<size_t, size_t> point2D_to_image;
  std::unordered_map<size_t, std::vector<size_t> > image_to_points2D;
  std::unordered_map<size_t, std::vector<size_t> > point3D_to_points2D;
  std::unordered_map<size_t, Eigen::Vector3d> rvecs;
  std::unordered_map<size_t, Eigen::Vector3d> tvecs;
  std::unordered_map<size_t, size_t> image_to_camera;
  std::unordered_map<size_t, std::vector<double> > camera_params;
 private:
  size_t num_cameras_, num_images_, num_points2D_, num_points3D_;
};
#endif
