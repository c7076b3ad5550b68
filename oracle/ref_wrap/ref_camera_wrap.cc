// C entry points around the REFERENCE's own camera models (compiled from
// /root/reference/src/base3d/camera_models.{h,cc} where they lie; nothing is copied).
// Output: oracle/_ref/libref_camera.so — used by tests to validate the oracle restatement.
#include <vector>
#include "base3d/camera_models.h"

extern "C" {
void ref_world2image(int code, const double* params, long n, const double* xyz, double* uv) {
  for (long i = 0; i < n; ++i)
    camera_model_world2image(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], uv[2 * i], uv[2 * i + 1], code, params);
}
void ref_image2world(int code, const double* params, long n, const double* uv, double* xyz) {
  for (long i = 0; i < n; ++i)
    camera_model_image2world(uv[2 * i], uv[2 * i + 1], xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], code, params);
}
void ref_image2world_normalized(int code, const double* params, long n, const double* uv, double* xy) {
  std::vector<Eigen::Vector2d> in(n), out;
  for (long i = 0; i < n; ++i) in[i] = Eigen::Vector2d(uv[2 * i], uv[2 * i + 1]);
  camera_model_image2world(in, out, code, params);
  for (long i = 0; i < n; ++i) { xy[2 * i] = out[i](0); xy[2 * i + 1] = out[i](1); }
}
int ref_name_to_code(const char* name) { return camera_model_name_to_code(name); }
double ref_threshold(double t, int code, const double* params) { return camera_model_image2world_threshold(t, code, params); }
}
