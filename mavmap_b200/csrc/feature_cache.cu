// feature_cache.cu — reader for the on-disk feature cache of mavmap/mavmap (src/base2d/feature_cache.cc:126-163), so that
// the all-pairs matcher can be fed from a directory the reference mapper has already populated, without OpenCV.
//
//   <image>.keypoints   : size_t num_bytes | num_bytes of cv::KeyPoint (OpenCV 2.4: Point2f pt, float size, angle, response,
//                         int octave, class_id = 28 bytes each; only pt is used by match_brute_force, feature.cc:23-49)
//   <image>.descriptors : size_t num_bytes | 8 bytes "rows" | 8 bytes "cols" | int type | num_bytes of data
// The writer stores `descriptors.rows` and `descriptors.cols` (both int) with sizeof(size_t) (feature_cache.cc:140-141), so
// each 8-byte field holds the value in its LOW 32 bits and whatever follows in cv::Mat in the high 32 (cols, then the low
// half of the data pointer); the reference's own reader survives because cv::Mat(int, int, int) truncates.  Same here.
#include <stdio.h>
#include <vector>
#include <string>
#include "common.cuh"

namespace {
constexpr int KEYPOINT_BYTES = 28;      // sizeof(cv::KeyPoint), OpenCV 2.4.x
constexpr int CV_32F_TYPE = 5;          // CV_32FC1

struct File {
  FILE* f = nullptr;
  explicit File(const char* path) { f = path ? fopen(path, "rb") : nullptr; }
  ~File() { if (f) fclose(f); }
  bool read(void* dst, size_t n) { return f && fread(dst, 1, n, f) == n; }
  // bytes left between the current position and the end of the file (headers announce payload sizes: check before allocating)
  uint64_t remaining() { if (!f) return 0; const long pos = ftell(f); fseek(f, 0, SEEK_END); const long end = ftell(f); fseek(f, pos, SEEK_SET); return end > pos ? (uint64_t)(end - pos) : 0; }
};

int read_desc_header(File& d, const char* path, uint64_t* nbytes, int32_t* rows, int32_t* cols, int32_t* type) {
  uint64_t r64 = 0, c64 = 0; int32_t t = 0;
  if (!d.read(nbytes, 8) || !d.read(&r64, 8) || !d.read(&c64, 8) || !d.read(&t, 4)) { mm::set_error("cannot read descriptor header of %s", path ? path : "(null)"); return MM_ERR_INVALID_ARG; }
  *rows = (int32_t)(r64 & 0xffffffffu); *cols = (int32_t)(c64 & 0xffffffffu); *type = t;
  if (*rows < 0 || *cols < 0) { mm::set_error("negative descriptor matrix size in %s", path); return MM_ERR_INVALID_ARG; }
  return MM_OK;
}
}  // namespace

extern "C" {

int mm_feature_cache_info(const char* keypoints_path, const char* descriptors_path, int32_t* n_keypoints, int32_t* rows, int32_t* cols, int32_t* cv_type) {
  if (!n_keypoints || !rows || !cols || !cv_type) { mm::set_error("null argument"); return MM_ERR_INVALID_ARG; }
  File k(keypoints_path), d(descriptors_path);
  uint64_t kb = 0, nbytes = 0;
  if (!k.read(&kb, 8)) { mm::set_error("cannot read %s", keypoints_path ? keypoints_path : "(null)"); return MM_ERR_INVALID_ARG; }
  if (kb % KEYPOINT_BYTES) { mm::set_error("%s: %llu bytes is not a whole number of cv::KeyPoint", keypoints_path, (unsigned long long)kb); return MM_ERR_INVALID_ARG; }
  *n_keypoints = (int32_t)(kb / KEYPOINT_BYTES);
  return read_desc_header(d, descriptors_path, &nbytes, rows, cols, cv_type);
}

int mm_feature_cache_read(const char* keypoints_path, const char* descriptors_path, float* xy, float* desc, int32_t cap_rows, int32_t cols_expected) {
  File k(keypoints_path), d(descriptors_path);
  uint64_t kb = 0, nbytes = 0; int32_t rows = 0, cols = 0, type = 0;
  if (!k.read(&kb, 8) || kb % KEYPOINT_BYTES) { mm::set_error("cannot read %s", keypoints_path ? keypoints_path : "(null)"); return MM_ERR_INVALID_ARG; }
  int rc = read_desc_header(d, descriptors_path, &nbytes, &rows, &cols, &type); if (rc) return rc;
  const int64_t n_kp = (int64_t)(kb / KEYPOINT_BYTES);
  if (rows > 0 && type != CV_32F_TYPE) { mm::set_error("%s: descriptor type %d is not CV_32F", descriptors_path, type); return MM_ERR_UNSUPPORTED; }
  if (nbytes != (uint64_t)rows * (uint64_t)cols * 4u) { mm::set_error("%s: %llu data bytes for a %d x %d float matrix", descriptors_path, (unsigned long long)nbytes, rows, cols); return MM_ERR_INVALID_ARG; }
  if (n_kp != rows) { mm::set_error("%lld keypoints but %d descriptor rows", (long long)n_kp, rows); return MM_ERR_INVALID_ARG; }
  if (rows > cap_rows || (rows > 0 && cols != cols_expected)) { mm::set_error("buffer too small or descriptor length mismatch (%d x %d)", rows, cols); return MM_ERR_INVALID_ARG; }
  if (kb > k.remaining() || nbytes > d.remaining()) { mm::set_error("truncated feature cache file (%s / %s)", keypoints_path, descriptors_path); return MM_ERR_INVALID_ARG; }
  std::vector<unsigned char> raw((size_t)kb);
  if (kb && !k.read(raw.data(), (size_t)kb)) { mm::set_error("truncated %s", keypoints_path); return MM_ERR_INVALID_ARG; }
  if (xy) for (int64_t i = 0; i < n_kp; ++i) memcpy(xy + 2 * i, raw.data() + (size_t)i * KEYPOINT_BYTES, 8);      // KeyPoint::pt
  if (nbytes && desc && !d.read(desc, (size_t)nbytes)) { mm::set_error("truncated %s", descriptors_path); return MM_ERR_INVALID_ARG; }
  return MM_OK;
}

// descriptors (and keypoints) of a whole sequence straight from the cache files into one resident match set
int mm_match_set_create_from_cache(const char* const* keypoints_paths, const char* const* descriptors_paths, int32_t n_images, mm_match_set** out) {
  if (!out || n_images < 0 || (n_images > 0 && (!keypoints_paths || !descriptors_paths))) { mm::set_error("invalid argument"); return MM_ERR_INVALID_ARG; }
  *out = nullptr;
  std::vector<int32_t> counts((size_t)n_images, 0);
  int32_t k = 0; int64_t total = 0;
  for (int i = 0; i < n_images; ++i) {
    int32_t nk = 0, r = 0, c = 0, t = 0;
    int rc = mm_feature_cache_info(keypoints_paths[i], descriptors_paths[i], &nk, &r, &c, &t); if (rc) return rc;
    if (r > 0) { if (k == 0) k = c; else if (c != k) { mm::set_error("descriptor length changes from %d to %d at image %d", k, c, i); return MM_ERR_INVALID_ARG; } }
    counts[i] = r; total += r;
  }
  if (k == 0) k = 64;
  std::vector<float> desc((size_t)total * k), xy((size_t)total * 2);
  int64_t off = 0;
  for (int i = 0; i < n_images; ++i) {
    int rc = mm_feature_cache_read(keypoints_paths[i], descriptors_paths[i], xy.data() + 2 * off, desc.data() + (size_t)off * k, counts[i], k); if (rc) return rc;
    off += counts[i];
  }
  return mm_match_set_create(desc.data(), xy.data(), counts.data(), n_images, k, out);
}

}  // extern "C"
