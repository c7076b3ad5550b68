// Test stand-in for OpenCV 2.4 core types used by match_brute_force's signature (cv::Mat CV_32F, cv::KeyPoint, cv::DMatch).
#ifndef TEST_CV_CORE_STUB_
#define TEST_CV_CORE_STUB_
#include <vector>
namespace cv {
enum { NORM_L2 = 4 };
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; };
struct DMatch {
  int queryIdx, trainIdx, imgIdx; float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
};
struct Mat {
  int rows, cols; std::vector<float> data;
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c) : rows(r), cols(c), data((size_t)r * c) {}
  bool isContinuous() const { return true; }
  Mat clone() const { return *this; }
  template <typename T> const T* ptr(int r) const { return reinterpret_cast<const T*>(data.data()) + (size_t)r * cols; }
  template <typename T> T* ptr(int r) { return reinterpret_cast<T*>(data.data()) + (size_t)r * cols; }
};
}  // namespace cv
#endif
