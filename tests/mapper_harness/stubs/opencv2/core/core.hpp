// stand-in for the OpenCV 2.4 headers (tests/mapper_harness): the types and free functions the reference's mapper layer names.
// cv::Mat is a row-major float/byte buffer with the members match_brute_force's shim reads (rows, cols, ptr<T>(), isContinuous).
// Image I/O and drawing only have to link; they do nothing.
#ifndef MAPPER_HARNESS_CV_CORE_
#define MAPPER_HARNESS_CV_CORE_
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_LOAD_IMAGE_COLOR 1
#define CV_GRAY2RGB 8
namespace cv {
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { COLOR_GRAY2RGB = 8, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8 };
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} template <typename U> Point_(const Point_<U>& o) : x(T(o.x)), y(T(o.y)) {} };
typedef Point_<float> Point2f; typedef Point_<int> Point; typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } static Scalar all(double a) { return Scalar(a, a, a, a); } double& operator[](int i) { return v[i]; } const double& operator[](int i) const { return v[i]; } };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} KeyPoint(float x, float y, float s) : pt(x, y), size(s), angle(-1), response(0), octave(0), class_id(-1) {} };
struct DMatch {
  int queryIdx, trainIdx, imgIdx; float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(0), distance(d) {}
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  bool operator<(const DMatch& o) const { return distance < o.distance; }
};
class Mat;
// reference-counted storage like cv::Mat: copies share the pixels, clone() copies them
class Mat {
 public:
  int rows, cols, type_; std::shared_ptr<std::vector<unsigned char>> store; unsigned char* data;
  Mat() : rows(0), cols(0), type_(CV_32F), data(nullptr) {}
  Mat(int r, int c, int t = CV_32F) : rows(0), cols(0), type_(t), data(nullptr) { create(r, c, t); }
  Mat(int r, int c, int t, const Scalar& s) : rows(0), cols(0), type_(t), data(nullptr) { create(r, c, t); setTo(s); }
  // wraps caller memory in OpenCV; here the bytes are copied (enough for the call sites that only read the mask afterwards)
  Mat(int r, int c, int t, void* user) : rows(0), cols(0), type_(t), data(nullptr) { create(r, c, t); std::copy(static_cast<unsigned char*>(user), static_cast<unsigned char*>(user) + store->size(), data); }
  Mat(Size sz, int t) : rows(0), cols(0), type_(t), data(nullptr) { create(sz.height, sz.width, t); }
  static size_t elemSize_of(int t) { return t == CV_32F ? 4 : (t == CV_64F ? 8 : (t == CV_8UC3 ? 3 : 1)); }
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  static Mat zeros(Size sz, int t) { return Mat(sz.height, sz.width, t); }
  static Mat ones(int r, int c, int t) { Mat m(r, c, t); m.setTo(Scalar::all(1)); return m; }
  void create(int r, int c, int t) { rows = r; cols = c; type_ = t; store.reset(new std::vector<unsigned char>((size_t)r * c * elemSize_of(t), 0)); data = store->data(); }
  int type() const { return type_; }
  int depth() const { return type_ == CV_8UC3 ? CV_8U : type_; }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  size_t elemSize() const { return elemSize_of(type_); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return rows == 0 || cols == 0 || !data; }
  bool isContinuous() const { return true; }
  Size size() const { return Size(cols, rows); }
  Mat clone() const { Mat m; if (store) { m.create(rows, cols, type_); *m.store = *store; m.data = m.store->data(); } return m; }
  void release() { rows = cols = 0; store.reset(); data = nullptr; }
  Mat& setTo(const Scalar& s, const Mat& /*mask*/ = Mat()) {
    for (size_t i = 0; i < total() * channels(); ++i) { const double v = s[(int)(i % channels())]; if (depth() == CV_32F) reinterpret_cast<float*>(data)[i] = (float)v; else if (depth() == CV_64F) reinterpret_cast<double*>(data)[i] = v; else data[i] = (unsigned char)v; }
    return *this;
  }
  Mat& operator=(const Scalar& s) { return setTo(s); }
  unsigned char* ptr(size_t r = 0) { return data + r * cols * elemSize(); }
  const unsigned char* ptr(size_t r = 0) const { return data + r * cols * elemSize(); }
  template <typename T> T* ptr(size_t r = 0) { return reinterpret_cast<T*>(data + r * cols * elemSize()); }
  template <typename T> const T* ptr(size_t r = 0) const { return reinterpret_cast<const T*>(data + r * cols * elemSize()); }
  template <typename T> T& at(size_t r, size_t c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(size_t r, size_t c) const { return ptr<T>(r)[c]; }
  Mat row(int r) const { Mat m(1, cols, type_); std::copy(ptr((size_t)r), ptr((size_t)r + 1), m.data); return m; }
  Mat operator()(const Rect& roi) const { Mat m(roi.height, roi.width, type_); const size_t e = elemSize(); for (int r = 0; r < roi.height; ++r) for (int c = 0; c < roi.width; ++c) { const int rr = roi.y + r, cc = roi.x + c; if (rr >= 0 && rr < rows && cc >= 0 && cc < cols) std::copy(data + ((size_t)rr * cols + cc) * e, data + ((size_t)rr * cols + cc + 1) * e, m.data + ((size_t)r * roi.width + c) * e); } return m; }
  void push_back(const Mat& m) { if (empty()) { *this = m.clone(); return; } Mat n(rows + m.rows, cols, type_); std::copy(data, data + total() * elemSize(), n.data); std::copy(m.data, m.data + m.total() * m.elemSize(), n.data + total() * elemSize()); *this = n; }
  void convertTo(Mat& dst, int) const { dst = clone(); }
  void copyTo(Mat& dst) const { dst = clone(); }
};
inline const Mat& noArray() { static Mat none; return none; }
typedef const Mat& InputArray; typedef Mat& OutputArray;
Scalar mean(const Mat& m);
void split(const Mat& m, Mat* planes);
void split(const Mat& m, std::vector<Mat>& planes);
void merge(const std::vector<Mat>& planes, Mat& m);
void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1, int line_type = 8, int shift = 0);
void line(Mat& img, Point a, Point b, const Scalar& color, int thickness = 1, int line_type = 8, int shift = 0);
}
#endif
