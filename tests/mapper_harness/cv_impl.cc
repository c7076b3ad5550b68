// Definitions behind the OpenCV stand-in headers (tests/mapper_harness/stubs/opencv2): enough for the reference's mapper to
// link.  Image I/O, drawing and SURF do nothing useful here (the harness feeds the mapper from the feature cache or not at
// all); cv::mean / cv::split work on the stand-in Mat; BFMatcher is an exact brute-force matcher.
#include <cmath>
#include <limits>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/nonfree/features2d.hpp>
namespace cv {
Scalar mean(const Mat& m) {
  Scalar s; const int ch = m.channels(); if (m.empty()) return s;
  for (size_t i = 0; i < m.total(); ++i) for (int c = 0; c < ch; ++c) s[c] += m.depth() == CV_32F ? reinterpret_cast<const float*>(m.data)[i * ch + c] : (m.depth() == CV_64F ? reinterpret_cast<const double*>(m.data)[i * ch + c] : m.data[i * ch + c]);
  for (int c = 0; c < ch; ++c) s[c] /= (double)m.total();
  return s;
}
void split(const Mat& m, Mat* planes) {
  const int ch = m.channels();
  for (int c = 0; c < ch; ++c) { planes[c].create(m.rows, m.cols, m.depth()); for (size_t i = 0; i < m.total(); ++i) planes[c].data[i] = m.data[i * ch + c]; }
}
void split(const Mat& m, std::vector<Mat>& planes) { planes.resize(m.channels()); split(m, planes.data()); }
void merge(const std::vector<Mat>& planes, Mat& m) { if (planes.empty()) return; m.create(planes[0].rows, planes[0].cols, planes.size() == 3 ? CV_8UC3 : planes[0].type()); for (size_t c = 0; c < planes.size(); ++c) for (size_t i = 0; i < m.total(); ++i) m.data[i * planes.size() + c] = planes[c].data[i]; }
void circle(Mat&, Point, int, const Scalar&, int, int, int) {}
void line(Mat&, Point, Point, const Scalar&, int, int, int) {}
void cvtColor(const Mat& src, Mat& dst, int, int) { dst = src.clone(); }
void resize(const Mat& src, Mat& dst, Size, double, double, int) { dst = src.clone(); }
Mat imread(const std::string&, int) { return Mat(); }
bool imwrite(const std::string&, const Mat&, const std::vector<int>&) { return false; }
void drawMatches(const Mat&, const std::vector<KeyPoint>&, const Mat&, const std::vector<KeyPoint>&, const std::vector<DMatch>&, Mat&, const Scalar&, const Scalar&, const std::vector<char>&, int) {}
void drawKeypoints(const Mat&, const std::vector<KeyPoint>&, Mat&, const Scalar&, int) {}
BFMatcher::BFMatcher(int norm_type, bool cross_check) : norm_(norm_type), cross_(cross_check) {}
void BFMatcher::knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& m, int k, const Mat& mask, bool) const {
  m.assign(q.rows, std::vector<DMatch>());
  for (int i = 0; i < q.rows; ++i) {
    std::vector<DMatch> best;
    for (int j = 0; j < t.rows; ++j) {
      if (!mask.empty() && !mask.at<unsigned char>(i, j)) continue;
      double s = 0; for (int c = 0; c < q.cols; ++c) { const double d = (double)q.at<float>(i, c) - t.at<float>(j, c); s += d * d; }
      DMatch dm(i, j, (float)std::sqrt((float)s));
      size_t pos = best.size(); while (pos > 0 && dm.distance < best[pos - 1].distance) --pos;
      best.insert(best.begin() + pos, dm); if ((int)best.size() > k) best.pop_back();
    }
    m[i] = best;
  }
}
void BFMatcher::match(const Mat& q, const Mat& t, std::vector<DMatch>& m, const Mat& mask) const {
  std::vector<std::vector<DMatch>> k; knnMatch(q, t, k, 1, mask); m.clear(); for (auto& r : k) if (!r.empty()) m.push_back(r[0]);
}
void SURF::detect(const Mat&, std::vector<KeyPoint>& kp, const Mat&) const { kp.clear(); }
void SURF::compute(const Mat&, std::vector<KeyPoint>& kp, Mat& desc) const { desc.create((int)kp.size(), descriptorSize(), CV_32F); }
void SURF::operator()(const Mat& img, const Mat& mask, std::vector<KeyPoint>& kp, Mat& desc, bool use_provided) const { if (!use_provided) detect(img, kp, mask); compute(img, kp, desc); }
}  // namespace cv
