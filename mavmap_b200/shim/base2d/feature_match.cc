// Drop-in body for match_brute_force() (reference: src/base2d/feature.cc:52-133, declared in
// src/base2d/feature.h:102-110).  Compile it instead of the reference's OpenCV BFMatcher code;
// AdaptiveSURF and median_feature_disparity of feature.cc are untouched (out of scope).
#include <stdexcept>
#include <string>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include "mavmap_b200.h"

void match_brute_force(const std::vector<cv::KeyPoint>& keypoints1, const cv::Mat& descriptors1,
                       const std::vector<cv::KeyPoint>& keypoints2, const cv::Mat& descriptors2,
                       std::vector<cv::DMatch>& matches, const bool ratio_test, const double max_ratio,
                       const double max_distance, const int norm_type) {
  matches.clear();                                                            // feature.cc:62
  if (norm_type != cv::NORM_L2) throw std::invalid_argument("mavmap_b200: only cv::NORM_L2 is built (the mapper's choice)");
  const int n1 = descriptors1.rows, n2 = descriptors2.rows, k = descriptors1.cols;
  if (n1 == 0 || n2 == 0) return;
  cv::Mat d1 = descriptors1.isContinuous() ? descriptors1 : descriptors1.clone();
  cv::Mat d2 = descriptors2.isContinuous() ? descriptors2 : descriptors2.clone();
  std::vector<float> xy1, xy2;
  if (max_distance != -1) {                                                   // feature.cc:23-49 reads only .pt
    xy1.resize(2 * keypoints1.size()); xy2.resize(2 * keypoints2.size());
    for (size_t i = 0; i < keypoints1.size(); ++i) { xy1[2 * i] = keypoints1[i].pt.x; xy1[2 * i + 1] = keypoints1[i].pt.y; }
    for (size_t i = 0; i < keypoints2.size(); ++i) { xy2[2 * i] = keypoints2[i].pt.x; xy2[2 * i + 1] = keypoints2[i].pt.y; }
  }
  mm_match_options o; mm_match_options_default(&o);
  o.ratio_test = ratio_test ? 1 : 0; o.max_ratio = max_ratio; o.max_distance = max_distance;
  const int cap = n1 < n2 ? n1 : n2;
  std::vector<int32_t> q(cap), t(cap); std::vector<float> dist(cap); int32_t n_out = 0;
  const int rc = mm_match_pair(d1.ptr<float>(0), n1, d2.ptr<float>(0), n2, k, xy1.empty() ? nullptr : xy1.data(), xy2.empty() ? nullptr : xy2.data(),
                               &o, q.data(), t.data(), dist.data(), &n_out);
  if (rc != MM_OK) throw std::runtime_error(std::string("mavmap_b200: ") + mm_last_error());
  matches.reserve(n_out);
  for (int i = 0; i < n_out; ++i) matches.push_back(cv::DMatch(q[i], t[i], dist[i]));
}
