#ifndef MAPPER_HARNESS_CV_FEATURES2D_
#define MAPPER_HARNESS_CV_FEATURES2D_
#include <opencv2/core/core.hpp>
namespace cv {
struct DrawMatchesFlags { enum { DEFAULT = 0, DRAW_OVER_OUTIMG = 1, NOT_DRAW_SINGLE_POINTS = 2, DRAW_RICH_KEYPOINTS = 4 }; };
void drawMatches(const Mat& img1, const std::vector<KeyPoint>& kp1, const Mat& img2, const std::vector<KeyPoint>& kp2, const std::vector<DMatch>& matches, Mat& out,
                 const Scalar& match_color = Scalar::all(-1), const Scalar& point_color = Scalar::all(-1), const std::vector<char>& mask = std::vector<char>(), int flags = 0);
void drawKeypoints(const Mat& img, const std::vector<KeyPoint>& kp, Mat& out, const Scalar& color = Scalar::all(-1), int flags = 0);
class BFMatcher {
 public:
  explicit BFMatcher(int norm_type = NORM_L2, bool cross_check = false);
  void match(const Mat& q, const Mat& t, std::vector<DMatch>& m, const Mat& mask = Mat()) const;
  void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& m, int k, const Mat& mask = Mat(), bool compact = false) const;
 private:
  int norm_; bool cross_;
};
}
#endif
