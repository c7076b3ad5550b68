#ifndef MAPPER_HARNESS_CV_NONFREE_
#define MAPPER_HARNESS_CV_NONFREE_
#include <opencv2/features2d/features2d.hpp>
namespace cv {
class SURF {
 public:
  double hessianThreshold; int nOctaves, nOctaveLayers; bool extended, upright;
  SURF() : hessianThreshold(100), nOctaves(4), nOctaveLayers(2), extended(true), upright(false) {}
  SURF(double h, int o = 4, int l = 2, bool e = true, bool u = false) : hessianThreshold(h), nOctaves(o), nOctaveLayers(l), extended(e), upright(u) {}
  int descriptorSize() const { return extended ? 128 : 64; }
  void detect(const Mat& img, std::vector<KeyPoint>& kp, const Mat& mask = Mat()) const;
  void compute(const Mat& img, std::vector<KeyPoint>& kp, Mat& desc) const;
  void operator()(const Mat& img, const Mat& mask, std::vector<KeyPoint>& kp, Mat& desc, bool use_provided = false) const;
};
}
#endif
