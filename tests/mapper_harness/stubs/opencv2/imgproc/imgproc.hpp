#ifndef MAPPER_HARNESS_CV_IMGPROC_
#define MAPPER_HARNESS_CV_IMGPROC_
#include <opencv2/core/core.hpp>
namespace cv {
void cvtColor(const Mat& src, Mat& dst, int code, int dst_cn = 0);
void resize(const Mat& src, Mat& dst, Size size, double fx = 0, double fy = 0, int interp = 1);
}
#endif
