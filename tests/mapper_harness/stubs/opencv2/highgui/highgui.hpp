#ifndef MAPPER_HARNESS_CV_HIGHGUI_
#define MAPPER_HARNESS_CV_HIGHGUI_
#include <opencv2/core/core.hpp>
namespace cv {
Mat imread(const std::string& path, int flags = 1);
bool imwrite(const std::string& path, const Mat& img, const std::vector<int>& params = std::vector<int>());
}
#endif
