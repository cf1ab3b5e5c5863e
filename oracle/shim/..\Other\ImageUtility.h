// Stand-in for src/Other/ImageUtility.h. CPUPhotomosaicGenerator.cpp only calls batchResizeMat(lib) between size steps
// (CPUPhotomosaicGenerator.cpp:95-99). The harness (oracle/ref_generator_harness.cpp) swaps in the next step's library,
// which the oracle's cv2 path has already halved (cv::resize is OpenCV, not reference code).
#pragma once
#include <opencv2/core.hpp>
#include "qt_standins.h"
namespace ImageUtility {
bool batchResizeMat(std::vector<cv::Mat> &t_images, const double t_ratio = 0.5);
// GridGenerator.cpp (:181-186) resizes the cell to the mask size and asks for its masked entropy. Both are OpenCV arithmetic
// (cv::resize, cvtColor BGR2GRAY) in the reference; the harness forwards the pair to a callback that evaluates them with cv2.
enum class ResizeType { INCLUSIVE, EXCLUSIVE, EXACT };  // ImageUtility.h:38
cv::Mat resizeImage(const cv::Mat &t_img, const int t_targetHeight, const int t_targetWidth, const ResizeType t_type);
double calculateEntropy(const cv::Mat &t_in, const cv::Mat &t_mask = cv::Mat());
[[maybe_unused]] const double MAX_ENTROPY = 8.0;  // ImageUtility.h:69
}
