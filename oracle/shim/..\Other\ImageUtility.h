// Stand-in for src/Other/ImageUtility.h. The reference's ImageUtility.cpp cannot be compiled here (Qt GUI types, CUDA
// warping), so the five functions the generator sources call are RESTATED line by line in oracle/ref_generator_harness.cpp
// on top of the shim's cv::resize / cv::cvtColor (which are the real OpenCV through a callback).
#pragma once
#include <opencv2/core.hpp>
#include "qt_standins.h"
namespace ImageUtility {
enum class ResizeType { INCLUSIVE, EXCLUSIVE, EXACT };  // ImageUtility.h:38
cv::Mat resizeImage(const cv::Mat &t_img, const int t_targetHeight, const int t_targetWidth, const ResizeType t_type);
bool batchResizeMat(std::vector<cv::Mat> &t_images, const double t_ratio = 0.5);
void addAlphaChannel(std::vector<cv::Mat> &t_images);
double calculateEntropy(const cv::Mat &t_in, const cv::Mat &t_mask = cv::Mat());
[[maybe_unused]] const double MAX_ENTROPY = 8.0;  // ImageUtility.h:69
}
