// Stand-in for src/Other/ImageUtility.h: CPUPhotomosaicGenerator.cpp only calls batchResizeMat(lib) between size steps
// (CPUPhotomosaicGenerator.cpp:95-99). The harness (oracle/ref_generator_harness.cpp) swaps in the next step's library,
// which the oracle's cv2 path has already halved (cv::resize is OpenCV, not reference code).
#pragma once
#include <opencv2/core.hpp>
#include "qt_standins.h"
namespace ImageUtility {
bool batchResizeMat(std::vector<cv::Mat> &t_images, const double t_ratio = 0.5);
}
