// Stand-in for src/Other/ImageUtility.h. The reference's ImageUtility.cpp cannot be compiled here (QPixmap / QImage / QProgressBar,
// feature detectors, CUDA warping), so the functions the compiled sources call are RESTATED line by line in
// oracle/ref_generator_harness.cpp on top of the shim's cv::resize / cv::cvtColor (the real OpenCV through a callback). The edge-cell
// helpers (edgeDetect, matMakeTransparent) only feed the GUI's grid preview: they return a plain copy.
#pragma once
#include <opencv2/core.hpp>
#include "qt_standins.h"
namespace ImageUtility {
enum class ResizeType { INCLUSIVE, EXCLUSIVE, EXACT };  // ImageUtility.h:38
cv::Mat resizeImage(const cv::Mat &t_img, const int t_targetHeight, const int t_targetWidth, const ResizeType t_type);
void batchResizeMat(const std::vector<cv::Mat> &t_src, std::vector<cv::Mat> &t_dst, const int t_targetHeight, const int t_targetWidth,
                    const ResizeType t_type, QProgressBar *progressBar = nullptr);
bool batchResizeMat(std::vector<cv::Mat> &t_images, const double t_ratio = 0.5);
void matMakeTransparent(const cv::Mat &t_src, cv::Mat &t_dst, const int t_targetValue);
void edgeDetect(const cv::Mat &t_src, cv::Mat &t_dst);
void addAlphaChannel(std::vector<cv::Mat> &t_images);
double calculateEntropy(const cv::Mat &t_in, const cv::Mat &t_mask = cv::Mat());
[[maybe_unused]] const double MAX_ENTROPY = 8.0;  // ImageUtility.h:69
enum class SquareMethod { PAD, CROP };  // ImageUtility.h:74
void imageToSquare(cv::Mat &t_img, const SquareMethod t_method);
}
