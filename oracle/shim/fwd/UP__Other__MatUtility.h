// Stand-in for src/Other/MatUtility.h (parallel element-wise visitors over two / three cv::Mat_ built on OpenCV internals:
// ParallelLoopBody, parallel_for_, n-dimensional indexing). Restated as plain row / column loops over 2-D mats -- the functor
// the reference passes is what matters, and that stays the reference's own code (ColourScheme.cpp:71-75, 96-100, 122-127, 151-156).
#pragma once
#include <opencv2/core.hpp>
namespace MatUtility {
template <typename T, typename F> inline void forEach_2_impl(cv::Mat_<T> *const m1, cv::Mat_<T> *const m2, const F &op)
{
    for (int y = 0; y < m1->rows; ++y)
        for (int x = 0; x < m1->cols; ++x) {
            const int pos[2] = {y, x};
            op(m1->template ptr<T>(y)[x], m2->template ptr<T>(y)[x], pos);
        }
}
template <typename T, typename F> inline void forEach_3_impl(cv::Mat_<T> *const m1, cv::Mat_<T> *const m2, cv::Mat_<T> *const m3, const F &op)
{
    for (int y = 0; y < m1->rows; ++y)
        for (int x = 0; x < m1->cols; ++x) {
            const int pos[2] = {y, x};
            op(m1->template ptr<T>(y)[x], m2->template ptr<T>(y)[x], m3->template ptr<T>(y)[x], pos);
        }
}
}
