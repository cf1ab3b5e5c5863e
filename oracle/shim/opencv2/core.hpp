// Minimal stand-in for the few OpenCV types the reference's ColourDifference.cpp, GridUtility.cpp
// and GridBounds.cpp touch, so those files can be compiled UNMODIFIED from /root/reference
// into oracle/_ref (test infrastructure only; see oracle/Makefile). Not OpenCV code.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>
typedef unsigned char uchar;
namespace cv {
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3 only"); val[0] = a; val[1] = b; val[2] = c; }
    template <typename U> Vec(const Vec<U, N> &o) { for (int i = 0; i < N; ++i) val[i] = static_cast<T>(o.val[i]); }
    const T &operator[](int i) const { return val[i]; }
    T &operator[](int i) { return val[i]; }
};
typedef Vec<double, 3> Vec3d;
typedef Vec<float, 3> Vec3f;
struct Point { int x = 0, y = 0; };
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
    Point tl() const { return {x, y}; }
    Point br() const { Point p; p.x = x + width; p.y = y + height; return p; }
    bool empty() const { return width <= 0 || height <= 0; }
};
// what GridBounds.cpp needs: equality and the bounding-box union (an empty operand contributes nothing)
inline bool operator==(const Rect &a, const Rect &b) { return a.x == b.x && a.y == b.y && a.width == b.width && a.height == b.height; }
inline Rect operator|(const Rect &a, const Rect &b)
{
    if (a.empty())
        return b;
    if (b.empty())
        return a;
    const int x1 = a.x < b.x ? a.x : b.x, y1 = a.y < b.y ? a.y : b.y;
    const int x2 = a.x + a.width > b.x + b.width ? a.x + a.width : b.x + b.width;
    const int y2 = a.y + a.height > b.y + b.height ? a.y + a.height : b.y + b.height;
    return Rect(x1, y1, x2 - x1, y2 - y1);
}
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct Range {
    int start = 0, end = 0;
    Range() {}
    Range(int s, int e) : start(s), end(e) {}
};
// Row-major 2-D array with shared storage: what CPUPhotomosaicGenerator.cpp and GridGenerator.cpp touch of cv::Mat (rows,
// cols, ptr<T>(row), empty(), channels(), sub-views by Range / Rect). Copies and views share the buffer, like cv::Mat headers.
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, size_t elem_bytes)
        : rows(r), cols(c), step_(c * elem_bytes), elem_(elem_bytes),
          buf_(new unsigned char[(size_t)r * c * elem_bytes + 1], std::default_delete<unsigned char[]>())
    {}
    Mat(const Mat &m, const Range &rowRange, const Range &colRange)
        : rows(rowRange.end - rowRange.start), cols(colRange.end - colRange.start), step_(m.step_), elem_(m.elem_),
          off_(m.off_ + (size_t)rowRange.start * m.step_ + (size_t)colRange.start * m.elem_), buf_(m.buf_)
    {}
    Mat(const Mat &m, const Rect &roi)
        : rows(roi.height), cols(roi.width), step_(m.step_), elem_(m.elem_),
          off_(m.off_ + (size_t)roi.y * m.step_ + (size_t)roi.x * m.elem_), buf_(m.buf_)
    {}
    bool empty() const { return rows <= 0 || cols <= 0 || !buf_; }
    int channels() const { return (int)elem_; }  // the shim only ever holds 8U images where channels() is asked
    size_t step() const { return step_; }
    unsigned char *data() { return buf_.get() + off_; }
    const unsigned char *data() const { return buf_.get() + off_; }
    template <typename T> T *ptr(int row = 0) { return reinterpret_cast<T *>(buf_.get() + off_ + (size_t)row * step_); }
    template <typename T> const T *ptr(int row = 0) const { return reinterpret_cast<const T *>(buf_.get() + off_ + (size_t)row * step_); }
private:
    size_t step_ = 0, elem_ = 0, off_ = 0;
    std::shared_ptr<unsigned char> buf_;
};
} // namespace cv
