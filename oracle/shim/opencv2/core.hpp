// Minimal stand-in for the OpenCV types and calls the reference's generator sources touch (ColourDifference.cpp,
// GridUtility.cpp, GridBounds.cpp, GridGenerator.cpp, CPUPhotomosaicGenerator.cpp, PhotomosaicGeneratorBase.cpp), so those
// files can be compiled UNMODIFIED from /root/reference into oracle/_ref (test infrastructure only; see oracle/Makefile).
// Not OpenCV code: cv::Mat here is a typed 2-D array with shared storage and views; anything that is OpenCV ARITHMETIC
// (cvtColor, resize) is forwarded through a callback to the real OpenCV (cv2) by oracle/oracle.py.
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
typedef Vec<unsigned char, 4> Vec4b;
struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
    Point2f &operator+=(const Point2f &o) { x += o.x; y += o.y; return *this; }
};
struct Point3f { float x = 0, y = 0, z = 0; };
struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
    Rect(const Point &p, const Size &s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    int area() const { return width * height; }
    bool contains(const Point2f &p) const { return p.x >= x && p.x < x + width && p.y >= y && p.y < y + height; }
    Point tl() const { return {x, y}; }
    Point br() const { Point p; p.x = x + width; p.y = y + height; return p; }
    bool empty() const { return width <= 0 || height <= 0; }
};
// what GridBounds.cpp needs: equality and the bounding-box union (an empty operand contributes nothing)
inline bool operator==(const Rect &a, const Rect &b) { return a.x == b.x && a.y == b.y && a.width == b.width && a.height == b.height; }
inline Rect operator&(const Rect &a, const Rect &b)
{
    const int x1 = a.x > b.x ? a.x : b.x, y1 = a.y > b.y ? a.y : b.y;
    const int x2 = a.x + a.width < b.x + b.width ? a.x + a.width : b.x + b.width;
    const int y2 = a.y + a.height < b.y + b.height ? a.y + a.height : b.y + b.height;
    return (x2 <= x1 || y2 <= y1) ? Rect() : Rect(x1, y1, x2 - x1, y2 - y1);
}
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
    double operator[](int i) const { return val[i]; }
};
struct Range {
    int start = 0, end = 0;
    Range() {}
    Range(int s, int e) : start(s), end(e) {}
};
} // namespace cv

// element type codes: the public encoding of the OpenCV API (depth in the low 3 bits, channels - 1 above)
#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(flags) ((flags) & 7)
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)

namespace cv {
// Typed row-major 2-D array with shared storage. Copies of a Mat and views (Range / Rect) share the buffer, like cv::Mat
// headers do -- which is what makes std::vector<cv::Mat>(n, cv::Mat(...)) alias one buffer.
// Public members follow the names callers written against OpenCV use: rows, cols, data, step.
class Mat {
public:
    struct Step {
        size_t v = 0;
        operator size_t() const { return v; }
    };
    // `m.size` is an object in OpenCV (compared with != and called as m.size()): it looks at its owner's rows / cols
    struct MatSize {
        const Mat *owner;
        explicit MatSize(const Mat *o) : owner(o) {}
        Size operator()() const { return Size(owner->cols, owner->rows); }
        bool operator==(const MatSize &o) const { return owner->rows == o.owner->rows && owner->cols == o.owner->cols; }
        bool operator!=(const MatSize &o) const { return !(*this == o); }
    };
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;  // first element of this header (view offset applied)
    Step step;                      // bytes per row
    MatSize size{this};
    Mat() {}
    Mat(const Mat &m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), type_(m.type_), buf_(m.buf_) {}
    Mat &operator=(const Mat &m)
    {
        rows = m.rows; cols = m.cols; data = m.data; step = m.step; type_ = m.type_; buf_ = m.buf_;
        return *this;
    }
    Mat(int r, int c, int type) { create(r, c, type); }
    // header over caller-owned memory (CustomQDataStream.h:70 reads RAW mats this way, then clones)
    Mat(int r, int c, int type, void *external) : rows(r), cols(c), data(static_cast<unsigned char *>(external)), type_(type)
    {
        step.v = (size_t)c * elemSize();
    }
    Mat(int r, int c, int type, const Scalar &s)
    {
        create(r, c, type);
        const int cn = channels();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols * cn; ++x) {
                const double v = s.val[x % cn];
                if (depth() == CV_8U)
                    ptr<unsigned char>(y)[x] = (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : std::lrint(v));
                else
                    ptr<float>(y)[x] = (float)v;
            }
    }
    Mat(Size sz, int type, const Scalar &s) : Mat(sz.height, sz.width, type, s) {}
    Mat(const Mat &m, const Range &rowRange, const Range &colRange)
        : rows(rowRange.end - rowRange.start), cols(colRange.end - colRange.start),
          data(m.data ? m.data + (size_t)rowRange.start * m.step + (size_t)colRange.start * m.elemSize() : nullptr), step(m.step),
          type_(m.type_), buf_(m.buf_)
    {}
    Mat(const Mat &m, const Rect &roi)
        : rows(roi.height), cols(roi.width), data(m.data ? m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize() : nullptr),
          step(m.step), type_(m.type_), buf_(m.buf_)
    {}
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0, 0, 0, 0)); }
    void create(int r, int c, int type)
    {
        rows = r;
        cols = c;
        type_ = type;
        step.v = (size_t)c * elemSize();
        buf_.reset(new unsigned char[(size_t)r * step.v + 1], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    bool empty() const { return rows <= 0 || cols <= 0 || !data; }
    bool isContinuous() const { return step.v == (size_t)cols * elemSize(); }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return (size_t)(depth() == CV_8U ? 1 : 4) * channels(); }
    unsigned char *ptr(int row = 0) { return data + (size_t)row * step.v; }
    const unsigned char *ptr(int row = 0) const { return data + (size_t)row * step.v; }
    template <typename T> T *ptr(int row = 0) { return reinterpret_cast<T *>(data + (size_t)row * step.v); }
    template <typename T> const T *ptr(int row = 0) const { return reinterpret_cast<const T *>(data + (size_t)row * step.v); }
    Mat operator()(const Range &rowRange, const Range &colRange) const { return Mat(*this, rowRange, colRange); }
    Mat operator()(const Rect &roi) const { return Mat(*this, roi); }
    Mat clone() const
    {
        Mat m;
        if (!empty()) {
            m.create(rows, cols, type_);
            for (int y = 0; y < rows; ++y)
                std::memcpy(m.ptr<unsigned char>(y), ptr<unsigned char>(y), (size_t)cols * elemSize());
        }
        return m;
    }
    // copyTo: into the destination's existing storage when it already has this size and type (that is how writing
    // through a view works), into fresh storage otherwise; with a mask only the elements whose mask byte is non-zero
    void copyTo(Mat &dst) const { copy_impl(dst, nullptr); }
    void copyTo(Mat &&dst) const { copy_impl(dst, nullptr); }
    void copyTo(Mat &dst, const Mat &mask) const { copy_impl(dst, &mask); }
    void copyTo(Mat &&dst, const Mat &mask) const { copy_impl(dst, &mask); }
    // convertTo between 8U and 32F with OpenCV's arithmetic for these pairs: to 32F dst = float(src) * float(alpha) + float(beta);
    // to 8U saturate_cast<uchar> = round half to even, then clamp (alpha / beta only ever 1 / 0 on that side here)
    void convertTo(Mat &dst, int rtype, double alpha = 1, double beta = 0) const
    {
        const int ddepth = CV_MAT_DEPTH(rtype);
        if ((depth() != CV_8U && depth() != CV_32F) || (ddepth != CV_8U && ddepth != CV_32F))
            throw std::invalid_argument("shim convertTo: only 8U / 32F");
        Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
        const float a = (float)alpha, b = (float)beta;
        const bool scale = alpha != 1 || beta != 0;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols * channels(); ++x) {
                float v = depth() == CV_8U ? (float)ptr<unsigned char>(y)[x] : ptr<float>(y)[x];
                if (scale)
                    v = v * a + b;
                if (ddepth == CV_32F)
                    out.ptr<float>(y)[x] = v;
                else {
                    const long r = std::lrint(v);  // default rounding mode: to nearest, ties to even
                    out.ptr<unsigned char>(y)[x] = (unsigned char)(r < 0 ? 0 : r > 255 ? 255 : r);
                }
            }
        dst = out;
    }
    // element-wise visitor (ColourScheme.cpp:49: forEach<cv::Point3f>)
    template <typename T, typename F> void forEach(const F &op)
    {
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                const int pos[2] = {y, x};
                op(ptr<T>(y)[x], pos);
            }
    }

private:
    void copy_impl(Mat &dst, const Mat *mask) const
    {
        if (dst.empty() || dst.rows != rows || dst.cols != cols || dst.type_ != type_) {
            dst.create(rows, cols, type_);
            if (mask)
                for (int y = 0; y < rows; ++y)
                    std::memset(dst.ptr<unsigned char>(y), 0, (size_t)cols * elemSize());
        }
        const size_t es = elemSize();
        for (int y = 0; y < rows; ++y) {
            const unsigned char *s = ptr<unsigned char>(y);
            unsigned char *d = dst.ptr<unsigned char>(y);
            if (!mask) {
                std::memmove(d, s, (size_t)cols * es);
                continue;
            }
            const unsigned char *m = mask->ptr<unsigned char>(y);
            for (int x = 0; x < cols; ++x)
                if (m[x])
                    std::memcpy(d + x * es, s + x * es, es);
        }
    }
    int type_ = 0;
    std::shared_ptr<unsigned char> buf_;
};

// Typed header (ColourScheme.cpp uses cv::Mat_<cv::Point3f>): constructing it from a Mat of another element type converts,
// like OpenCV's Mat_(const Mat&) does through convertTo
template <typename T> struct DataTypeOf;
template <> struct DataTypeOf<Point3f> { enum { type = CV_32FC3 }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataTypeOf<T>::type) {}
    Mat_(const Mat &m)
    {
        if (m.type() == DataTypeOf<T>::type)
            Mat::operator=(m);
        else {
            Mat tmp;
            m.convertTo(tmp, DataTypeOf<T>::type);
            Mat::operator=(tmp);
        }
    }
};

// conversion / interpolation codes: the public values of the OpenCV API, handed unchanged to cv2 by the callback
enum ColorConversionCodes { COLOR_BGR2BGRA = 0, COLOR_BGR2GRAY = 6, COLOR_GRAY2RGBA = 9, COLOR_BGR2Lab = 44, COLOR_BGR2HSV_FULL = 66, COLOR_HSV2BGR_FULL = 70 };
enum InterpolationFlags { INTER_CUBIC = 2, INTER_AREA = 3 };
// OpenCV arithmetic: evaluated by the real OpenCV through the callback the harness installs (oracle/ref_generator_harness.cpp)
void cvtColor(const Mat &src, Mat &dst, int code);
void resize(const Mat &src, Mat &dst, Size dsize, double fx = 0, double fy = 0, int interpolation = 1);
// PNG container of .mcs / .mil (CustomQDataStream.h:38-41, 76-81): the real OpenCV codec through the harness callback
enum ImreadModes { IMREAD_UNCHANGED = -1 };
bool imencode(const std::string &ext, const Mat &img, std::vector<uchar> &buf);
Mat imdecode(const std::vector<uchar> &buf, int flags);
// the 8U helpers CellShape.cpp uses on masks: binary threshold, flips, inequality + sum (operator==)
enum ThresholdTypes { THRESH_BINARY = 0 };
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int)
{
    Mat out(src.rows, src.cols, src.type());
    for (int y = 0; y < src.rows; ++y)
        for (size_t x = 0; x < (size_t)src.cols * src.elemSize(); ++x)
            out.ptr<unsigned char>(y)[x] = src.ptr<unsigned char>(y)[x] > thresh ? (unsigned char)maxval : 0;
    dst = out;
    return thresh;
}
inline void flip(const Mat &src, Mat &dst, int flipCode)  // 0: around the x axis (rows), > 0: around the y axis, < 0: both
{
    Mat out(src.rows, src.cols, src.type());
    const size_t es = src.elemSize();
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            const int sy = flipCode <= 0 ? src.rows - 1 - y : y, sx = flipCode != 0 ? src.cols - 1 - x : x;
            std::memcpy(out.ptr<unsigned char>(y) + x * es, src.ptr<unsigned char>(sy) + sx * es, es);
        }
    dst = out;
}
inline Mat operator!=(const Mat &a, const Mat &b)
{
    Mat out(a.rows, a.cols, CV_MAKETYPE(CV_8U, a.channels()));
    for (int y = 0; y < a.rows; ++y)
        for (size_t x = 0; x < (size_t)a.cols * a.elemSize(); ++x)
            out.ptr<unsigned char>(y)[x] = a.ptr<unsigned char>(y)[x] != b.ptr<unsigned char>(y)[x] ? 255 : 0;
    return out;
}
inline Scalar sum(const Mat &m)
{
    Scalar s;
    const int cn = m.channels();
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols * cn; ++x)
            s.val[x % cn] += m.ptr<unsigned char>(y)[x];
    return s;
}
// ---- what else ImageUtility.cpp mentions. On the generator path: copyMakeBorder (imageToSquare PAD), split / merge
// (addAlphaChannel). Off the path (GUI edge overlay, feature / face based cropping): simple or refusing stand-ins.
enum BorderTypes { BORDER_CONSTANT = 0 };
inline void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int, const Scalar &value = Scalar())
{
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type(), value);
    src.copyTo(out(Rect(left, top, src.cols, src.rows)));
    dst = out;
}
inline void split(const Mat &m, std::vector<Mat> &channels)
{
    const int cn = m.channels();
    channels.assign(cn, Mat());
    for (int c = 0; c < cn; ++c) {
        channels[c].create(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
        for (int y = 0; y < m.rows; ++y)
            for (int x = 0; x < m.cols; ++x)
                channels[c].ptr<unsigned char>(y)[x] = m.ptr<unsigned char>(y)[x * cn + c];
    }
}
inline void merge(const std::vector<Mat> &channels, Mat &dst)
{
    const int cn = (int)channels.size();
    Mat out(channels.at(0).rows, channels.at(0).cols, CV_MAKETYPE(channels.at(0).depth(), cn));
    for (int c = 0; c < cn; ++c)
        for (int y = 0; y < out.rows; ++y)
            for (int x = 0; x < out.cols; ++x)
                out.ptr<unsigned char>(y)[x * cn + c] = channels[c].ptr<unsigned char>(y)[x];
    dst = out;
}
inline void bitwise_and(const Mat &a, const Mat &b, Mat &dst)
{
    Mat out(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (size_t x = 0; x < (size_t)a.cols * a.elemSize(); ++x)
            out.ptr<unsigned char>(y)[x] = a.ptr<unsigned char>(y)[x] & b.ptr<unsigned char>(y)[x];
    dst = out;
}
// 8U single channel, small float kernel, zero border: enough for the 3 x 3 edge kernel of ImageUtility::edgeDetect
inline void filter2D(const Mat &src, Mat &dst, int, const Mat &kernel, Point = Point(-1, -1), double delta = 0, int = BORDER_CONSTANT)
{
    Mat out(src.rows, src.cols, src.type());
    const int kr = kernel.rows / 2, kc = kernel.cols / 2;
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            double acc = delta;
            for (int j = 0; j < kernel.rows; ++j)
                for (int i = 0; i < kernel.cols; ++i) {
                    const int sy = y + j - kr, sx = x + i - kc;
                    if (sy >= 0 && sy < src.rows && sx >= 0 && sx < src.cols)
                        acc += kernel.ptr<float>(j)[i] * src.ptr<unsigned char>(sy)[sx];
                }
            const long r = std::lrint(acc);
            out.ptr<unsigned char>(y)[x] = (unsigned char)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    dst = out;
}
enum MorphShapes { MORPH_RECT = 0 };
inline Mat getStructuringElement(int, Size ksize) { return Mat(ksize.height, ksize.width, CV_8UC1, Scalar(1)); }
inline void dilate(const Mat &src, Mat &dst, const Mat &kernel)
{
    Mat out(src.rows, src.cols, src.type());
    const int kr = kernel.rows / 2, kc = kernel.cols / 2;
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            unsigned char m = 0;
            for (int sy = std::max(0, y - kr); sy <= std::min(src.rows - 1, y + kr); ++sy)
                for (int sx = std::max(0, x - kc); sx <= std::min(src.cols - 1, x + kc); ++sx)
                    m = std::max(m, src.ptr<unsigned char>(sy)[sx]);
            out.ptr<unsigned char>(y)[x] = m;
        }
    dst = out;
}
inline void equalizeHist(const Mat &, Mat &) { throw std::runtime_error("equalizeHist: not available in the checker"); }
struct KeyPoint {
    Point2f pt;
};
class FastFeatureDetector {
public:
    static std::shared_ptr<FastFeatureDetector> create() { return std::make_shared<FastFeatureDetector>(); }
    void detect(const Mat &, std::vector<KeyPoint> &) { throw std::runtime_error("FastFeatureDetector: not available in the checker"); }
};
class CascadeClassifier {
public:
    bool empty() const { return true; }
    void detectMultiScale(const Mat &, std::vector<Rect> &) { throw std::runtime_error("CascadeClassifier: not available in the checker"); }
};
// 8U element-wise logic (buildPhotomosaic's coverage masks)
inline void bitwise_or(const Mat &a, const Mat &b, Mat &dst)
{
    Mat out = (dst.rows == a.rows && dst.cols == a.cols && dst.type() == a.type() && !dst.empty()) ? dst : Mat(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (size_t x = 0; x < (size_t)a.cols * a.elemSize(); ++x)
            out.ptr<unsigned char>(y)[x] = a.ptr<unsigned char>(y)[x] | b.ptr<unsigned char>(y)[x];
    dst = out;
}
inline void bitwise_not(const Mat &a, Mat &dst)
{
    Mat out = (dst.rows == a.rows && dst.cols == a.cols && dst.type() == a.type() && !dst.empty()) ? dst : Mat(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; ++y)
        for (size_t x = 0; x < (size_t)a.cols * a.elemSize(); ++x)
            out.ptr<unsigned char>(y)[x] = (unsigned char)~a.ptr<unsigned char>(y)[x];
    dst = out;
}
} // namespace cv
