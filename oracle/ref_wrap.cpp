// C entry points over the reference's own ColourDifference.cpp, GridUtility.cpp and GridBounds.cpp, which
// oracle/Makefile compiles unmodified from /root/reference into oracle/_ref/libref_core.so.
// Test infrastructure only: used to validate the oracle's restatement (tests/test_oracle_ref.py).
#include <optional>
#include "ColourDifference.h"
#include "GridBounds.h"
#include "GridUtility.h"
#include "ref_group.h"

static CellShape make_shape(const int *p) { return ref_make_shape(p, nullptr); }

extern "C" {
double ref_rgb_euclidean(const double *a, const double *b)
{
    return ColourDifference::calculateRGBEuclidean(cv::Vec3d(a[0], a[1], a[2]), cv::Vec3d(b[0], b[1], b[2]));
}
double ref_ciede2000(const double *a, const double *b)
{
    return ColourDifference::calculateCIEDE2000(cv::Vec3d(a[0], a[1], a[2]), cv::Vec3d(b[0], b[1], b[2]));
}
// same call the generator makes: a std::function taking Vec3d fed with Vec3f pixels
double ref_diff_f32(int type, const float *a, const float *b)
{
    const auto f = ColourDifference::getFunction(static_cast<ColourDifference::Type>(type));
    return f(cv::Vec3f(a[0], a[1], a[2]), cv::Vec3f(b[0], b[1], b[2]));
}
void ref_grid_size(const int *shape, int w, int h, int pad, int *gx, int *gy)
{
    const cv::Point p = GridUtility::calculateGridSize(make_shape(shape), w, h, pad);
    *gx = p.x; *gy = p.y;
}
void ref_rect_at(const int *shape, int x, int y, int *rect)
{
    const cv::Rect r = GridUtility::getRectAt(make_shape(shape), x, y);
    rect[0] = r.x; rect[1] = r.y; rect[2] = r.width; rect[3] = r.height;
}
int ref_flip_at(const int *shape, int x, int y)
{
    const auto f = GridUtility::getFlipStateAt(make_shape(shape), x, y);
    return (f.horizontal ? 1 : 0) + (f.vertical ? 2 : 0);
}
// GridBounds::addBound x n, mergeBounds; rects are x, y, w, h quadruples. Returns the number of merged bounds written.
int ref_merge_bounds(const int *rects, int n, int *out, int out_capacity)
{
    GridBounds b;
    for (int i = 0; i < n; ++i)
        b.addBound(cv::Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]));
    if (!b.empty())
        b.mergeBounds();
    int m = 0;
    for (auto it = b.cbegin(); it != b.cend(); ++it, ++m)
        if (m < out_capacity) {
            out[4 * m] = it->x; out[4 * m + 1] = it->y; out[4 * m + 2] = it->width; out[4 * m + 3] = it->height;
        }
    return m;
}
}
