// Minimal stand-in for the few OpenCV types the reference's ColourDifference.cpp and
// GridUtility.cpp touch, so those files can be compiled UNMODIFIED from /root/reference
// into oracle/_ref (test infrastructure only; see oracle/Makefile). Not OpenCV code.
#pragma once
#include <cmath>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>
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
    Point tl() const { return {x, y}; }
    Point br() const { Point p; p.x = x + width; p.y = y + height; return p; }
};
} // namespace cv
