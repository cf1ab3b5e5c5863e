// Host model: cell shapes, grid geometry, grid-state generation. See host_model.h for the reference map.
#include "host_model.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <cfloat>

namespace mm {

// ------------------------------------------------------------------ OpenCV-compatible 8U INTER_AREA

namespace {

struct Tap {
    int di, si;
    float alpha;
};

// cv::computeResizeAreaTab: fractional source coverage of every destination sample
std::vector<Tap> area_tab(int ssize, int dsize, int cn, double scale)
{
    std::vector<Tap> t;
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell_width = std::min(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3)
            t.push_back({dx * cn, (sx1 - 1) * cn, (float)((sx1 - fsx1) / cell_width)});
        for (int sx = sx1; sx < sx2; ++sx)
            t.push_back({dx * cn, sx * cn, (float)(1.0 / cell_width)});
        if (fsx2 - sx2 > 1e-3)
            t.push_back({dx * cn, sx2 * cn, (float)(std::min(std::min(fsx2 - sx2, 1.0), cell_width) / cell_width)});
    }
    return t;
}

inline uint8_t sat_u8(float v)
{
    const long r = lrintf(v);  // cvRound: round half to even in the default rounding mode
    return (uint8_t)std::min(255l, std::max(0l, r));
}

}  // namespace

AreaTable make_area_table(int ssize, int dsize)
{
    AreaTable t;
    const std::vector<Tap> taps = area_tab(ssize, dsize, 1, (double)ssize / dsize);
    t.start.assign(dsize + 1, 0);
    for (const Tap &tp : taps)
        t.start[tp.di + 1]++;
    for (int d = 0; d < dsize; ++d)
        t.start[d + 1] += t.start[d];
    for (const Tap &tp : taps) {  // taps are already ordered by destination, then source
        t.si.push_back(tp.si);
        t.alpha.push_back(tp.alpha);
    }
    return t;
}

bool resize_area_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw)
{
    if (dh <= 0 || dw <= 0 || sh < dh || sw < dw)
        return false;
    if (sh == dh && sw == dw) {
        memcpy(dst, src, (size_t)sh * sw * cn);
        return true;
    }
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    const int kx = (int)floor(scale_x + 0.5), ky = (int)floor(scale_y + 0.5);  // saturate_cast<int>(scale)
    const bool fast = fabs(scale_x - kx) < DBL_EPSILON && fabs(scale_y - ky) < DBL_EPSILON;
    if (fast) {
        // resizeAreaFast_: integer block sums; 2x2 uses the (s + 2) >> 2 vector path, otherwise sum * float(1/area)
        const int area = kx * ky;
        const float scale = 1.0f / (float)area;
        for (int y = 0; y < dh; ++y)
            for (int x = 0; x < dw; ++x)
                for (int c = 0; c < cn; ++c) {
                    int sum = 0;
                    for (int yy = 0; yy < ky; ++yy)
                        for (int xx = 0; xx < kx; ++xx)
                            sum += src[((size_t)(y * ky + yy) * sw + (x * kx + xx)) * cn + c];
                    dst[((size_t)y * dw + x) * cn + c] =
                        (kx == 2 && ky == 2 && (cn == 1 || cn == 3 || cn == 4)) ? (uint8_t)((sum + 2) >> 2) : sat_u8((float)sum * scale);
                }
        return true;
    }
    // resizeArea_: separable fractional coverage, float accumulation in table order
    const std::vector<Tap> xt = area_tab(sw, dw, cn, scale_x), yt = area_tab(sh, dh, 1, scale_y);
    std::vector<float> buf((size_t)dw * cn), sum((size_t)dw * cn, 0.0f);
    int prev_dy = yt.empty() ? 0 : yt[0].di;
    for (const Tap &ty : yt) {
        const uint8_t *S = src + (size_t)ty.si * sw * cn;
        std::fill(buf.begin(), buf.end(), 0.0f);
        for (const Tap &tx : xt)
            for (int c = 0; c < cn; ++c) {
                const float prod = (float)S[tx.si + c] * tx.alpha;
                buf[tx.di + c] = buf[tx.di + c] + prod;
            }
        if (ty.di != prev_dy) {
            for (int i = 0; i < dw * cn; ++i) {
                dst[(size_t)prev_dy * dw * cn + i] = sat_u8(sum[i]);
                sum[i] = ty.alpha * buf[i];
            }
            prev_dy = ty.di;
        } else {
            for (int i = 0; i < dw * cn; ++i) {
                const float prod = ty.alpha * buf[i];
                sum[i] = sum[i] + prod;
            }
        }
    }
    for (int i = 0; i < dw * cn; ++i)
        dst[(size_t)prev_dy * dw * cn + i] = sat_u8(sum[i]);
    return true;
}

// ------------------------------------------------------------------ INTER_CUBIC (8U)
//
// cv::resize(..., INTER_CUBIC) for 8U as OpenCV's own code computes it (imgproc/resize.cpp: interpolateCubic with
// A = -0.75 in float, coefficients rounded to 11-bit fixed point, horizontal pass in int, vertical pass
// VResizeCubicVec_32s8u in float for the first (row elements / 8) * 8 elements of a row and FixedPtCast integer
// arithmetic for the row tail). Probed against cv2 4.13 with IPP switched off (cv2.ipp.setUseIPP(False)): IPP builds
// replace this function by a closed float kernel whose results differ by +-1 LSB in ~5 % of the samples.
CubicTable make_cubic_table(int ssize, int dsize)
{
    CubicTable t;
    t.idx.resize((size_t)dsize * 4);
    t.coef.resize((size_t)dsize * 4);
    const double scale = 1.0 / ((double)dsize / ssize);  // resize(): scale_x = 1. / inv_scale_x
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        const int s0 = (int)floorf(f);
        f -= (float)s0;
        const float A = -0.75f;
        float c[4];
        c[0] = ((A * (f + 1) - 5 * A) * (f + 1) + 8 * A) * (f + 1) - 4 * A;
        c[1] = ((A + 2) * f - (A + 3)) * f * f + 1;
        c[2] = ((A + 2) * (1 - f) - (A + 3)) * (1 - f) * (1 - f) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
        for (int k = 0; k < 4; ++k) {
            t.idx[(size_t)d * 4 + k] = std::min(std::max(s0 - 1 + k, 0), ssize - 1);  // border: replicate
            t.coef[(size_t)d * 4 + k] = (int16_t)std::min(std::max(lrintf(c[k] * 2048.0f), -32768L), 32767L);
        }
    }
    return t;
}

bool resize_cubic_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw)
{
    if (sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0 || cn <= 0)
        return false;
    const CubicTable xt = make_cubic_table(sw, dw), yt = make_cubic_table(sh, dh);
    const int row = dw * cn, n_vec = row / 8 * 8;
    std::vector<int> h((size_t)sh * row);
    for (int y = 0; y < sh; ++y)
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                int v = 0;
                for (int k = 0; k < 4; ++k)
                    v += src[((size_t)y * sw + xt.idx[(size_t)x * 4 + k]) * cn + c] * xt.coef[(size_t)x * 4 + k];
                h[(size_t)y * row + x * cn + c] = v;
            }
    const float scale = 1.0f / (2048.0f * 2048.0f);
    for (int y = 0; y < dh; ++y) {
        const int *S[4];
        int16_t b[4];
        for (int k = 0; k < 4; ++k) {
            S[k] = h.data() + (size_t)yt.idx[(size_t)y * 4 + k] * row;
            b[k] = yt.coef[(size_t)y * 4 + k];
        }
        for (int i = 0; i < row; ++i)
            dst[(size_t)y * row + i] = cubic_vertical_u8(S[0][i], S[1][i], S[2][i], S[3][i], b, scale, i < n_vec);
    }
    return true;
}

void bgr_to_gray_u8(const uint8_t *bgr, size_t n, uint8_t *gray)
{
    // OpenCV RGB2Gray<uchar>: 15-bit fixed point, B 3735, G 19235, R 9798 (probed against cv2 4.13)
    for (size_t i = 0; i < n; ++i)
        gray[i] = (uint8_t)((bgr[3 * i] * 3735 + bgr[3 * i + 1] * 19235 + bgr[3 * i + 2] * 9798 + (1 << 14)) >> 15);
}

double masked_entropy(const uint8_t *gray, const uint8_t *mask, size_t n)
{
    if (n == 0)
        return 0;
    size_t hist[256] = {0}, count = 0;
    for (size_t i = 0; i < n; ++i)
        if (!mask || mask[i] != 0) {
            ++hist[gray[i]];
            ++count;
        }
    double e = 0;
    for (int b = 0; b < 256; ++b) {
        const double p = hist[b] / (double)count;
        if (p > 0)
            e -= p * log2(p);
    }
    return e;
}

// ------------------------------------------------------------------ Shape / Group

void Shape::set_mask(const uint8_t *m, int s)
{
    size = s;
    mask.resize((size_t)s * s);
    for (size_t i = 0; i < mask.size(); ++i)
        mask[i] = m[i] > 127 ? 255 : 0;
}

void Shape::set_mask_as_stored(const uint8_t *m, int s)
{
    size = s;
    mask.assign(m, m + (size_t)s * s);
}

std::vector<uint8_t> Shape::masks4() const
{
    std::vector<uint8_t> out((size_t)4 * size * size);
    for (int f = 0; f < 4; ++f)
        for (int y = 0; y < size; ++y)
            for (int x = 0; x < size; ++x) {
                const int sx = (f & 1) ? size - 1 - x : x, sy = (f & 2) ? size - 1 - y : y;
                out[((size_t)f * size + y) * size + x] = mask[(size_t)sy * size + sx];
            }
    return out;
}

bool Shape::resized(int new_size, Shape &out, std::string &err) const
{
    if (empty() || new_size == size) {
        out = *this;
        return true;
    }
    if (new_size < 1) {
        err = "cell size must be >= 1";
        return false;
    }
    // ImageUtility::resizeImage (ImageUtility.cpp:34-62): INTER_AREA when shrinking, INTER_CUBIC when growing
    std::vector<uint8_t> rm((size_t)new_size * new_size);
    if (new_size > size)
        resize_cubic_u8(mask.data(), size, size, 1, rm.data(), new_size, new_size);
    else
        resize_area_u8(mask.data(), size, size, 1, rm.data(), new_size, new_size);
    out = Shape();
    out.set_mask(rm.data(), new_size);  // CellShape(const cv::Mat&) -> setCellMask -> threshold
    const double ratio = (double)new_size / size;
    auto fl = [&](int v) { return (int)floor(v * ratio); };
    out.row_spacing = std::max(fl(row_spacing), 1);
    out.col_spacing = std::max(fl(col_spacing), 1);
    out.alt_row_spacing = std::max(fl(alt_row_spacing), 1);
    out.alt_col_spacing = std::max(fl(alt_col_spacing), 1);
    out.alt_row_offset = fl(alt_row_offset);
    out.alt_col_offset = fl(alt_col_offset);
    out.alt_col_flip_h = alt_col_flip_h;
    out.alt_col_flip_v = alt_col_flip_v;
    out.alt_row_flip_h = alt_row_flip_h;
    out.alt_row_flip_v = alt_row_flip_v;
    return true;
}

bool Group::build(const Shape &top, int detail_percent, int steps, std::string &err)
{
    if (top.empty()) {
        err = "cell shape has no mask";
        return false;
    }
    if (detail_percent < 1 || detail_percent > 100) {
        err = "detail must be in 1..100 percent";
        return false;
    }
    if (steps < 0) {
        err = "size steps must be >= 0";
        return false;
    }
    detail = detail_percent / 100.0;
    size_steps = steps;
    cells.assign(1, top);
    detail_cells.assign(1, Shape());
    if (!top.resized(std::max((int)(top.size * detail), 1), detail_cells[0], err))  // CellGroup.cpp:79-80
        return false;
    int size = top.size;
    for (int s = 1; s <= steps; ++s) {
        size /= 2;  // CellGroup.cpp:105
        if (size < 1) {
            err = "too many size steps for this cell size";
            return false;
        }
        Shape n, d;
        if (!cells[s - 1].resized(size, n, err))
            return false;
        if (!n.resized(std::max((int)(size * detail), 1), d, err))  // CellGroup.cpp:114
            return false;
        cells.push_back(n);
        detail_cells.push_back(d);
    }
    return true;
}

// ------------------------------------------------------------------ grid geometry

void grid_size(const Shape &s, int image_w, int image_h, int pad, int &gx, int &gy)
{
    if (s.col_spacing != s.alt_col_spacing)
        gx = 2 * ((image_w + s.col_spacing + s.alt_col_spacing - 1) / (s.col_spacing + s.alt_col_spacing));
    else
        gx = (image_w + s.col_spacing - 1) / s.col_spacing;
    if (s.row_spacing != s.alt_row_spacing)
        gy = 2 * ((image_h + s.row_spacing + s.alt_row_spacing - 1) / (s.row_spacing + s.alt_row_spacing));
    else
        gy = (image_h + s.row_spacing - 1) / s.row_spacing;
    gx += pad;
    gy += pad;
}

Rect rect_at(const Shape &s, int x, int y)
{
    // C++ integer division / remainder truncate toward zero, as in the reference (x, y may be negative)
    const int nx = x / 2, ax = x - nx;
    const int ny = y / 2, ay = y - ny;
    Rect r;
    r.x = (x < 0) ? ax * s.col_spacing + nx * s.alt_col_spacing : nx * s.col_spacing + ax * s.alt_col_spacing;
    if (y % 2 != 0)
        r.x += s.alt_row_offset;
    r.y = (y < 0) ? ay * s.row_spacing + ny * s.alt_row_spacing : ny * s.row_spacing + ay * s.alt_row_spacing;
    if (x % 2 != 0)
        r.y += s.alt_col_offset;
    r.w = r.h = s.size;
    return r;
}

int flip_at(const Shape &s, int x, int y)
{
    bool h = false, v = false;
    if (s.alt_col_flip_h && x % 2 != 0) h = !h;
    if (s.alt_row_flip_h && y % 2 != 0) h = !h;
    if (s.alt_col_flip_v && x % 2 != 0) v = !v;
    if (s.alt_row_flip_v && y % 2 != 0) v = !v;
    return (h ? 1 : 0) + (v ? 2 : 0);
}

static inline int clampi(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }

Rect detail_bound(const Shape &normal, int detail_size, double detail, int x, int y, int image_w, int image_h,
                  Rect *clamped_global, Rect *local_out)
{
    const Rect r = rect_at(normal, x, y);
    const int y0 = clampi(r.y, 0, image_h), y1 = clampi(r.y + r.h, 0, image_h);
    const int x0 = clampi(r.x, 0, image_w), x1 = clampi(r.x + r.w, 0, image_w);
    Rect local{x0 - r.x, y0 - r.y, x1 - x0, y1 - y0};
    if (clamped_global)
        *clamped_global = Rect{x0, y0, x1 - x0, y1 - y0};
    if (local_out)
        *local_out = local;
    Rect d;
    d.x = std::min(detail_size - 1, (int)(local.x * detail));
    d.y = std::min(detail_size - 1, (int)(local.y * detail));
    d.w = std::max(1, (int)(local.w * detail));
    d.h = std::max(1, (int)(local.h * detail));
    d.w = std::min(detail_size - d.x, d.w);
    d.h = std::min(detail_size - d.y, d.h);
    return d;
}

// ------------------------------------------------------------------ grid state

namespace {

Rect rect_union(const Rect &a, const Rect &b)
{
    // cv::Rect operator| : an empty operand yields the other one
    if (a.w <= 0 || a.h <= 0)
        return b;
    if (b.w <= 0 || b.h <= 0)
        return a;
    const int x0 = std::min(a.x, b.x), y0 = std::min(a.y, b.y);
    const int x1 = std::max(a.x + a.w, b.x + b.w), y1 = std::max(a.y + a.h, b.y + b.h);
    return Rect{x0, y0, x1 - x0, y1 - y0};
}
bool rect_eq(const Rect &a, const Rect &b) { return a.x == b.x && a.y == b.y && a.w == b.w && a.h == b.h; }

}  // namespace

// GridBounds::mergeBounds (GridBounds.cpp:39-104), same iteration order
void merge_bounds(std::vector<Rect> &b)
{
    bool merged_any = true;
    while (merged_any) {
        merged_any = false;
        for (size_t i = 0; i + 1 < b.size();) {
            for (size_t j = i + 1; j < b.size() && i + 1 < b.size();) {
                bool merge = false;
                if (b[i].x == b[j].x && b[i].w == b[j].w) {
                    const int yd = b[j].y - b[i].y;
                    merge = yd == 0 || (yd > 0 && yd <= b[i].h) || (yd < 0 && -yd <= b[j].h);
                } else if (b[i].y == b[j].y && b[i].h == b[j].h) {
                    const int xd = b[j].x - b[i].x;
                    merge = xd == 0 || (xd > 0 && xd <= b[i].w) || (xd < 0 && -xd <= b[j].w);
                } else {
                    const Rect u = rect_union(b[i], b[j]);
                    merge = rect_eq(u, b[i]) || rect_eq(u, b[j]);
                }
                if (merge) {
                    b[i] = rect_union(b[i], b[j]);
                    b.erase(b.begin() + j);
                    merged_any = true;
                } else
                    ++j;
            }
            if (i + 1 < b.size())
                ++i;
            else
                break;
        }
    }
}

namespace {

}  // namespace

bool compute_grid_state(const Group &g, int rows, int cols, const EntropyEvaluator &evaluate, std::vector<GridStep> &out,
                        std::string &err)
{
    out.clear();
    if (g.cells.empty() || g.cells[0].empty()) {
        err = "no cell group set";
        return false;
    }
    const int gh = rows, gw = cols;
    std::vector<Rect> active{Rect{0, 0, gw, gh}}, next;
    for (int step = 0; step <= g.size_steps && !active.empty(); ++step) {
        const Shape &shape = g.cells[step];
        const Shape &dshape = g.detail_cells[step];
        int gx, gy;
        grid_size(shape, gw, gh, kPadGrid, gx, gy);
        GridStep gs;
        gs.rows = gy;
        gs.cols = gx;
        gs.v.assign((size_t)gx * gy, -1);
        // cells that intersect an active bound, raster order (findCellState, GridGenerator.cpp:113-140)
        std::vector<EntropyCandidate> cand;
        for (int y = -kPadGrid; y < gy - kPadGrid; ++y)
            for (int x = -kPadGrid; x < gx - kPadGrid; ++x) {
                const Rect r = rect_at(shape, x, y);
                bool in_bounds = false;
                for (const Rect &b : active) {
                    const int y0 = clampi(r.y, b.y, b.y + b.h), y1 = clampi(r.y + r.h, b.y, b.y + b.h);
                    const int x0 = clampi(r.x, b.x, b.x + b.w), x1 = clampi(r.x + r.w, b.x, b.x + b.w);
                    if (y0 != y1 && x0 != x1) {
                        in_bounds = true;
                        break;
                    }
                }
                if (!in_bounds)
                    continue;
                EntropyCandidate c;
                c.x = x;
                c.y = y;
                c.db = detail_bound(shape, dshape.size, g.detail, x, y, gw, gh, &c.cg, nullptr);
                c.flip = flip_at(shape, x, y);
                if (c.cg.w > 0 && c.cg.h > 0 && (c.db.h > c.cg.h || c.db.w > c.cg.w)) {
                    err = "grid state: cell up-scaling (INTER_CUBIC) is not implemented";
                    return false;
                }
                cand.push_back(c);
            }
        // entropy rule only below the last size step (GridGenerator.cpp:143)
        std::vector<uint8_t> split(cand.size(), 0);
        if (step < g.size_steps && evaluate && !cand.empty())
            if (!evaluate(step, cand, split, err))
                return false;
        next.clear();
        for (size_t i = 0; i < cand.size(); ++i) {
            const EntropyCandidate &c = cand[i];
            if (split[i] && c.cg.w > 0 && c.cg.h > 0) {
                next.push_back(c.cg);  // the cell rect clamped to the image (GridGenerator.cpp:73-97)
            } else if (!split[i]) {
                gs.v[(size_t)(c.y + kPadGrid) * gx + (c.x + kPadGrid)] = 0;
            }
        }
        out.push_back(std::move(gs));
        active.swap(next);
        if (!active.empty())
            merge_bounds(active);
    }
    return true;
}

EntropyEvaluator host_entropy_evaluator(const Group &g, const uint8_t *main_bgr, size_t row_stride)
{
    return [&g, main_bgr, row_stride](int step, const std::vector<EntropyCandidate> &cand, std::vector<uint8_t> &split,
                                      std::string &) -> bool {
        const Shape &dshape = g.detail_cells[step];
        const std::vector<uint8_t> dmasks = dshape.masks4();
        std::vector<uint8_t> cell, small, gray, bmask;
        for (size_t i = 0; i < cand.size(); ++i) {
            const Rect &cg = cand[i].cg, &db = cand[i].db;
            if (cg.w <= 0 || cg.h <= 0)
                continue;  // empty image part: entropy 0 (ImageUtility.cpp:191-192)
            cell.resize((size_t)cg.w * cg.h * 3);
            for (int yy = 0; yy < cg.h; ++yy)
                memcpy(&cell[(size_t)yy * cg.w * 3], main_bgr + (size_t)(cg.y + yy) * row_stride + (size_t)cg.x * 3, (size_t)cg.w * 3);
            const uint8_t *img = cell.data();
            if (!(cg.h == db.h && cg.w == db.w)) {  // ImageUtility::resizeImage EXACT (ImageUtility.cpp:40-48)
                small.resize((size_t)db.w * db.h * 3);
                resize_area_u8(cell.data(), cg.h, cg.w, 3, small.data(), db.h, db.w);
                img = small.data();
            }
            gray.resize((size_t)db.h * db.w);
            bgr_to_gray_u8(img, gray.size(), gray.data());
            bmask.resize(gray.size());
            const uint8_t *m = dmasks.data() + (size_t)cand[i].flip * dshape.size * dshape.size;
            for (int yy = 0; yy < db.h; ++yy)
                memcpy(&bmask[(size_t)yy * db.w], m + (size_t)(db.y + yy) * dshape.size + db.x, db.w);
            split[i] = masked_entropy(gray.data(), bmask.data(), gray.size()) >= 8.0 * 0.7;  // MAX_ENTROPY * 0.7
        }
        return true;
    };
}

}  // namespace mm
