/*
 * mosaic_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's CPU photomosaic best-fit path. Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file; the product (mosaicmagnifique_b200/csrc) never does.
 *
 * What is restated (reference = /root/reference, MorganGrundy/MosaicMagnifique):
 *   - colour differences            src/Photomosaic/ColourDifference.cpp:28-158
 *   - grid geometry                 src/Grid/GridUtility.cpp:25-52, 86-132
 *   - repeat penalties              src/Photomosaic/CPUPhotomosaicGenerator.cpp:185-225
 *   - per-cell best fit             src/Photomosaic/CPUPhotomosaicGenerator.cpp:116-181
 *   - raster loop over the grid     src/Photomosaic/CPUPhotomosaicGenerator.cpp:49-100
 *   - detail-space cell bounds      src/Photomosaic/PhotomosaicGeneratorBase.cpp:293-329
 *   - entropy                       src/Other/ImageUtility.cpp:189-242
 *
 * Arithmetic follows the reference: pixels are f32, every difference and every
 * sum is f64 (cv::Vec3f -> cv::Vec3d at the std::function call,
 * ColourDifference.h:31).
 *
 * Parity pinning: mo_rgb_euclidean / mo_ciede2000 are checked against the 48
 * known-answer vectors of test/tst_ColourDifference.h:26-108 (tests/test_oracle_*.py)
 * and against the reference's own ColourDifference.cpp compiled unmodified
 * (oracle/_ref, see oracle/Makefile). The OpenCV arithmetic either side of this
 * core (Lab conversion, INTER_AREA) is unpinned by the reference's tests; the
 * oracle uses cv2 4.13 for it (oracle/oracle.py) -- see DESIGN.md.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define MO_PAD_GRID 2 /* GridUtility.h:33 */

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ colour */

/* ColourDifference.cpp:28-33 (also CIE76, ColourDifference.h:41) */
double mo_rgb_euclidean(const double a[3], const double b[3])
{
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
}

static double mo_rad(double deg) { return (deg * M_PI) / 180.0; } /* :36-39 */

/* hue angle in [0, 2pi), 0 for the achromatic point -- ColourDifference.cpp:61-85 */
static double mo_hue(double b, double a_prime)
{
    if (b == 0.0 && a_prime == 0.0)
        return 0.0;
    double h = atan2(b, a_prime);
    if (h < 0.0)
        h += mo_rad(360.0);
    return h;
}

/* ColourDifference.cpp:42-158; x = (L, a, b) of the first colour, y of the second */
double mo_ciede2000(const double x[3], const double y[3])
{
    const double two_pi = mo_rad(360.0), pi = mo_rad(180.0);
    const double k25_7 = 6103515625.0; /* :46 */

    /* step 1: chroma, G, a', C', h' (:48-85) */
    const double c1 = sqrt(x[1] * x[1] + x[2] * x[2]);
    const double c2 = sqrt(y[1] * y[1] + y[2] * y[2]);
    const double cbar = (c1 + c2) / 2.0;
    const double cbar7 = cbar * cbar * cbar * cbar * cbar * cbar * cbar;
    const double g = 0.5 * (1.0 - sqrt(cbar7 / (cbar7 + k25_7)));
    const double a1p = (1.0 + g) * x[1];
    const double a2p = (1.0 + g) * y[1];
    const double c1p = sqrt(a1p * a1p + x[2] * x[2]);
    const double c2p = sqrt(a2p * a2p + y[2] * y[2]);
    const double h1p = mo_hue(x[2], a1p);
    const double h2p = mo_hue(y[2], a2p);

    /* step 2: deltas (:87-102) */
    const double dl = y[0] - x[0];
    const double dc = c2p - c1p;
    const double cprod = c1p * c2p;
    double dh = 0.0;
    if (cprod != 0.0) {
        dh = h2p - h1p;
        if (dh < -pi)
            dh += two_pi;
        else if (dh > pi)
            dh -= two_pi;
    }
    const double dH = 2.0 * sqrt(cprod) * sin(dh / 2.0);

    /* step 3: means, weights (:104-146) */
    const double lbar = (x[0] + y[0]) / 2.0;
    const double cpbar = (c1p + c2p) / 2.0;
    const double hsum = h1p + h2p;
    double hbar;
    if (c1p * c2p == 0.0)
        hbar = hsum;
    else if (fabs(h1p - h2p) <= pi)
        hbar = hsum / 2.0;
    else if (hsum < two_pi)
        hbar = (hsum + two_pi) / 2.0;
    else
        hbar = (hsum - two_pi) / 2.0;

    const double t = 1.0 - (0.17 * cos(hbar - mo_rad(30.0))) + (0.24 * cos(2.0 * hbar)) +
                     (0.32 * cos((3.0 * hbar) + mo_rad(6.0))) - (0.20 * cos((4.0 * hbar) - mo_rad(63.0)));
    /* the reference expands ((hbar-275)/25)^2 as (hbar*(hbar-550)+275^2)/25^2, all in radians (:132-133) */
    const double dtheta = mo_rad(30.0) *
        exp(-((hbar * (hbar - mo_rad(550.0)) + (mo_rad(275.0) * mo_rad(275.0))) / (mo_rad(25.0) * mo_rad(25.0))));
    const double cpbar7 = cpbar * cpbar * cpbar * cpbar * cpbar * cpbar * cpbar;
    const double rc = 2.0 * sqrt(cpbar7 / (cpbar7 + k25_7));
    /* (lbar-50)^2 written as lbar*(lbar-100)+2500 (:138-139) */
    const double sl = 1.0 + ((0.015 * (lbar * (lbar - 100.0) + 2500.0)) / sqrt(20 + lbar * (lbar - 100.0) + 2500.0));
    const double sc = 1.0 + (0.045 * cpbar);
    const double sh = 1.0 + (0.015 * cpbar * t);
    const double rt = (-sin(2.0 * dtheta)) * rc;

    /* :153-157 (kL = kC = kH = 1) */
    return sqrt((dl * dl) / (sl * sl) + (dc * dc) / (sc * sc) + (dH * dH) / (sh * sh) +
                (rt * (dc / sc) * (dH / sh)));
}

/* type: 0 RGB_EUCLIDEAN, 1 CIE76, 2 CIEDE2000 (ColourDifference.h:13-19, .cpp:16-25) */
static double mo_diff_f32(int type, const float *p, const float *q)
{
    const double a[3] = {p[0], p[1], p[2]}, b[3] = {q[0], q[1], q[2]};
    return type == 2 ? mo_ciede2000(a, b) : mo_rgb_euclidean(a, b);
}

/* batch per-pixel differences (mirrors the reference's *_CPUvsBatchCUDA tests,
 * test/tst_ColourDifference.h:389-470): out[i] = diff(a[i], b[i]) */
void mo_diff_batch(int type, const float *a, const float *b, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = mo_diff_f32(type, a + 3 * i, b + 3 * i);
}

/* ------------------------------------------------------------------ grid geometry */

typedef struct {
    int size;           /* mask rows (CellShape::getSize, CellShape.cpp:152) */
    int row_spacing, col_spacing;
    int alt_row_spacing, alt_col_spacing;
    int alt_row_offset, alt_col_offset;
    int alt_col_flip_h, alt_col_flip_v, alt_row_flip_h, alt_row_flip_v;
} mo_shape;

/* GridUtility.cpp:25-52 */
void mo_grid_size(const mo_shape *s, int image_w, int image_h, int pad, int *gx, int *gy)
{
    if (s->col_spacing != s->alt_col_spacing)
        *gx = 2 * ((image_w + s->col_spacing + s->alt_col_spacing - 1) / (s->col_spacing + s->alt_col_spacing));
    else
        *gx = (image_w + s->col_spacing - 1) / s->col_spacing;
    if (s->row_spacing != s->alt_row_spacing)
        *gy = 2 * ((image_h + s->row_spacing + s->alt_row_spacing - 1) / (s->row_spacing + s->alt_row_spacing));
    else
        *gy = (image_h + s->row_spacing - 1) / s->row_spacing;
    *gx += pad;
    *gy += pad;
}

/* GridUtility.cpp:86-115. C's / and % truncate toward zero exactly like the reference's C++. */
void mo_rect_at(const mo_shape *s, int x, int y, int rect[4] /* x, y, w, h */)
{
    const int nx = x / 2, ax = x - nx;
    const int ny = y / 2, ay = y - ny;
    rect[0] = (x < 0) ? ax * s->col_spacing + nx * s->alt_col_spacing
                      : nx * s->col_spacing + ax * s->alt_col_spacing;
    if (y % 2 != 0)
        rect[0] += s->alt_row_offset;
    rect[1] = (y < 0) ? ay * s->row_spacing + ny * s->alt_row_spacing
                      : ny * s->row_spacing + ay * s->alt_row_spacing;
    if (x % 2 != 0)
        rect[1] += s->alt_col_offset;
    rect[2] = s->size;
    rect[3] = s->size;
}

/* GridUtility.cpp:118-132; returns flip_h + 2*flip_v (mask index order of
 * CUDAPhotomosaicGenerator.cpp:201 / CellGroup.cpp:146) */
int mo_flip_at(const mo_shape *s, int x, int y)
{
    int h = 0, v = 0;
    if (s->alt_col_flip_h && x % 2 != 0) h = !h;
    if (s->alt_row_flip_h && y % 2 != 0) h = !h;
    if (s->alt_col_flip_v && x % 2 != 0) v = !v;
    if (s->alt_row_flip_v && y % 2 != 0) v = !v;
    return h + 2 * v;
}

static int mo_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* PhotomosaicGeneratorBase.cpp:296-326: image-clamped global rect -> local bound ->
 * detail-space bound. global_clamped = {xStart, yStart, xEnd, yEnd}; local = {x,y,w,h};
 * dbound = {x,y,w,h} at detail resolution. */
void mo_cell_bounds(const mo_shape *normal, int detail_size, double detail, int x, int y,
                    int image_w, int image_h, int global_clamped[4], int local[4], int dbound[4])
{
    int r[4];
    mo_rect_at(normal, x, y, r);
    const int y0 = mo_clampi(r[1], 0, image_h), y1 = mo_clampi(r[1] + r[3], 0, image_h);
    const int x0 = mo_clampi(r[0], 0, image_w), x1 = mo_clampi(r[0] + r[2], 0, image_w);
    global_clamped[0] = x0; global_clamped[1] = y0; global_clamped[2] = x1; global_clamped[3] = y1;
    local[0] = x0 - r[0]; local[1] = y0 - r[1]; local[2] = x1 - x0; local[3] = y1 - y0;
    int bx = (int)(local[0] * detail), by = (int)(local[1] * detail);
    int bw = (int)(local[2] * detail), bh = (int)(local[3] * detail);
    if (bx > detail_size - 1) bx = detail_size - 1;
    if (by > detail_size - 1) by = detail_size - 1;
    if (bw < 1) bw = 1;
    if (bh < 1) bh = 1;
    if (bw > detail_size - bx) bw = detail_size - bx;
    if (bh > detail_size - by) bh = detail_size - by;
    dbound[0] = bx; dbound[1] = by; dbound[2] = bw; dbound[3] = bh;
}

/* ------------------------------------------------------------------ repeats */

/* CPUPhotomosaicGenerator.cpp:185-225. grid is rows x cols of int64 (-1 = nullopt),
 * (x, y) already padded coordinates. Writes (index, penalty) pairs; returns count.
 * ids/pen need room for min(range*(2*range+1)+range, rows*cols) entries. */
int mo_calculate_repeats(const int64_t *grid, int rows, int cols, int x, int y,
                         int range, int addition, int64_t *ids, int64_t *pen)
{
    int n = 0;
    const int y0 = mo_clampi(y - range, 0, rows);
    const int x0 = mo_clampi(x - range, 0, cols);
    const int x1 = mo_clampi(x + range, 0, cols - 1);
    for (int pass = 0; pass < 2; ++pass) {
        /* pass 0: rows above, full window width (:196-209); pass 1: same row, left only (:212-222) */
        const int ry0 = pass == 0 ? y0 : y, ry1 = pass == 0 ? y : y + 1;
        for (int ry = ry0; ry < ry1; ++ry) {
            const int rx1 = pass == 0 ? x1 + 1 : x;
            for (int rx = x0; rx < rx1; ++rx) {
                const int64_t v = grid[(int64_t)ry * cols + rx];
                if (v < 0)
                    continue;
                int k = 0;
                while (k < n && ids[k] != v)
                    ++k;
                if (k == n) {
                    ids[n] = v;
                    pen[n] = addition;
                    ++n;
                } else
                    pen[k] += addition;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------ best fit */

typedef struct {
    int64_t visited;  /* pixel-differences actually evaluated (early exit on) */
    int64_t nominal;  /* active in-bound pixels x N x V */
} mo_stats;

/* CPUPhotomosaicGenerator.cpp:116-181 for one cell.
 *   cell   : V images, dsize x dsize x 3 f32 (already at detail size, getCellAt)
 *   lib    : N images, dsize x dsize x 3 f32
 *   mask   : dsize x dsize u8 (already the flipped variant for this cell)
 *   bound  : detail-space {x, y, w, h}
 *   early_exit: 1 = reference behaviour (stop summing once >= best)
 *   row_out: optional N doubles: min over variants of the FULL masked sum
 *            (no early exit, no repeat penalty) -- the D matrix row used for parity.
 * Returns best index or -1. best_score / second_score receive the winning and
 * runner-up penalised scores when row_out is requested (full sums). */
int64_t mo_find_cell_best_fit(int type, const float *cell, int V, const float *lib, int64_t N,
                              int dsize, const uint8_t *mask, const int bound[4],
                              const int64_t *rep_ids, const int64_t *rep_pen, int n_rep,
                              int early_exit, double *row_out, double *best_score, double *second_score,
                              mo_stats *stats)
{
    const int64_t P3 = (int64_t)dsize * dsize * 3;
    int64_t best = -1;
    double best_variant = DBL_MAX;
    const int bx0 = bound[0], by0 = bound[1], bx1 = bound[0] + bound[2], by1 = bound[1] + bound[3];

    if (stats) {
        int64_t act = 0;
        for (int r = by0; r < by1; ++r)
            for (int c = bx0; c < bx1; ++c)
                act += mask[r * dsize + c] != 0;
        stats->nominal += act * N * V;
    }

    for (int64_t i = 0; early_exit && i < N; ++i) {
        const float *im = lib + i * P3;
        double rep = 0;
        for (int k = 0; k < n_rep; ++k)
            if (rep_ids[k] == i)
                rep = (double)rep_pen[k];
        for (int v = 0; v < V; ++v) {
            const float *cv = cell + v * P3;
            double variant = rep;
            for (int r = by0; r < by1 && variant < best_variant; ++r) {
                const float *pm = cv + (int64_t)r * dsize * 3;
                const float *pi = im + (int64_t)r * dsize * 3;
                const uint8_t *pk = mask + (int64_t)r * dsize;
                for (int c = bx0; c < bx1 && variant < best_variant; ++c) {
                    if (pk[c] != 0) {
                        variant += mo_diff_f32(type, pm + 3 * c, pi + 3 * c);
                        if (stats)
                            stats->visited++;
                    }
                }
            }
            if (variant < best_variant) {
                best_variant = variant;
                best = i;
            }
        }
    }

    if (row_out || !early_exit) {
        /* full sums, same pixel order, no early exit */
        double b1 = DBL_MAX, b2 = DBL_MAX;
        int64_t bi = -1;
        for (int64_t i = 0; i < N; ++i) {
            const float *im = lib + i * P3;
            double rep = 0;
            for (int k = 0; k < n_rep; ++k)
                if (rep_ids[k] == i)
                    rep = (double)rep_pen[k];
            double dmin = DBL_MAX;
            for (int v = 0; v < V; ++v) {
                const float *cv = cell + v * P3;
                double s = 0;
                for (int r = by0; r < by1; ++r)
                    for (int c = bx0; c < bx1; ++c)
                        if (mask[r * dsize + c] != 0)
                            s += mo_diff_f32(type, cv + ((int64_t)r * dsize + c) * 3, im + ((int64_t)r * dsize + c) * 3);
                if (s < dmin)
                    dmin = s;
            }
            if (row_out)
                row_out[i] = dmin;
            const double pen = rep + dmin;
            if (pen < b1) { b2 = b1; b1 = pen; bi = i; }
            else if (pen < b2) b2 = pen;
        }
        if (best_score) *best_score = b1;
        if (second_score) *second_score = b2;
        if (!early_exit)
            best = bi;
    }
    return best;
}

/* CPUPhotomosaicGenerator.cpp:49-100 for ONE size step, raster order.
 *   cells     : per VALID cell (raster order) V x dsize x dsize x 3 f32, packed
 *   bounds    : per valid cell {x,y,w,h}
 *   flips     : per valid cell mask index (h + 2v)
 *   masks     : 4 x dsize x dsize
 *   grid      : rows x cols int64, in: -1 nullopt / >=0 valid; out: best fits
 *   D_out     : optional n_valid x N doubles; margins optional n_valid x 2
 *   y_begin/y_end: padded row range to process (timed slices use a sub-range)
 */
int mo_generate_step(int type, const float *cells, const int *bounds, const int *flips, int V,
                     const float *lib, int64_t N, int dsize, const uint8_t *masks,
                     int64_t *grid, int rows, int cols, int range, int addition,
                     int early_exit, int y_begin, int y_end, double *D_out, double *margins, mo_stats *stats)
{
    const int64_t cell_stride = (int64_t)V * dsize * dsize * 3;
    /* distinct images in a window never exceed the number of grid cells (the UI allows ranges up to 100000) */
    const size_t max_rep = (size_t)rows * (size_t)cols + 1;
    int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * max_rep);
    int64_t *pen = (int64_t *)malloc(sizeof(int64_t) * max_rep);
    if (!ids || !pen)
        return -1;
    int64_t vi = 0; /* running index of valid cells */
    for (int y = 0; y < rows; ++y) {
        for (int x = 0; x < cols; ++x) {
            if (grid[(int64_t)y * cols + x] < 0)
                continue;
            const int64_t me = vi++;
            if (y < y_begin || y >= y_end)
                continue;
            const int n_rep = mo_calculate_repeats(grid, rows, cols, x, y, range, addition, ids, pen);
            double b1 = 0, b2 = 0;
            const int64_t fit = mo_find_cell_best_fit(
                type, cells + me * cell_stride, V, lib, N, dsize,
                masks + (int64_t)flips[me] * dsize * dsize, bounds + 4 * me, ids, pen, n_rep, early_exit,
                D_out ? D_out + me * N : NULL, &b1, &b2, stats);
            if (margins) {
                margins[2 * me] = b1;
                margins[2 * me + 1] = b2;
            }
            /* a -1 here would be the reference's "should never happen" nullopt (:174-178) */
            grid[(int64_t)y * cols + x] = fit;
        }
    }
    free(ids);
    free(pen);
    return 0;
}

/* Selection only, from a precomputed D matrix (f64): the reference's repeat rule and
 * strict-< lowest-index argmin applied in raster order. Used to test the GPU selection
 * stage in isolation (mirrors test/tst_CUDAKernel.h:20-165). */
void mo_select_from_D(const double *D, int64_t N, int64_t *grid, int rows, int cols, int range, int addition)
{
    const size_t max_rep = (size_t)rows * (size_t)cols + 1;
    int64_t *ids = (int64_t *)malloc(sizeof(int64_t) * max_rep);
    int64_t *pen = (int64_t *)malloc(sizeof(int64_t) * max_rep);
    int64_t vi = 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            if (grid[(int64_t)y * cols + x] < 0)
                continue;
            const double *row = D + (vi++) * N;
            const int n_rep = mo_calculate_repeats(grid, rows, cols, x, y, range, addition, ids, pen);
            double bestv = DBL_MAX;
            int64_t best = -1;
            for (int64_t i = 0; i < N; ++i) {
                double rep = 0;
                for (int k = 0; k < n_rep; ++k)
                    if (ids[k] == i)
                        rep = (double)pen[k];
                const double v = rep + row[i];
                if (v < bestv) { bestv = v; best = i; }
            }
            grid[(int64_t)y * cols + x] = best;
        }
    free(ids);
    free(pen);
}

/* ------------------------------------------------------------------ entropy */

/* ImageUtility.cpp:189-242 after BGR2GRAY: gray and mask are n bytes (mask may be NULL). */
double mo_entropy(const uint8_t *gray, const uint8_t *mask, int64_t n)
{
    if (n <= 0)
        return 0;
    size_t hist[256];
    size_t count = 0;
    memset(hist, 0, sizeof hist);
    for (int64_t i = 0; i < n; ++i)
        if (!mask || mask[i] != 0) {
            hist[gray[i]]++;
            count++;
        }
    double e = 0;
    for (int b = 0; b < 256; ++b) {
        const double p = hist[b] / (double)count;
        if (p > 0)
            e -= p * log2(p);
    }
    return e;
}
