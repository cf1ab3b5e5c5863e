// Harness around the reference's OWN CPUPhotomosaicGenerator.cpp, which oracle/Makefile compiles UNMODIFIED from
// /root/reference into oracle/_ref/libref_core.so (stand-in headers: oracle/shim). Test infrastructure only.
//
// What runs from the reference's object code: CPUPhotomosaicGenerator::generateBestFits (the step / row / column loops,
// progress weights, library halving point), findCellBestFit (masked, bounded sum with the early exit, variant loop,
// strict-< argmin) and calculateRepeats -- CPUPhotomosaicGenerator.cpp:33-225 -- with ColourDifference.cpp and
// GridUtility.cpp, also unmodified.
// What the harness supplies instead of PhotomosaicGeneratorBase.cpp (which is OpenCV calls: cvtColor, resize): the
// preprocessed main images / libraries and the per-cell getCellAt results, computed by the oracle's cv2 path
// (oracle/oracle.py) and handed in as plain arrays. So this pins the oracle's restatement of the generator LOGIC on the
// reference itself; the OpenCV numerics stay pinned on cv2.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>

#include "CPUPhotomosaicGenerator.h"
#include "GridGenerator.h"

int g_ref_message_boxes = 0;

namespace {
struct CellEntry {
    std::vector<cv::Mat> cells;  // V variants, detail size
    cv::Rect bounds;
};
struct State {
    std::vector<std::vector<cv::Mat>> libs;                         // per step
    std::vector<cv::Mat> mains;                                     // V (only their count matters to the reference code)
    std::map<std::tuple<int, int, int>, CellEntry> cells;           // (cell size, x, y) -> getCellAt result
    size_t lib_step = 0;
    std::vector<int> progress;
} g;

cv::Mat mat_from(const void *src, int rows, int cols, size_t elem)
{
    cv::Mat m(rows, cols, elem);
    std::memcpy(m.data(), src, (size_t)rows * cols * elem);
    return m;
}
}  // namespace

// ---- the parts of PhotomosaicGeneratorBase the reference's CPU generator calls, supplied by the harness
PhotomosaicGeneratorBase::PhotomosaicGeneratorBase()
    : m_progress(0), m_wasCanceled(false), m_colourDiffType(ColourDifference::Type::RGB_EUCLIDEAN),
      m_colourSchemeType(ColourScheme::Type::NONE), m_repeatRange(0), m_repeatAddition(0)
{}
PhotomosaicGeneratorBase::~PhotomosaicGeneratorBase() {}
bool PhotomosaicGeneratorBase::generateBestFits() { return false; }
void PhotomosaicGeneratorBase::progress(const int t_progressStep) { g.progress.push_back(t_progressStep); }
std::vector<cv::Mat> PhotomosaicGeneratorBase::preprocessMainImage() { return g.mains; }
std::vector<cv::Mat> PhotomosaicGeneratorBase::preprocessLibraryImages() { return g.libs.at(0); }
std::pair<std::vector<cv::Mat>, cv::Rect> PhotomosaicGeneratorBase::getCellAt(const CellShape &t_cellShape, const CellShape &,
                                                                              const int x, const int y,
                                                                              const std::vector<cv::Mat> &) const
{
    const CellEntry &e = g.cells.at(std::make_tuple(t_cellShape.getSize(), x, y));
    return {e.cells, e.bounds};
}
// ---- GridGenerator.cpp's two OpenCV-backed helpers: remembered / forwarded to the cv2 callback
typedef double (*ref_entropy_fn)(const unsigned char *cell, int rows, int cols, long stride, int target_h, int target_w,
                                 const unsigned char *mask, int mask_rows, int mask_cols, long mask_stride);
static ref_entropy_fn g_entropy_cb = nullptr;
static int g_target_h = 0, g_target_w = 0;
cv::Mat ImageUtility::resizeImage(const cv::Mat &t_img, const int t_targetHeight, const int t_targetWidth, const ResizeType)
{
    g_target_h = t_targetHeight;  // the resize itself happens inside the callback (cv2), together with the entropy
    g_target_w = t_targetWidth;
    return t_img;
}
double ImageUtility::calculateEntropy(const cv::Mat &t_in, const cv::Mat &t_mask)
{
    if (t_in.empty())
        return 0;  // ImageUtility.cpp:191-192
    return g_entropy_cb(t_in.data(), t_in.rows, t_in.cols, (long)t_in.step(), g_target_h, g_target_w, t_mask.data(), t_mask.rows,
                        t_mask.cols, (long)t_mask.step());
}
bool ImageUtility::batchResizeMat(std::vector<cv::Mat> &t_images, const double)
{
    t_images = g.libs.at(++g.lib_step);  // the next step's library, halved by the oracle's cv2 path
    return true;
}

namespace {
struct Runner : public CPUPhotomosaicGenerator {
    void configure(int diff_type, int repeat_range, int repeat_addition, const CellGroup &cells,
                   const GridUtility::MosaicBestFit &state)
    {
        m_colourDiffType = static_cast<ColourDifference::Type>(diff_type);
        m_colourDiffFunc = ColourDifference::getFunction(m_colourDiffType);
        m_repeatRange = repeat_range;
        m_repeatAddition = repeat_addition;
        m_cells = cells;
        m_bestFits = state;
    }
    const GridUtility::MosaicBestFit &fits() const { return m_bestFits; }
};
}  // namespace

extern "C" {
// One whole generateBestFits() of the reference's CPU generator.
//   per step s: shapes[s] = 11 ints (size, rowSp, colSp, altRowSp, altColSp, altRowOff, altColOff, colFlipH, colFlipV, rowFlipH,
//   rowFlipV) of the NORMAL cell; ds[s] detail size; masks[s] = 4 x ds x ds u8 (index flip_h + 2 flip_v); libs[s] = N x ds x ds x 3
//   f32; grid_rows/cols[s]; grids[s] = rows x cols int64 (-1 nullopt, else valid) IN/OUT; n_cells[s]; cell_xy[s] = n x 2 (unpadded
//   x, y); cell_bounds[s] = n x 4 (x, y, w, h in detail space); cell_px[s] = n x V x ds x ds x 3 f32.
// progress_out (optional) receives up to progress_cap emitted progress values; returns their count, or -1 when the reference
// returned false, -2 on a missing cell (harness misuse).
int ref_cpu_generate(int n_steps, int diff_type, int repeat_range, int repeat_addition, int n_lib, int n_variants,
                     const int *const *shapes, const int *ds, const unsigned char *const *masks, const float *const *libs,
                     const int *grid_rows, const int *grid_cols, long long *const *grids, const int *n_cells,
                     const int *const *cell_xy, const int *const *cell_bounds, const float *const *cell_px, int *progress_out,
                     int progress_cap)
{
    g = State();
    g_ref_message_boxes = 0;
    CellGroup group;
    GridUtility::MosaicBestFit state;
    for (int s = 0; s < n_steps; ++s) {
        CellShape normal, detail;
        const int *p = shapes[s];
        normal.size = p[0]; normal.rowSpacing = p[1]; normal.colSpacing = p[2]; normal.altRowSpacing = p[3];
        normal.altColSpacing = p[4]; normal.altRowOffset = p[5]; normal.altColOffset = p[6];
        normal.colFlipH = p[7]; normal.colFlipV = p[8]; normal.rowFlipH = p[9]; normal.rowFlipV = p[10];
        detail = normal;
        detail.size = ds[s];
        for (int f = 0; f < 4; ++f)
            detail.masks[f] = mat_from(masks[s] + (size_t)f * ds[s] * ds[s], ds[s], ds[s], 1);
        group.cells.push_back(normal);
        group.detailCells.push_back(detail);

        std::vector<cv::Mat> lib;
        for (int i = 0; i < n_lib; ++i)
            lib.push_back(mat_from(libs[s] + (size_t)i * ds[s] * ds[s] * 3, ds[s], ds[s], sizeof(cv::Vec3f)));
        g.libs.push_back(lib);

        GridUtility::StepBestFit step(grid_rows[s], std::vector<GridUtility::CellBestFit>(grid_cols[s]));
        for (int y = 0; y < grid_rows[s]; ++y)
            for (int x = 0; x < grid_cols[s]; ++x)
                if (grids[s][(size_t)y * grid_cols[s] + x] >= 0)
                    step[y][x] = 0;
        state.push_back(step);

        const size_t cell_elems = (size_t)ds[s] * ds[s] * 3;
        for (int c = 0; c < n_cells[s]; ++c) {
            CellEntry e;
            for (int v = 0; v < n_variants; ++v)
                e.cells.push_back(mat_from(cell_px[s] + ((size_t)c * n_variants + v) * cell_elems, ds[s], ds[s], sizeof(cv::Vec3f)));
            const int *b = cell_bounds[s] + 4 * c;
            e.bounds = cv::Rect(b[0], b[1], b[2], b[3]);
            g.cells[std::make_tuple(normal.size, cell_xy[s][2 * c], cell_xy[s][2 * c + 1])] = e;
        }
    }
    g.mains.assign(n_variants, cv::Mat());

    Runner r;
    r.configure(diff_type, repeat_range, repeat_addition, group, state);
    bool ok;
    try {
        ok = r.generateBestFits();
    } catch (const std::out_of_range &) {
        return -2;
    }
    if (!ok)
        return -1;
    for (int s = 0; s < n_steps; ++s)
        for (int y = 0; y < grid_rows[s]; ++y)
            for (int x = 0; x < grid_cols[s]; ++x) {
                const auto &v = r.fits()[s][y][x];
                grids[s][(size_t)y * grid_cols[s] + x] = v.has_value() ? (long long)v.value() : -1;
            }
    const int n = (int)g.progress.size();
    for (int i = 0; i < n && i < progress_cap; ++i)
        progress_out[i] = g.progress[i];
    return n;
}
int ref_cpu_message_boxes(void) { return g_ref_message_boxes; }
// GridGenerator::getGridState (GridGenerator.cpp:29-110) from the reference's object code.
//   shapes[s] (11 ints, normal cell of step s), ds[s] + masks[s] (4 x ds x ds u8) for n_steps = size_steps + 1 levels,
//   detail in (0, 1]; bgr may be NULL (then height / width give the grid area); entropy = cv2-backed callback.
//   out: steps written back to back (-1 nullopt, 0 valid), step_rows / step_cols per generated step.
//   Returns the number of generated steps, or -1 when out_capacity is too small.
int ref_grid_state(int n_steps, const int *const *shapes, const int *ds, const unsigned char *const *masks, double detail,
                   const unsigned char *bgr, int rows, int cols, long stride, int height, int width, ref_entropy_fn entropy,
                   long long *out, long long out_capacity, int *step_rows, int *step_cols)
{
    g_entropy_cb = entropy;
    CellGroup group;
    group.detail = detail;
    for (int s = 0; s < n_steps; ++s) {
        CellShape normal, dcell;
        const int *p = shapes[s];
        normal.size = p[0]; normal.rowSpacing = p[1]; normal.colSpacing = p[2]; normal.altRowSpacing = p[3];
        normal.altColSpacing = p[4]; normal.altRowOffset = p[5]; normal.altColOffset = p[6];
        normal.colFlipH = p[7]; normal.colFlipV = p[8]; normal.rowFlipH = p[9]; normal.rowFlipV = p[10];
        normal.masks[0] = cv::Mat(1, 1, 1);  // getGridState only asks whether the top mask is empty (:35)
        dcell = normal;
        dcell.size = ds[s];
        for (int f = 0; f < 4; ++f)
            dcell.masks[f] = mat_from(masks[s] + (size_t)f * ds[s] * ds[s], ds[s], ds[s], 1);
        group.cells.push_back(normal);
        group.detailCells.push_back(dcell);
    }
    cv::Mat image;
    if (bgr) {
        image = cv::Mat(rows, cols, 3);
        for (int y = 0; y < rows; ++y)
            std::memcpy(image.ptr<unsigned char>(y), bgr + (size_t)y * stride, (size_t)cols * 3);
    }
    const GridUtility::MosaicBestFit state = GridGenerator::getGridState(group, image, height, width);
    long long used = 0;
    for (size_t s = 0; s < state.size(); ++s) {
        const int r = (int)state[s].size(), c = r ? (int)state[s][0].size() : 0;
        if (used + (long long)r * c > out_capacity)
            return -1;
        step_rows[s] = r;
        step_cols[s] = c;
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < c; ++x)
                out[used++] = state[s][y][x].has_value() ? (long long)state[s][y][x].value() : -1;
    }
    return (int)state.size();
}
}
