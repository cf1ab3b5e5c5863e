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
}
