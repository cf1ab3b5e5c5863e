// Harness around the reference's OWN generator sources, which oracle/Makefile compiles UNMODIFIED from /root/reference
// into oracle/_ref/libref_core.so against the stand-in headers of oracle/shim. Test infrastructure only.
//
// Reference object code in the library:
//   PhotomosaicGeneratorBase.cpp   setters, preprocessMainImage / preprocessLibraryImages, getCellAt, buildPhotomosaic,
//                                  getMaxProgress (PhotomosaicGeneratorBase.cpp:33-329)
//   CPUPhotomosaicGenerator.cpp    generateBestFits, findCellBestFit, calculateRepeats (:33-225)
//   GridGenerator.cpp              getGridState, findCellState (:29-193)
//   ColourDifference.cpp, GridUtility.cpp, GridBounds.cpp
// What is NOT reference code here, and why:
//   * cv::cvtColor / cv::resize are OpenCV arithmetic: forwarded through a callback to the real OpenCV (cv2, oracle.py);
//   * ImageUtility.cpp (Qt GUI types, CUDA warping) and ColourScheme.cpp (cv::Mat_ / forEach templates) cannot be compiled
//     against the stand-ins: the five ImageUtility functions the generator calls are restated below line by line, and the
//     colour-scheme variants come from the oracle's cv2 restatement through a second callback;
//   * CellShape / CellGroup are plain holders of the masks and tiling parameters the oracle derives (their resizing is
//     cv::resize + threshold, pinned separately against cv2).
#include <cstdint>
#include <cstring>

#include "CPUPhotomosaicGenerator.h"
#include "GridGenerator.h"

int g_ref_message_boxes = 0;

// ---- callbacks into the real OpenCV (set per call)
//   op 0: cvtColor(code); op 1: resize(interpolation = code). dst is allocated by the caller side here.
typedef int (*ref_cv_fn)(int op, int code, const unsigned char *src, int rows, int cols, int type, long step, unsigned char *dst,
                         int drows, int dcols, int dtype, long dstep);
//   colour-scheme variants 1 .. V-1 of an 8U BGR image (variant 0 is the image itself), written back to back into dst
typedef int (*ref_scheme_fn)(int scheme, const unsigned char *src, int rows, int cols, long step, unsigned char *dst);
static ref_cv_fn g_cv = nullptr;
static ref_scheme_fn g_scheme = nullptr;
static std::vector<int> g_progress;

static void call_cv(int op, int code, const cv::Mat &src, cv::Mat &dst, int drows, int dcols, int dtype)
{
    cv::Mat out(drows, dcols, dtype);  // fresh storage: src and dst may be the same Mat object
    if (!g_cv || g_cv(op, code, src.data, src.rows, src.cols, src.type(), (long)(size_t)src.step, out.data, drows, dcols, dtype,
                      (long)(size_t)out.step) != 0)
        throw std::runtime_error("OpenCV callback failed");
    dst = out;
}
void cv::cvtColor(const cv::Mat &src, cv::Mat &dst, int code)
{
    int dtype = src.type();
    if (code == cv::COLOR_BGR2GRAY)
        dtype = CV_MAKETYPE(src.depth(), 1);
    else if (code == cv::COLOR_BGR2BGRA)
        dtype = CV_MAKETYPE(src.depth(), 4);
    call_cv(0, code, src, dst, src.rows, src.cols, dtype);
}
void cv::resize(const cv::Mat &src, cv::Mat &dst, cv::Size dsize, double, double, int interpolation)
{
    call_cv(1, interpolation, src, dst, dsize.height, dsize.width, src.type());
}

// ---- ImageUtility, restated (the reference file cannot be compiled here, see the header comment)
// ImageUtility.cpp:34-62
cv::Mat ImageUtility::resizeImage(const cv::Mat &t_img, const int t_targetHeight, const int t_targetWidth, const ResizeType t_type)
{
    double resizeFactor = static_cast<double>(t_targetHeight) / t_img.rows;
    if ((t_type == ResizeType::EXCLUSIVE && t_targetWidth < resizeFactor * t_img.cols) ||
        (t_type == ResizeType::INCLUSIVE && t_targetWidth > resizeFactor * t_img.cols) ||
        (t_type == ResizeType::EXACT && resizeFactor == 1.0))
        resizeFactor = static_cast<double>(t_targetWidth) / t_img.cols;
    if (resizeFactor == 1.0)
        return t_img;  // the SAME Mat: with detail 100 % every colour-scheme variant of a cell keeps aliasing one buffer (Q1)
    const cv::InterpolationFlags flags = (resizeFactor < 1) ? cv::INTER_AREA : cv::INTER_CUBIC;
    cv::Mat result;
    if (t_type == ResizeType::EXACT)
        cv::resize(t_img, result, cv::Size(t_targetWidth, t_targetHeight), 0, 0, flags);
    else
        cv::resize(t_img, result, cv::Size((int)std::round(resizeFactor * t_img.cols), (int)std::round(resizeFactor * t_img.rows)), 0, 0,
                   flags);
    return result;
}
// ImageUtility.cpp:66-101
bool ImageUtility::batchResizeMat(std::vector<cv::Mat> &t_images, const double t_ratio)
{
    if (t_images.empty())
        return false;
    const int h = (int)std::round(t_ratio * t_images.front().rows), w = (int)std::round(t_ratio * t_images.front().cols);
    for (auto &im : t_images)
        im = resizeImage(im, h, w, ResizeType::EXACT);
    return true;
}
// ImageUtility.cpp:173-186 (split + constant 255 plane + merge = BGR -> BGRA)
void ImageUtility::addAlphaChannel(std::vector<cv::Mat> &t_images)
{
    for (auto &image : t_images)
        cv::cvtColor(image, image, cv::COLOR_BGR2BGRA);
}
// ImageUtility.cpp:189-242
double ImageUtility::calculateEntropy(const cv::Mat &t_in, const cv::Mat &t_mask)
{
    if (t_in.empty())
        return 0;
    if (!t_mask.empty() && (t_mask.rows != t_in.rows || t_mask.cols != t_in.cols || t_mask.channels() != 1))
        return 0;
    cv::Mat grayImage;
    cv::cvtColor(t_in, grayImage, cv::COLOR_BGR2GRAY);
    size_t pixelCount = 0;
    std::vector<size_t> histogram(256, 0);
    for (int row = 0; row < grayImage.rows; ++row) {
        const uchar *p_im = grayImage.ptr<uchar>(row);
        const uchar *p_mask = t_mask.empty() ? nullptr : t_mask.ptr<uchar>(row);
        for (int col = 0; col < grayImage.cols; ++col)
            if (!p_mask || p_mask[col] != 0) {
                ++histogram.at(p_im[col]);
                ++pixelCount;
            }
    }
    double entropy = 0;
    for (auto value : histogram) {
        const double probability = value / static_cast<double>(pixelCount);
        if (probability > 0)
            entropy -= probability * std::log2(probability);
    }
    return entropy;
}

// ---- ColourScheme::getFunction (ColourScheme.cpp:20-33): variants from the oracle's cv2 restatement
ColourScheme::FunctionType ColourScheme::getFunction(const Type &t_type)
{
    const int scheme = static_cast<int>(t_type);
    static const int kVariants[] = {1, 2, 3, 3, 4, 4};  // NONE, COMPLEMENTARY, TRIADIC, COMPOUND, TETRADIC, ANALAGOUS
    if (scheme < 0 || scheme > 5)
        throw std::invalid_argument("No function for given type");
    return [scheme](const cv::Mat &t_image) {
        std::vector<cv::Mat> out{t_image};  // the original image is always the first variant
        const int extra = kVariants[scheme] - 1;
        if (extra > 0) {
            cv::Mat all(t_image.rows * extra, t_image.cols, t_image.type());
            if (!g_scheme || g_scheme(scheme, t_image.data, t_image.rows, t_image.cols, (long)(size_t)t_image.step, all.data) != 0)
                throw std::runtime_error("colour-scheme callback failed");
            for (int v = 0; v < extra; ++v)
                out.push_back(cv::Mat(all, cv::Range(v * t_image.rows, (v + 1) * t_image.rows), cv::Range(0, t_image.cols)).clone());
        }
        return out;
    };
}

// ---- the moc-generated signal body
void PhotomosaicGeneratorBase::progress(const int t_progressStep) { g_progress.push_back(t_progressStep); }

namespace {
cv::Mat mat_from(const void *src, int rows, int cols, int type, size_t src_step = 0)
{
    cv::Mat m(rows, cols, type);
    const size_t row_bytes = (size_t)cols * m.elemSize();
    for (int y = 0; y < rows; ++y)
        std::memcpy(m.ptr<unsigned char>(y), (const unsigned char *)src + (size_t)y * (src_step ? src_step : row_bytes), row_bytes);
    return m;
}

// shapes[s]: 11 ints of the NORMAL cell of step s; masks4[s]: 4 x S x S u8 (index flip_h + 2 flip_v) of the normal cell;
// dmasks4[s]: 4 x ds x ds of the detail cell
CellGroup make_group(int n_steps, const int *const *shapes, const unsigned char *const *masks4, const int *ds,
                     const unsigned char *const *dmasks4, double detail)
{
    CellGroup group;
    group.detail = detail;
    for (int s = 0; s < n_steps; ++s) {
        CellShape normal;
        const int *p = shapes[s];
        normal.size = p[0]; normal.rowSpacing = p[1]; normal.colSpacing = p[2]; normal.altRowSpacing = p[3];
        normal.altColSpacing = p[4]; normal.altRowOffset = p[5]; normal.altColOffset = p[6];
        normal.colFlipH = p[7]; normal.colFlipV = p[8]; normal.rowFlipH = p[9]; normal.rowFlipV = p[10];
        CellShape dcell = normal;
        dcell.size = ds[s];
        for (int f = 0; f < 4; ++f) {
            normal.masks[f] = mat_from(masks4[s] + (size_t)f * p[0] * p[0], p[0], p[0], CV_8UC1);
            dcell.masks[f] = mat_from(dmasks4[s] + (size_t)f * ds[s] * ds[s], ds[s], ds[s], CV_8UC1);
        }
        group.cells.push_back(normal);
        group.detailCells.push_back(dcell);
    }
    return group;
}

struct Runner : public CPUPhotomosaicGenerator {
    using PhotomosaicGeneratorBase::getCellAt;
    using PhotomosaicGeneratorBase::preprocessMainImage;
};

struct Session {
    Runner gen;
    CellGroup group;
    std::vector<cv::Mat> mains;  // preprocessMainImage() of the current inputs (for ref_get_cell_at)
};
}  // namespace

extern "C" {
void ref_set_callbacks(ref_cv_fn cv_cb, ref_scheme_fn scheme_cb)
{
    g_cv = cv_cb;
    g_scheme = scheme_cb;
}

// A generator object configured like MainWindow.cpp:584-607 / tst_Generator.h:111-136 does it.
void *ref_session_create(const unsigned char *bgr, int rows, int cols, long stride, const unsigned char *lib, int n_lib, int lib_size,
                         int n_steps, const int *const *shapes, const unsigned char *const *masks4, const int *ds,
                         const unsigned char *const *dmasks4, double detail, int diff_type, int scheme, int repeat_range,
                         int repeat_addition)
{
    try {
        Session *s = new Session;
        s->group = make_group(n_steps, shapes, masks4, ds, dmasks4, detail);
        s->gen.setMainImage(mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride));
        std::vector<cv::Mat> library;
        for (int i = 0; i < n_lib; ++i)
            library.push_back(mat_from(lib + (size_t)i * lib_size * lib_size * 3, lib_size, lib_size, CV_8UC3));
        s->gen.setLibrary(library);
        s->gen.setColourDifference(static_cast<ColourDifference::Type>(diff_type));
        s->gen.setColourScheme(static_cast<ColourScheme::Type>(scheme));
        s->gen.setCellGroup(s->group);
        s->gen.setRepeat(repeat_range, repeat_addition);
        return s;
    } catch (const std::exception &) {
        return nullptr;
    }
}
void ref_session_destroy(void *h) { delete static_cast<Session *>(h); }

// setGridState + generateBestFits + getBestFits. grids[s]: rows x cols int64, in: -1 nullopt / >= 0 valid, out: best fits.
// Returns the number of progress emissions (values into progress_out up to progress_cap), -1 if the reference returned false,
// -3 on an exception.
int ref_session_generate(void *h, int n_steps, const int *grid_rows, const int *grid_cols, long long *const *grids, int *progress_out,
                         int progress_cap, int *max_progress)
{
    Session *s = static_cast<Session *>(h);
    try {
        GridUtility::MosaicBestFit state;
        for (int st = 0; st < n_steps; ++st) {
            GridUtility::StepBestFit step(grid_rows[st], std::vector<GridUtility::CellBestFit>(grid_cols[st]));
            for (int y = 0; y < grid_rows[st]; ++y)
                for (int x = 0; x < grid_cols[st]; ++x)
                    if (grids[st][(size_t)y * grid_cols[st] + x] >= 0)
                        step[y][x] = 0;
            state.push_back(step);
        }
        s->gen.setGridState(state);
        if (max_progress)
            *max_progress = s->gen.getMaxProgress();
        g_progress.clear();
        g_ref_message_boxes = 0;
        if (!s->gen.generateBestFits())
            return -1;
        const GridUtility::MosaicBestFit fits = s->gen.getBestFits();
        for (int st = 0; st < n_steps; ++st)
            for (int y = 0; y < grid_rows[st]; ++y)
                for (int x = 0; x < grid_cols[st]; ++x) {
                    const auto &v = fits[st][y][x];
                    grids[st][(size_t)y * grid_cols[st] + x] = v.has_value() ? (long long)v.value() : -1;
                }
        const int n = (int)g_progress.size();
        for (int i = 0; i < n && i < progress_cap; ++i)
            progress_out[i] = g_progress[i];
        return n;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_message_boxes(void) { return g_ref_message_boxes; }

// buildPhotomosaic(background) on the best fits of the last generate: out is rows x cols x 4 (BGRA), contiguous
int ref_session_build(void *h, const double background[4], unsigned char *out)
{
    Session *s = static_cast<Session *>(h);
    try {
        const cv::Mat m = s->gen.buildPhotomosaic(cv::Scalar(background[0], background[1], background[2], background[3]));
        for (int y = 0; y < m.rows; ++y)
            std::memcpy(out + (size_t)y * m.cols * 4, m.ptr<unsigned char>(y), (size_t)m.cols * 4);
        return 0;
    } catch (const std::exception &) {
        return -3;
    }
}

// getCellAt(cell shape of `step`, detail cell shape, x, y, preprocessMainImage()): cells_out = V x ds x ds x 3 f32,
// bounds_out = x, y, w, h in detail space. Returns V.
int ref_session_get_cell_at(void *h, int step, int x, int y, float *cells_out, int *bounds_out)
{
    Session *s = static_cast<Session *>(h);
    try {
        if (s->mains.empty())
            s->mains = s->gen.preprocessMainImage();
        const auto r = s->gen.getCellAt(s->group.getCell(step), s->group.getCell(step, true), x, y, s->mains);
        const int ds = s->group.getCellSize(step, true);
        for (size_t v = 0; v < r.first.size(); ++v)
            for (int row = 0; row < ds; ++row)
                std::memcpy(cells_out + (v * ds + row) * (size_t)ds * 3, r.first[v].ptr<float>(row), (size_t)ds * 3 * sizeof(float));
        bounds_out[0] = r.second.x; bounds_out[1] = r.second.y; bounds_out[2] = r.second.width; bounds_out[3] = r.second.height;
        return (int)r.first.size();
    } catch (const std::exception &) {
        return -3;
    }
}

// GridGenerator::getGridState (GridGenerator.cpp:29-110). bgr may be NULL (then height / width give the grid area).
// out: steps written back to back (-1 nullopt, 0 valid). Returns the number of generated steps, -1 when out is too small.
int ref_grid_state(int n_steps, const int *const *shapes, const unsigned char *const *masks4, const int *ds,
                   const unsigned char *const *dmasks4, double detail, const unsigned char *bgr, int rows, int cols, long stride,
                   int height, int width, long long *out, long long out_capacity, int *step_rows, int *step_cols)
{
    try {
        const CellGroup group = make_group(n_steps, shapes, masks4, ds, dmasks4, detail);
        cv::Mat image;
        if (bgr)
            image = mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride);
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
    } catch (const std::exception &) {
        return -3;
    }
}
}
