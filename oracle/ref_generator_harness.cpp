// Harness around the reference's OWN generator sources, which oracle/Makefile compiles UNMODIFIED from /root/reference
// into oracle/_ref/libref_core.so against the stand-in headers of oracle/shim. Test infrastructure only.
//
// Reference object code in the library:
//   PhotomosaicGeneratorBase.cpp   setters, preprocessMainImage / preprocessLibraryImages, getCellAt, buildPhotomosaic,
//                                  getMaxProgress (PhotomosaicGeneratorBase.cpp:33-329)
//   CPUPhotomosaicGenerator.cpp    generateBestFits, findCellBestFit, calculateRepeats (:33-225)
//   GridGenerator.cpp              getGridState, findCellState (:29-193)
//   ImageUtility.cpp               resizeImage, batchResizeMat, imageToSquare, addAlphaChannel, calculateEntropy, edgeDetect ...
//   ColourScheme.cpp               the colour-scheme variants (hue rotations in float HSV_FULL, :20-177)
//   ColourDifference.cpp, GridUtility.cpp, GridBounds.cpp
// What is NOT reference code here, and why:
//   * cv::cvtColor / cv::resize are OpenCV arithmetic: forwarded through a callback to the real OpenCV (cv2, oracle.py);
//   * MatUtility.h's parallel visitors (built on OpenCV internals: ParallelLoopBody, n-dimensional indexing) are plain loops in
//     the stand-in of the same name -- the functors they run are ColourScheme.cpp's own;
//   * the 8U helpers of the stand-in cv namespace (threshold, flip, split / merge, copyMakeBorder, bitwise ops) are written
//     out in oracle/shim/opencv2/core.hpp; the GUI-only ones (filter2D / dilate for the edge overlay) are simple versions, the
//     feature / face detectors refuse to run;
//   * cv::imencode / cv::imdecode (the PNG payload of .mcs / .mil) are the real OpenCV codec through a third callback.
// Also reference object code in the library: CellShape.cpp, CellGroup.cpp (per-step cell derivation, .mcs load / save through
// the reference's CustomQDataStream.h) and ImageLibrary.cpp (addImage, setImageSize, .mil load / save), on the Qt stand-ins of
// oracle/shim/qt_standins.h.
#include <cstdint>
#include <cstring>

#include "CPUPhotomosaicGenerator.h"
#include "GridGenerator.h"
#include "ImageLibrary.h"
#include "ref_group.h"

int g_ref_message_boxes = 0;

// ---- callbacks into the real OpenCV (set per call)
//   op 0: cvtColor(code); op 1: resize(interpolation = code). dst is allocated by the caller side here.
typedef int (*ref_cv_fn)(int op, int code, const unsigned char *src, int rows, int cols, int type, long step, unsigned char *dst,
                         int drows, int dcols, int dtype, long dstep);
static ref_cv_fn g_cv = nullptr;
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
    else if (code == cv::COLOR_BGR2BGRA || code == cv::COLOR_GRAY2RGBA)
        dtype = CV_MAKETYPE(src.depth(), 4);
    call_cv(0, code, src, dst, src.rows, src.cols, dtype);
}
void cv::resize(const cv::Mat &src, cv::Mat &dst, cv::Size dsize, double, double, int interpolation)
{
    call_cv(1, interpolation, src, dst, dsize.height, dsize.width, src.type());
}

// ---- PNG codec (imgcodecs) through the real OpenCV
//   op 0: encode `src` (rows x cols, type, step) as PNG, returns the byte count; op 1: copy the pending bytes to dst;
//   op 2: decode the n bytes at src, info = {rows, cols, type}; op 3: copy the pending pixels to dst (contiguous)
typedef long (*ref_codec_fn)(int op, const unsigned char *src, long n, int rows, int cols, int type, long step, unsigned char *dst,
                             int *info);
static ref_codec_fn g_codec = nullptr;
bool cv::imencode(const std::string &, const cv::Mat &img, std::vector<uchar> &buf)
{
    if (!g_codec)
        throw std::runtime_error("no codec callback");
    const long n = g_codec(0, img.data, 0, img.rows, img.cols, img.type(), (long)(size_t)img.step, nullptr, nullptr);
    if (n < 0)
        return false;
    buf.resize((size_t)n);
    return g_codec(1, nullptr, n, 0, 0, 0, 0, buf.data(), nullptr) == n;
}
cv::Mat cv::imdecode(const std::vector<uchar> &buf, int)
{
    if (!g_codec)
        throw std::runtime_error("no codec callback");
    int info[3] = {0, 0, 0};
    if (g_codec(2, buf.data(), (long)buf.size(), 0, 0, 0, 0, nullptr, info) < 0)
        return cv::Mat();
    cv::Mat m(info[0], info[1], info[2]);
    g_codec(3, nullptr, 0, info[0], info[1], info[2], (long)(size_t)m.step, m.data, nullptr);
    return m;
}

// ---- the moc-generated signal body. A QProgressDialog connected to the signal calls the generator's cancel() slot from inside
// the emission when the user presses Cancel (MainWindow.cpp:597-603); ref_cancel_after(n) plays that user at the n-th emission.
static int g_cancel_after = 0;
void PhotomosaicGeneratorBase::progress(const int t_progressStep)
{
    g_progress.push_back(t_progressStep);
    if (g_cancel_after > 0 && (int)g_progress.size() == g_cancel_after)
        cancel();
}
extern "C" {
void ref_cancel_after(int n_emissions) { g_cancel_after = n_emissions; }
void ref_progress_clear(void) { g_progress.clear(); }
int ref_progress_get(int *out, int cap)
{
    const int n = (int)g_progress.size();
    for (int i = 0; i < n && i < cap; ++i)
        out[i] = g_progress[i];
    return n;
}
}

namespace {
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
void ref_set_callbacks(ref_cv_fn cv_cb, ref_codec_fn codec_cb)
{
    g_cv = cv_cb;
    g_codec = codec_cb;
}

// ColourScheme::getFunction(type)(image) (ColourScheme.cpp:20-177): the V variants of an 8U BGR image, written back to back into
// out (capacity in images). Returns V, -3 on an exception.
int ref_colour_scheme_variants(int scheme, const unsigned char *bgr, int rows, int cols, long stride, unsigned char *out, int capacity)
{
    try {
        const std::vector<cv::Mat> v =
            ColourScheme::getFunction(static_cast<ColourScheme::Type>(scheme))(ref_mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride));
        for (size_t i = 0; i < v.size() && (int)i < capacity; ++i)
            for (int y = 0; y < rows; ++y)
                std::memcpy(out + (i * rows + y) * (size_t)cols * 3, v[i].ptr<unsigned char>(y), (size_t)cols * 3);
        return (int)v.size();
    } catch (const std::exception &) {
        return -3;
    }
}

// A generator object configured like MainWindow.cpp:584-607 / tst_Generator.h:111-136 does it.
void *ref_session_create(const unsigned char *bgr, int rows, int cols, long stride, const unsigned char *lib, int n_lib, int lib_size,
                         const int *shape, const unsigned char *mask, int detail_percent, int size_steps, int diff_type, int scheme,
                         int repeat_range, int repeat_addition)
{
    try {
        Session *s = new Session;
        s->group = ref_make_group(shape, mask, detail_percent, size_steps);
        s->gen.setMainImage(ref_mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride));
        std::vector<cv::Mat> library;
        for (int i = 0; i < n_lib; ++i)
            library.push_back(ref_mat_from(lib + (size_t)i * lib_size * lib_size * 3, lib_size, lib_size, CV_8UC3));
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
int ref_grid_state(const int *shape, const unsigned char *mask, int detail_percent, int size_steps, const unsigned char *bgr, int rows,
                   int cols, long stride, int height, int width, long long *out, long long out_capacity, int *step_rows, int *step_cols)
{
    try {
        const CellGroup group = ref_make_group(shape, mask, detail_percent, size_steps);
        cv::Mat image;
        if (bgr)
            image = ref_mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride);
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

// ---- CellShape / CellGroup (src/CellShape/CellShape.cpp, CellGroup.cpp)
static void describe(const CellShape &c, int *params, unsigned char *mask4, long mask_capacity)
{
    params[0] = c.getSize(); params[1] = c.getRowSpacing(); params[2] = c.getColSpacing();
    params[3] = c.getAlternateRowSpacing(); params[4] = c.getAlternateColSpacing();
    params[5] = c.getAlternateRowOffset(); params[6] = c.getAlternateColOffset();
    params[7] = c.getAlternateColFlipHorizontal(); params[8] = c.getAlternateColFlipVertical();
    params[9] = c.getAlternateRowFlipHorizontal(); params[10] = c.getAlternateRowFlipVertical();
    const long n = (long)c.getSize() * c.getSize();
    for (int f = 0; f < 4 && mask4 && (f + 1) * n <= mask_capacity; ++f) {
        const cv::Mat &m = c.getCellMask(f & 1, (f >> 1) & 1);  // index flip_h + 2 flip_v
        for (int y = 0; y < m.rows; ++y)
            std::memcpy(mask4 + f * n + (long)y * m.cols, m.ptr<unsigned char>(y), (size_t)m.cols);
    }
}
// the normal (detail = 0) or detail (detail = 1) cell of size step `step` of the group the application would build:
// params = 11 ints, mask4 = 4 x size x size (may be NULL). Returns the cell size, -3 on an exception.
int ref_cell_group_cell(const int *shape, const unsigned char *mask, int detail_percent, int size_steps, int step, int detail,
                        int *params, unsigned char *mask4, long mask_capacity)
{
    try {
        const CellGroup group = ref_make_group(shape, mask, detail_percent, size_steps);
        describe(group.getCell((size_t)step, detail != 0), params, mask4, mask_capacity);
        return params[0];
    } catch (const std::exception &) {
        return -3;
    }
}
// CellShape::resized (CellShape.cpp:281-312)
int ref_cell_shape_resized(const int *shape, const unsigned char *mask, int new_size, int *params, unsigned char *mask4, long mask_capacity)
{
    try {
        describe(ref_make_shape(shape, mask).resized(new_size), params, mask4, mask_capacity);
        return params[0];
    } catch (const std::exception &) {
        return -3;
    }
}
// CellShape::loadFromFile (CellShape.cpp:363-434): params + the four masks of the stored shape; name_utf8 receives the name.
// Returns the cell size, -1 when the reference throws std::invalid_argument, -3 on any other exception.
int ref_mcs_load(const char *path, int *params, unsigned char *mask4, long mask_capacity, char *name_utf8, int name_capacity)
{
    try {
        CellShape c;
        c.loadFromFile(QString(path));
        describe(c, params, mask4, mask_capacity);
        if (name_utf8 && name_capacity > 0) {
            std::strncpy(name_utf8, c.getName().toStdString().c_str(), (size_t)name_capacity - 1);
            name_utf8[name_capacity - 1] = 0;
        }
        return params[0];
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}
// What the application does with a shape file (MainWindow: loadFromFile, then the cell-size spin box -> resized, then the CellGroup):
// the normal / detail cell of one size step. cell_size <= 0 keeps the stored size. Returns the cell size, -1 / -3 as ref_mcs_load.
int ref_mcs_group_cell(const char *path, int cell_size, int detail_percent, int size_steps, int step, int detail, int *params,
                       unsigned char *mask4, long mask_capacity)
{
    try {
        CellShape c;
        c.loadFromFile(QString(path));
        CellGroup g;
        g.setCellShape(cell_size > 0 ? c.resized(cell_size) : c);
        g.setDetail(detail_percent);
        g.setSizeSteps(static_cast<size_t>(size_steps));
        describe(g.getCell((size_t)step, detail != 0), params, mask4, mask_capacity);
        return params[0];
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}
// CellShape::saveToFile (CellShape.cpp:321-360)
int ref_mcs_save(const char *path, const int *shape, const unsigned char *mask, const char *name_utf8)
{
    try {
        CellShape c = ref_make_shape(shape, mask);
        c.setName(QString::fromStdString(name_utf8 ? name_utf8 : ""));
        c.saveToFile(QString(path));
        return 0;
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}

// ---- ImageLibrary (src/ImageLibrary/ImageLibrary.cpp)
void *ref_library_create(int image_size) { return new ImageLibrary((size_t)image_size); }
void ref_library_destroy(void *h) { delete static_cast<ImageLibrary *>(h); }
// addImage (ImageLibrary.cpp:62-86): returns the (random) index the image was inserted at, -1 for an empty image
long ref_library_add(void *h, const unsigned char *bgr, int rows, int cols, long stride, const char *name_utf8)
{
    try {
        cv::Mat im;
        if (bgr && rows > 0 && cols > 0)
            im = ref_mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride);
        return (long)static_cast<ImageLibrary *>(h)->addImage(im, QString::fromStdString(name_utf8 ? name_utf8 : ""));
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_library_set_image_size(void *h, int size)
{
    try {
        static_cast<ImageLibrary *>(h)->setImageSize((size_t)size);
        return 0;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_library_image_size(void *h) { return (int)static_cast<ImageLibrary *>(h)->getImageSize(); }
long ref_library_count(void *h) { return (long)static_cast<ImageLibrary *>(h)->getImages().size(); }
// image i (size x size x 3, contiguous) and its name
int ref_library_get(void *h, long i, unsigned char *out, char *name_utf8, int name_capacity)
{
    try {
        const ImageLibrary *lib = static_cast<ImageLibrary *>(h);
        const cv::Mat &m = lib->getImages().at((size_t)i);
        for (int y = 0; y < m.rows; ++y)
            std::memcpy(out + (size_t)y * m.cols * 3, m.ptr<unsigned char>(y), (size_t)m.cols * 3);
        if (name_utf8 && name_capacity > 0) {
            std::strncpy(name_utf8, lib->getNames().at((size_t)i).toStdString().c_str(), (size_t)name_capacity - 1);
            name_utf8[name_capacity - 1] = 0;
        }
        return m.rows;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_library_remove(void *h, long i)
{
    try {
        static_cast<ImageLibrary *>(h)->removeAtIndex((size_t)i);
        return 0;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_library_save(void *h, const char *path)
{
    try {
        static_cast<ImageLibrary *>(h)->saveToFile(QString(path));
        return 0;
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}
int ref_library_load(void *h, const char *path)
{
    try {
        static_cast<ImageLibrary *>(h)->loadFromFile(QString(path));
        return 0;
    } catch (const std::invalid_argument &) {
        return -1;
    } catch (const std::exception &) {
        return -3;
    }
}
}
