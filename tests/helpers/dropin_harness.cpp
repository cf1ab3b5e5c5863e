// Drop-in check (test infrastructure): the reference-side binding integration/B200PhotomosaicGenerator.{h,cpp} compiled
// against the reference's REAL PhotomosaicGeneratorBase.h (Qt / OpenCV = the stand-ins of oracle/shim) and driven through
// the base-class API exactly like the reference drives its own back-ends (MainWindow.cpp:584-607, tst_CUDAGenerator.h:
// CPU generator vs CUDA generator on the same inputs). The base-class code (setters, getBestFits, buildPhotomosaic) is the
// reference's own PhotomosaicGeneratorBase.cpp from oracle/_ref/libref_core.so; the CPU back-end it is compared with is the
// reference's own CPUPhotomosaicGenerator.cpp from the same library.
// Built by oracle/Makefile into oracle/_ref/libdropin_b200.so (needs /root/reference for the headers; the prebuilt library
// travels to the GPU box).
#include <cstring>
#include <memory>

#include "B200PhotomosaicGenerator.h"
#include "CPUPhotomosaicGenerator.h"

namespace {
cv::Mat mat_from(const void *src, int rows, int cols, int type, size_t src_step = 0)
{
    cv::Mat m(rows, cols, type);
    const size_t row_bytes = (size_t)cols * m.elemSize();
    for (int y = 0; y < rows; ++y)
        std::memcpy(m.ptr<unsigned char>(y), (const unsigned char *)src + (size_t)y * (src_step ? src_step : row_bytes), row_bytes);
    return m;
}
}  // namespace

extern "C" {
// backend 0: the reference's CPUPhotomosaicGenerator, 1: B200PhotomosaicGenerator(device). Both behind
// std::shared_ptr<PhotomosaicGeneratorBase>, configured with the reference's own setters.
//   shapes / masks4 / ds / dmasks4: as ref_session_create (oracle/ref_generator_harness.cpp)
//   grids[s]: rows x cols int64, in: -1 nullopt / >= 0 valid, out: getBestFits()
//   mosaic_out (optional): buildPhotomosaic(background), rows x cols x 4
// Returns 0, 1 when generateBestFits() returned false, -3 on an exception.
int dropin_run(int backend, int device, const unsigned char *bgr, int rows, int cols, long stride, const unsigned char *lib, int n_lib,
               int lib_size, int n_steps, const int *const *shapes, const unsigned char *const *masks4, const int *ds,
               const unsigned char *const *dmasks4, double detail, int diff_type, int scheme, int repeat_range, int repeat_addition,
               const int *grid_rows, const int *grid_cols, long long *const *grids, const double background[4],
               unsigned char *mosaic_out, int *last_progress)
{
    try {
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
        std::shared_ptr<PhotomosaicGeneratorBase> generator;  // MainWindow.cpp:584-595 picks the back-end the same way
        if (backend == 0)
            generator = std::make_shared<CPUPhotomosaicGenerator>();
        else
            generator = std::make_shared<B200PhotomosaicGenerator>(device);

        generator->setMainImage(mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride));
        std::vector<cv::Mat> library;
        for (int i = 0; i < n_lib; ++i)
            library.push_back(mat_from(lib + (size_t)i * lib_size * lib_size * 3, lib_size, lib_size, CV_8UC3));
        generator->setLibrary(library);
        generator->setColourDifference(static_cast<ColourDifference::Type>(diff_type));
        generator->setColourScheme(static_cast<ColourScheme::Type>(scheme));
        generator->setCellGroup(group);
        GridUtility::MosaicBestFit state;
        for (int st = 0; st < n_steps; ++st) {
            GridUtility::StepBestFit step(grid_rows[st], std::vector<GridUtility::CellBestFit>(grid_cols[st]));
            for (int y = 0; y < grid_rows[st]; ++y)
                for (int x = 0; x < grid_cols[st]; ++x)
                    if (grids[st][(size_t)y * grid_cols[st] + x] >= 0)
                        step[y][x] = 0;
            state.push_back(step);
        }
        generator->setGridState(state);
        generator->setRepeat(repeat_range, repeat_addition);
        if (!generator->generateBestFits())
            return 1;
        const GridUtility::MosaicBestFit fits = generator->getBestFits();
        for (int st = 0; st < n_steps; ++st)
            for (int y = 0; y < grid_rows[st]; ++y)
                for (int x = 0; x < grid_cols[st]; ++x) {
                    const auto &v = fits[st][y][x];
                    grids[st][(size_t)y * grid_cols[st] + x] = v.has_value() ? (long long)v.value() : -1;
                }
        if (mosaic_out) {
            const cv::Mat m = generator->buildPhotomosaic(cv::Scalar(background[0], background[1], background[2], background[3]));
            for (int y = 0; y < m.rows; ++y)
                std::memcpy(mosaic_out + (size_t)y * m.cols * 4, m.ptr<unsigned char>(y), (size_t)m.cols * 4);
        }
        (void)last_progress;
        return 0;
    } catch (const std::exception &) {
        return -3;
    }
}
}
