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
#include "ref_group.h"  // oracle/ref_group.h: real CellShape / CellGroup objects from plain arrays

extern "C" {
// backend 0: the reference's CPUPhotomosaicGenerator, 1: B200PhotomosaicGenerator(device). Both behind
// std::shared_ptr<PhotomosaicGeneratorBase>, configured with the reference's own setters.
//   shape (11 ints) / mask / detail_percent / size_steps: the top-level cell, as ref_session_create (oracle/ref_generator_harness.cpp);
//   the CellGroup is built by the reference's own CellGroup.cpp
//   grids[s]: rows x cols int64, in: -1 nullopt / >= 0 valid, out: getBestFits()
//   mosaic_out (optional): buildPhotomosaic(background), rows x cols x 4
// Returns 0, 1 when generateBestFits() returned false, -3 on an exception.
int dropin_run(int backend, int device, const unsigned char *bgr, int rows, int cols, long stride, const unsigned char *lib, int n_lib,
               int lib_size, const int *shape, const unsigned char *mask, int detail_percent, int size_steps, int diff_type, int scheme,
               int repeat_range, int repeat_addition, int n_steps, const int *grid_rows, const int *grid_cols, long long *const *grids,
               const double background[4], unsigned char *mosaic_out)
{
    try {
        const CellGroup group = ref_make_group(shape, mask, detail_percent, size_steps);
        std::shared_ptr<PhotomosaicGeneratorBase> generator;  // MainWindow.cpp:584-595 picks the back-end the same way
        if (backend == 0)
            generator = std::make_shared<CPUPhotomosaicGenerator>();
        else
            generator = std::make_shared<B200PhotomosaicGenerator>(device);

        generator->setMainImage(ref_mat_from(bgr, rows, cols, CV_8UC3, (size_t)stride));
        std::vector<cv::Mat> library;
        for (int i = 0; i < n_lib; ++i)
            library.push_back(ref_mat_from(lib + (size_t)i * lib_size * lib_size * 3, lib_size, lib_size, CV_8UC3));
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
        return 0;
    } catch (const std::exception &) {
        return -3;
    }
}
}
