// B200PhotomosaicGenerator.cpp -- see the header. Everything the reference's back-ends do inside generateBestFits()
// (CPUPhotomosaicGenerator.cpp:33-112, CUDA/CUDAPhotomosaicGenerator.cpp:40-370) happens behind the C ABI.
#include "B200PhotomosaicGenerator.h"

#include <cmath>
#include <cstring>

#include "..\..\Other\Logger.h"

B200PhotomosaicGenerator::B200PhotomosaicGenerator(const int device)
{
    if (mosaic_create(device, &m_engine) != MOSAIC_OK)
        m_engine = nullptr;  // no usable CUDA device: generateBestFits() returns false, there is no CPU fallback
}

B200PhotomosaicGenerator::~B200PhotomosaicGenerator()
{
    mosaic_destroy(m_engine);
}

const char *B200PhotomosaicGenerator::lastError() const
{
    return m_engine ? mosaic_last_error(m_engine) : "no CUDA device";
}

bool B200PhotomosaicGenerator::generateBestFits()
{
    if (!m_engine || m_lib.empty() || m_bestFits.empty())
        return false;
    if (m_wasCanceled)  // cancel() before the call: the reference's loops leave at their first check (CPUPhotomosaicGenerator.cpp:52)
        return false;
    bool ok = true;
    auto chk = [&](int rc) { ok = ok && rc == MOSAIC_OK; };

    // inputs held by PhotomosaicGeneratorBase (PhotomosaicGeneratorBase.h:83-100)
    const cv::Mat img = m_img.isContinuous() ? m_img : m_img.clone();  // 8U BGR
    chk(mosaic_set_main_image(m_engine, img.data, img.rows, img.cols, img.step));

    const int S = m_lib.front().rows;  // library already at cell size (MainWindow.cpp:575-581)
    std::vector<uchar> lib(m_lib.size() * size_t(S) * S * 3);
    for (size_t i = 0; i < m_lib.size(); ++i)  // contiguous n x S x S x 3
        std::memcpy(&lib[i * size_t(S) * S * 3], m_lib[i].clone().data, size_t(S) * S * 3);
    chk(mosaic_set_library(m_engine, lib.data(), int64_t(m_lib.size()), S));

    chk(mosaic_set_colour_difference(m_engine, static_cast<int>(m_colourDiffType)));
    chk(mosaic_set_colour_scheme(m_engine, static_cast<int>(m_colourSchemeType)));

    const CellShape &top = m_cells.getCell(0);
    const mosaic_cell_shape cs{top.getSize(), top.getRowSpacing(), top.getColSpacing(),
                               top.getAlternateRowSpacing(), top.getAlternateColSpacing(),
                               top.getAlternateRowOffset(), top.getAlternateColOffset(),
                               top.getAlternateColFlipHorizontal(), top.getAlternateColFlipVertical(),
                               top.getAlternateRowFlipHorizontal(), top.getAlternateRowFlipVertical()};
    const cv::Mat mask = top.getCellMask(false, false).clone();
    chk(mosaic_set_cell_group(m_engine, &cs, mask.data, 0, int(std::lround(m_cells.getDetail() * 100)),
                              int(m_cells.getSizeSteps())));

    for (size_t step = 0; step < m_bestFits.size(); ++step)  // setGridState: optional -> validity map
    {
        const int rows = int(m_bestFits[step].size()), cols = int(m_bestFits[step][0].size());
        std::vector<uchar> valid(size_t(rows) * cols);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                valid[size_t(y) * cols + x] = m_bestFits[step][y][x].has_value();
        chk(mosaic_set_grid_state(m_engine, int(step), rows, cols, valid.data()));
    }
    chk(mosaic_set_repeat(m_engine, m_repeatRange, m_repeatAddition));
    // progress(int) is emitted from inside mosaic_generate on this thread. A connected QProgressDialog runs the cancel() slot
    // during the emission (it only sets m_wasCanceled, PhotomosaicGeneratorBase.cpp:217-220), so the flag is looked at right after
    // and handed to the engine, whose running kernel then stops scheduling work.
    mosaic_set_progress_callback(
        m_engine,
        [](int p, void *self) {
            B200PhotomosaicGenerator *gen = static_cast<B200PhotomosaicGenerator *>(self);
            emit gen->progress(p);
            if (gen->m_wasCanceled)
                mosaic_cancel(gen->m_engine);
        },
        this);

    const int rc = ok ? mosaic_generate(m_engine) : MOSAIC_ERR_INVALID_ARGUMENT;
    if (rc == MOSAIC_ERR_CANCELLED || m_wasCanceled)
        return false;  // as both reference back-ends: `return !m_wasCanceled` (CPUPhotomosaicGenerator.cpp:107-112)
    if (rc != MOSAIC_OK)
    {
        LogCritical(mosaic_last_error(m_engine));  // errors: text instead of the modal box of CUDAUtility.h:34-62
        return false;
    }

    for (size_t step = 0; step < m_bestFits.size(); ++step)  // getBestFits: -1 -> std::nullopt
    {
        const int rows = int(m_bestFits[step].size()), cols = int(m_bestFits[step][0].size());
        std::vector<int64_t> out(size_t(rows) * cols);
        if (mosaic_get_best_fits(m_engine, int(step), out.data(), rows, cols) != MOSAIC_OK)
            return false;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                m_bestFits[step][y][x] = out[size_t(y) * cols + x] < 0
                                             ? std::nullopt
                                             : std::optional<size_t>(size_t(out[size_t(y) * cols + x]));
    }
    return !m_wasCanceled;
}
