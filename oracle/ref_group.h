// Shared by the harnesses (oracle/ref_wrap.cpp, oracle/ref_generator_harness.cpp, tests/helpers/dropin_harness.cpp): builds the
// reference's REAL CellShape / CellGroup objects (src/CellShape/CellShape.cpp, CellGroup.cpp compiled unmodified) from plain
// arrays, through the reference's own constructors and setters. Test infrastructure only.
#pragma once
#include <cstring>

#include "CellGroup.h"

// rows x cols elements of `type` copied out of caller memory (src_step = bytes per source row, 0 = contiguous)
inline cv::Mat ref_mat_from(const void *src, int rows, int cols, int type, size_t src_step = 0)
{
    cv::Mat m(rows, cols, type);
    const size_t row_bytes = (size_t)cols * m.elemSize();
    for (int y = 0; y < rows; ++y)
        std::memcpy(m.ptr<unsigned char>(y), (const unsigned char *)src + (size_t)y * (src_step ? src_step : row_bytes), row_bytes);
    return m;
}

// p: size, rowSpacing, colSpacing, altRowSpacing, altColSpacing, altRowOffset, altColOffset, colFlipH, colFlipV, rowFlipH, rowFlipV;
// mask: size x size 8U (NULL = the default all-255 square cell, CellShape(size_t))
inline CellShape ref_make_shape(const int *p, const unsigned char *mask)
{
    CellShape s = mask ? CellShape(ref_mat_from(mask, p[0], p[0], CV_8UC1)) : CellShape(static_cast<size_t>(p[0]));
    s.setRowSpacing(p[1]);
    s.setColSpacing(p[2]);
    s.setAlternateRowSpacing(p[3]);
    s.setAlternateColSpacing(p[4]);
    s.setAlternateRowOffset(p[5]);
    s.setAlternateColOffset(p[6]);
    s.setAlternateColFlipHorizontal(p[7] != 0);
    s.setAlternateColFlipVertical(p[8] != 0);
    s.setAlternateRowFlipHorizontal(p[9] != 0);
    s.setAlternateRowFlipVertical(p[10] != 0);
    return s;
}

// CellGroup as the application builds it (MainWindow / tst_Generator.h:111-125): setCellShape, setDetail, setSizeSteps
inline CellGroup ref_make_group(const int *shape, const unsigned char *mask, int detail_percent, int size_steps)
{
    CellGroup g;
    g.setCellShape(ref_make_shape(shape, mask));
    g.setDetail(detail_percent);
    g.setSizeSteps(static_cast<size_t>(size_steps));
    return g;
}
