// Stand-in for the reference's CellShape (src/CellShape/CellShape.h) exposing only the
// getters GridUtility.cpp and CPUPhotomosaicGenerator.cpp read; found through the reference's Windows-style include
// "..\CellShape\CellShape.h" (a literal file name on Linux).
#pragma once
#include <opencv2/core.hpp>
class CellShape {
public:
    int size = 0, rowSpacing = 0, colSpacing = 0, altRowSpacing = 0, altColSpacing = 0;
    int altRowOffset = 0, altColOffset = 0;
    bool colFlipH = false, colFlipV = false, rowFlipH = false, rowFlipV = false;
    int getSize() const { return size; }
    int getRowSpacing() const { return rowSpacing; }
    int getColSpacing() const { return colSpacing; }
    int getAlternateRowSpacing() const { return altRowSpacing; }
    int getAlternateColSpacing() const { return altColSpacing; }
    int getAlternateRowOffset() const { return altRowOffset; }
    int getAlternateColOffset() const { return altColOffset; }
    bool getAlternateColFlipHorizontal() const { return colFlipH; }
    bool getAlternateColFlipVertical() const { return colFlipV; }
    bool getAlternateRowFlipHorizontal() const { return rowFlipH; }
    bool getAlternateRowFlipVertical() const { return rowFlipV; }
    // the four flipped masks, index = horizontal + 2 * vertical (CellShape::getCellMask, CellShape.cpp:138-152)
    cv::Mat masks[4];
    const cv::Mat &getCellMask(const bool t_flippedHorizontal, const bool t_flippedVertical) const
    {
        return masks[(t_flippedHorizontal ? 1 : 0) + (t_flippedVertical ? 2 : 0)];
    }
};
