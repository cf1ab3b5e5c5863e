// Stand-in for src/CellShape/CellGroup.h: per size step a normal and a detail CellShape (CellGroup.h:47-49).
#pragma once
#include "..\CellShape\CellShape.h"
class CellGroup {
public:
    std::vector<CellShape> cells, detailCells;
    CellShape &getCell(const size_t t_sizeStep, const bool t_detail = false) { return t_detail ? detailCells.at(t_sizeStep) : cells.at(t_sizeStep); }
    const CellShape &getCell(const size_t t_sizeStep, const bool t_detail = false) const { return t_detail ? detailCells.at(t_sizeStep) : cells.at(t_sizeStep); }
    size_t getSizeSteps() const { return cells.empty() ? 0 : cells.size() - 1; }
    int getCellSize(const size_t t_sizeStep, const bool t_detail = false) const { return getCell(t_sizeStep, t_detail).getSize(); }
    double detail = 1.0;
    double getDetail() const { return detail; }
};
