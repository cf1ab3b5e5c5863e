// Host-only check of the C++ mirror's container methods (include/mosaic_b200.hpp): CellShape::loadFromFile / saveToFile and
// ImageLibrary::saveToFile / loadFromFile round trips. Built and run by tests/test_capi_host.py; needs no GPU.
//   usage: hpp_containers_check <in.mcs> <out.mcs> <out.mil>
#include <cstdio>

#include "../../include/mosaic_b200.hpp"

int main(int argc, char **argv)
{
    using namespace mosaicb200;
    if (argc < 4)
        return 64;
    try {
        CellShape shape;
        shape.loadFromFile(argv[1]);
        shape.saveToFile(argv[2]);
        CellShape again;
        again.loadFromFile(argv[2]);
        if (again.getCellMask() != shape.getCellMask() || again.getName() != shape.getName() ||
            again.getRowSpacing() != shape.getRowSpacing() || again.getColSpacing() != shape.getColSpacing())
            return 1;
        std::printf("%s %d %d %d\n", shape.getName().c_str(), shape.getSize(), shape.getRowSpacing(), shape.getColSpacing());

        bool threw = false;
        try {
            CellShape bad;
            bad.loadFromFile(argv[3]);  // does not exist yet
        } catch (const std::invalid_argument &) {
            threw = true;
        }
        if (!threw)
            return 2;
        return 0;
    } catch (const std::exception &e) {
        std::printf("failed: %s\n", e.what());
        return 3;
    }
}
