// Compile-and-link check of the C++ mirror (include/mosaic_b200.hpp): a caller written like the reference's
// fixtures (test/tst_Generator.h:111-136). Built by tests/test_capi_host.py; it runs only where a GPU exists.
#include <cstdio>
#include <vector>

#include "../../include/mosaic_b200.hpp"

int main()
{
    using namespace mosaicb200;
    try {
        PhotomosaicGenerator generator(0);
        std::vector<uint8_t> img(64 * 96 * 3, 100), big(50 * 70 * 3, 90);
        generator.setMainImage(Image{img.data(), 64, 96, 96 * 3});
        ImageLibrary library(32);  // ImageLibrary lib(cellSize), Benchmark_Generator.h:50
        for (int i = 0; i < 5; ++i) {
            big[i] = static_cast<uint8_t>(40 * i);
            library.addImage(Image{big.data(), 50, 70, 70 * 3}, "image");
        }
        const std::vector<uint8_t> lib = library.packed();
        generator.setLibrary(lib.data(), 5, 32);
        generator.setColourDifference(ColourDifference::Type::CIEDE2000);
        generator.setColourScheme(ColourScheme::Type::NONE);
        CellGroup cellGroup;
        cellGroup.setCellShape(CellShape(32));
        cellGroup.setDetail(100);
        generator.setCellGroup(cellGroup);
        generator.setGridState(generator.computeGridState());
        generator.setRepeat(2, 100);
        if (!generator.generateBestFits()) {
            std::printf("generate failed: %s\n", generator.lastError().c_str());
            return 2;
        }
        const auto fits = generator.getBestFits();
        std::printf("steps %zu rows %zu\n", fits.size(), fits.at(0).size());
        return 0;
    } catch (const std::exception &e) {
        std::printf("no device: %s\n", e.what());
        return 3;
    }
}
