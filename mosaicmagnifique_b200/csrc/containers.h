// Qt-free, OpenCV-free readers / writers of the reference's on-disk containers (host side of the path):
//   .mcs cell shape      CellShape::saveToFile / loadFromFile      src/CellShape/CellShape.cpp:321-434
//   .mil image library   ImageLibrary::saveToFile / loadFromFile   src/ImageLibrary/ImageLibrary.cpp:117-236
//   cv::Mat in a stream  CustomQDataStream                         src/Other/CustomQDataStream.h:22-87
// Both are QDataStream (Qt_5_0) streams: big-endian integers, QString = u32 byte length + UTF-16BE (0xFFFFFFFF = null),
// QByteArray = u32 length + bytes, bool = one byte. Images are PNG (what cv::imencode(".png") writes and cv::imdecode reads:
// 8-bit grey / RGB / RGBA, non-interlaced) or, in older files, raw (type, rows, cols, bytes). The PNG codec (inflate, the five
// row filters, CRC) is written out here so that the library keeps no dependency beyond the CUDA runtime.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "host_model.h"

namespace mm {

// 8-bit image, channels in OpenCV order (grey, BGR, BGRA)
struct Image8 {
    int rows = 0, cols = 0, channels = 0;
    std::vector<uint8_t> px;
};

bool png_decode(const uint8_t *data, size_t n, Image8 &out, std::string &err);
// per-row filter choice (None / Sub / Up / Average / Paeth by minimum absolute residual) + one deflate block: LZ77 matches over a
// 32 KB window, fixed or dynamic Huffman codes, whichever is shorter. A 512 x 512 cell mask is a few KB like the files the reference
// ships, 128 px photographic tiles ~35 % of their raw size (round 1 wrote stored blocks: 100 %)
void png_encode(const Image8 &img, std::vector<uint8_t> &out);

struct McsFile {
    std::string name;  // UTF-8
    Shape shape;       // mask + tiling parameters (mask is NOT re-thresholded: loadFromFile stores it as decoded)
    uint32_t version = 0;
};
bool load_mcs(const char *path, McsFile &out, std::string &err);
bool save_mcs(const char *path, const McsFile &in, std::string &err);

struct MilFile {
    int image_size = 0;
    uint32_t version = 0;
    std::vector<std::string> names;  // UTF-8
    std::vector<uint8_t> images;     // n x image_size x image_size x 3 (BGR)
};
bool load_mil(const char *path, MilFile &out, std::string &err);
bool save_mil(const char *path, const MilFile &in, std::string &err);

}  // namespace mm
