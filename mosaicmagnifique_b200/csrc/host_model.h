// Qt-free, OpenCV-free host model of the reference's domain types on the best-fit path:
//   CellShape      src/CellShape/CellShape.{h,cpp}   (mask + tiling parameters, resized())
//   CellGroup      src/CellShape/CellGroup.{h,cpp}   (per size step: normal cell + detail cell)
//   GridUtility    src/Grid/GridUtility.{h,cpp}      (grid size, cell rect, flip state)
//   GridGenerator  src/Grid/GridGenerator.{h,cpp}    (valid / split decision per cell), GridBounds
// plus the OpenCV arithmetic those use on the host (INTER_AREA for 8U, BGR2GRAY, entropy).
#pragma once
#include <math.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

namespace mm {

constexpr int kPadGrid = 2;  // GridUtility::PAD_GRID, GridUtility.h:33

struct Shape {
    std::vector<uint8_t> mask;  // size x size; non-zero = active (0 / 255 unless the shape was loaded from a file with grey values)
    int size = 0;
    int row_spacing = 0, col_spacing = 0, alt_row_spacing = 0, alt_col_spacing = 0;
    int alt_row_offset = 0, alt_col_offset = 0;
    bool alt_col_flip_h = false, alt_col_flip_v = false, alt_row_flip_h = false, alt_row_flip_v = false;

    bool empty() const { return size == 0; }
    // setCellMask: THRESH_BINARY at 127 (CellShape.cpp:116-135)
    void set_mask(const uint8_t *m, int s);
    // CellShape::loadFromFile keeps the decoded mask as stored (CellShape.cpp:405-410): no threshold. Every later test is
    // `mask != 0` (CPUPhotomosaicGenerator.cpp:150), and resized() thresholds the RESIZED mask, not this one
    void set_mask_as_stored(const uint8_t *m, int s);
    // the four flipped masks, index = flip_h + 2 * flip_v (CellShape::getCellMask, CellShape.cpp:138-152)
    std::vector<uint8_t> masks4() const;
    // CellShape::resized (CellShape.cpp:281-312); false + err when the resize is not supported
    bool resized(int new_size, Shape &out, std::string &err) const;
};

struct Group {
    std::vector<Shape> cells, detail_cells;
    double detail = 1.0;
    int size_steps = 0;
    // CellGroup::setCellShape / setDetail / setSizeSteps (CellGroup.cpp:29-128)
    bool build(const Shape &top, int detail_percent, int steps, std::string &err);
};

struct Rect {
    int x = 0, y = 0, w = 0, h = 0;
};

void grid_size(const Shape &s, int image_w, int image_h, int pad, int &gx, int &gy);  // GridUtility.cpp:25-52
Rect rect_at(const Shape &s, int x, int y);                                            // GridUtility.cpp:86-115
int flip_at(const Shape &s, int x, int y);                                             // GridUtility.cpp:118-132

// GridBounds::mergeBounds (GridBounds.cpp:39-104), same iteration order
void merge_bounds(std::vector<Rect> &b);

// getCellAt's bound arithmetic (PhotomosaicGeneratorBase.cpp:296-326): detail-space bound of cell (x, y)
Rect detail_bound(const Shape &normal, int detail_size, double detail, int x, int y, int image_w, int image_h,
                  Rect *clamped_global = nullptr, Rect *local = nullptr);

// cv::resize(..., INTER_AREA) for 8U, cn channels, scale >= 1 in both directions (OpenCV's resizeAreaFast_ for
// integer ratios, resizeArea_ otherwise). Returns false for up-scaling.
bool resize_area_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw);

// cv::computeResizeAreaTab as a CSR table for the GPU kernels: destination sample d covers source samples
// si[start[d] .. start[d+1]) with weights alpha (OpenCV's fractional-coverage weights, float)
struct AreaTable {
    std::vector<int> start, si;
    std::vector<float> alpha;
};
AreaTable make_area_table(int ssize, int dsize);

// cv::resize(..., INTER_CUBIC) for 8U, cn channels, as OpenCV's own (non-IPP) code computes it: per destination sample the
// four clamped source indices and the 11-bit fixed-point cubic coefficients (A = -0.75)
struct CubicTable {
    std::vector<int> idx;       // [dsize][4]
    std::vector<int16_t> coef;  // [dsize][4]
};
CubicTable make_cubic_table(int ssize, int dsize);
bool resize_cubic_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw);
// vertical pass of one sample from the four horizontally filtered rows (int, scaled by 2^11): the first (row / 8) * 8
// elements of a row go through OpenCV's float SIMD path (unfused mul + add, round half to even), the tail through the
// fixed-point cast. Shared by the host implementation and the CUDA kernel.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint8_t cubic_vertical_u8(int s0, int s1, int s2, int s3, const int16_t b[4], float scale, bool vec_path)
{
    int r;
    if (vec_path) {
        const float b0 = (float)b[0] * scale, b1 = (float)b[1] * scale, b2 = (float)b[2] * scale, b3 = (float)b[3] * scale;
#if defined(__CUDA_ARCH__)
        float v = __fmul_rn((float)s3, b3);
        v = __fadd_rn(__fmul_rn((float)s2, b2), v);
        v = __fadd_rn(__fmul_rn((float)s1, b1), v);
        v = __fadd_rn(__fmul_rn((float)s0, b0), v);
        r = __float2int_rn(v);
#else
        volatile float v = (float)s3 * b3;  // volatile: no contraction, every product and sum rounded on its own
        volatile float p = (float)s2 * b2;
        v = p + v;
        p = (float)s1 * b1;
        v = p + v;
        p = (float)s0 * b0;
        v = p + v;
        r = (int)lrintf(v);
#endif
    } else {
        r = (s0 * b[0] + s1 * b[1] + s2 * b[2] + s3 * b[3] + (1 << 21)) >> 22;
    }
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

// cvtColor(COLOR_BGR2GRAY) for 8U (OpenCV fixed point) and ImageUtility::calculateEntropy (ImageUtility.cpp:189-242)
void bgr_to_gray_u8(const uint8_t *bgr, size_t n, uint8_t *gray);
double masked_entropy(const uint8_t *gray, const uint8_t *mask, size_t n);

// GridGenerator::getGridState (GridGenerator.cpp:29-110). grids[step][y * cols + x] = -1 (nullopt) or 0 (valid).
struct GridStep {
    int rows = 0, cols = 0;
    std::vector<int64_t> v;
};
// one grid cell that lies inside an active bound: what the entropy rule needs to decide "split or keep"
struct EntropyCandidate {
    int x = 0, y = 0;  // unpadded grid coordinates
    Rect cg;           // cell rect clamped to the image (the pixels the reference crops, GridGenerator.cpp:145-164)
    Rect db;           // detail-space bound = size the crop is resized to and window of the detail mask (:166-185)
    int flip = 0;      // mask index: flip_h + 2 * flip_v
};
// split[i] = 1 <=> masked grey-level entropy of candidate i >= 0.7 * 8 bits (GridGenerator.cpp:187-188)
using EntropyEvaluator =
    std::function<bool(int step, const std::vector<EntropyCandidate> &, std::vector<uint8_t> &split, std::string &err)>;
// rows / cols: image size. A null evaluator never splits (what the reference does without a main image).
bool compute_grid_state(const Group &g, int rows, int cols, const EntropyEvaluator &evaluate, std::vector<GridStep> &out,
                        std::string &err);
// the reference's host arithmetic (OpenCV-compatible INTER_AREA, BGR2GRAY, f64 entropy); the product uses the GPU evaluator
EntropyEvaluator host_entropy_evaluator(const Group &g, const uint8_t *main_bgr, size_t row_stride);

}  // namespace mm
