// Entropy rule of the grid-state generator on the GPU: one CTA per candidate cell.
//
// Replaces the per-cell host work of GridGenerator::findCellState (src/Grid/GridGenerator.cpp:143-188) and
// ImageUtility::calculateEntropy (src/Other/ImageUtility.cpp:189-242): crop the visible part of the 8U main image,
// cv::resize it (INTER_AREA) to the size of the cell's detail-space bound, BGR2GRAY, 256-bin histogram of the pixels
// under the (flipped, bounded) detail mask, Shannon entropy in f64, split <=> entropy >= 0.7 * 8 bits.
// OpenCV arithmetic is reproduced exactly (tests compare the resulting grid state with the cv2-based oracle):
//   * integer ratios: block sums, (s + 2) >> 2 for 2 x 2, otherwise cvRound(sum * float(1 / area));
//   * other ratios: resizeArea_'s fractional-coverage taps, computed on the fly in f64 like computeResizeAreaTab,
//     accumulated in f32 in table order, cvRound;
//   * BGR2GRAY: 15-bit fixed point (3735, 19235, 9798).
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

namespace mm {

namespace {

constexpr int kGridThreads = 256;

// taps of destination sample d (cv::computeResizeAreaTab): up to `first partial`, full samples [sx1, sx2), `last partial`
struct Taps {
    int sx1, sx2;        // full-weight samples
    int pre;             // index of the leading partial sample or -1
    int post;            // index of the trailing partial sample or -1
    float a_pre, a_full, a_post;
};

__device__ __forceinline__ Taps make_taps(int ssize, double scale, int d)
{
    Taps t;
    const double fsx1 = d * scale, fsx2 = fsx1 + scale;
    const double cell_width = fmin(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = min(sx2, ssize - 1);
    sx1 = min(sx1, sx2);
    t.sx1 = sx1;
    t.sx2 = sx2;
    t.pre = (sx1 - fsx1 > 1e-3) ? sx1 - 1 : -1;
    t.a_pre = (float)((sx1 - fsx1) / cell_width);
    t.a_full = (float)(1.0 / cell_width);
    t.post = (fsx2 - sx2 > 1e-3) ? sx2 : -1;
    t.a_post = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell_width) / cell_width);
    return t;
}

// one row of the horizontal pass: buf = sum_k S[sx_k] * alpha_k, accumulated from 0 in tap order
template <typename F>
__device__ __forceinline__ float row_pass(const Taps &tx, F pix)
{
    float buf = 0.0f;
    if (tx.pre >= 0)
        buf = __fadd_rn(buf, __fmul_rn(pix(tx.pre), tx.a_pre));
    for (int sx = tx.sx1; sx < tx.sx2; ++sx)
        buf = __fadd_rn(buf, __fmul_rn(pix(sx), tx.a_full));
    if (tx.post >= 0)
        buf = __fadd_rn(buf, __fmul_rn(pix(tx.post), tx.a_post));
    return buf;
}

__device__ __forceinline__ int round_u8(float v) { return min(max(__float2int_rn(v), 0), 255); }

}  // namespace

__global__ void __launch_bounds__(kGridThreads)
grid_entropy_kernel(const uint8_t *__restrict__ main_bgr, int W, const GridCandidate *__restrict__ cand, const uint8_t *__restrict__ masks4,
                    int ds, double threshold, uint8_t *__restrict__ split)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned count;
    const GridCandidate c = cand[blockIdx.x];
    hist[threadIdx.x] = 0;
    if (threadIdx.x == 0)
        count = 0;
    __syncthreads();
    if (c.cw <= 0 || c.ch <= 0) {  // nothing of the cell is inside the image: entropy 0, never split
        if (threadIdx.x == 0)
            split[blockIdx.x] = 0;
        return;
    }
    const uint8_t *mask = masks4 + (size_t)c.flip * ds * ds;
    const uint8_t *src = main_bgr + ((size_t)c.cy * W + c.cx) * 3;
    const bool same = c.ch == c.bh && c.cw == c.bw;
    const double scale_x = (double)c.cw / c.bw, scale_y = (double)c.ch / c.bh;
    const int kx = (int)floor(scale_x + 0.5), ky = (int)floor(scale_y + 0.5);
    const bool fast = fabs(scale_x - kx) < DBL_EPSILON && fabs(scale_y - ky) < DBL_EPSILON;

    for (int i = threadIdx.x; i < c.bw * c.bh; i += kGridThreads) {
        const int dy = i / c.bw, dx = i - dy * c.bw;
        if (mask[(size_t)(c.by + dy) * ds + c.bx + dx] == 0)
            continue;
        int px[3];
        if (same) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                px[ch] = src[((size_t)dy * W + dx) * 3 + ch];
        } else if (fast) {
            const float scale = 1.0f / (float)(kx * ky);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                int sum = 0;
                for (int yy = 0; yy < ky; ++yy)
                    for (int xx = 0; xx < kx; ++xx)
                        sum += src[((size_t)(dy * ky + yy) * W + dx * kx + xx) * 3 + ch];
                px[ch] = (kx == 2 && ky == 2) ? ((sum + 2) >> 2) : round_u8(__fmul_rn((float)sum, scale));
            }
        } else {
            const Taps tx = make_taps(c.cw, scale_x, dx), ty = make_taps(c.ch, scale_y, dy);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float sum = 0.0f;
                bool first = true;
                auto add_row = [&](int sy, float beta) {
                    const uint8_t *row = src + (size_t)sy * W * 3 + ch;
                    const float buf = row_pass(tx, [&](int sx) { return (float)row[(size_t)sx * 3]; });
                    const float term = __fmul_rn(beta, buf);
                    sum = first ? term : __fadd_rn(sum, term);
                    first = false;
                };
                if (ty.pre >= 0)
                    add_row(ty.pre, ty.a_pre);
                for (int sy = ty.sx1; sy < ty.sx2; ++sy)
                    add_row(sy, ty.a_full);
                if (ty.post >= 0)
                    add_row(ty.post, ty.a_post);
                px[ch] = round_u8(sum);
            }
        }
        const int gray = (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
        atomicAdd(&hist[gray], 1u);
        atomicAdd(&count, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // ImageUtility.cpp:233-240: bins in ascending order, f64
        double e = 0.0;
        const double n = (double)count;
        for (int b = 0; b < 256; ++b) {
            const double p = hist[b] / n;
            if (p > 0)
                e -= p * log2(p);
        }
        split[blockIdx.x] = (count > 0 && e >= threshold) ? 1 : 0;
    }
}

cudaError_t launch_grid_entropy(const uint8_t *main_bgr, int W, const GridCandidate *cand, int n_cand, const uint8_t *masks4, int ds,
                                double threshold, uint8_t *split, cudaStream_t stream)
{
    if (n_cand <= 0)
        return cudaSuccess;
    grid_entropy_kernel<<<n_cand, kGridThreads, 0, stream>>>(main_bgr, W, cand, masks4, ds, threshold, split);
    return cudaGetLastError();
}

}  // namespace mm
