// Fused masked difference-sum kernel: D[cell, lib] = sum_p w[cell,p] * diff(cell[p], lib[p]).
//
// Replaces, in ONE launch per size step, the reference's per-(cell, library image) chain
//   imageDifference / imageDifferenceEdge  src/Photomosaic/CUDA/PhotomosaicGenerator.cu:35-72
//   reduceAdd tree + D2D copies            src/Photomosaic/CUDA/Reduction.cu:24-211
//   flattenKernel                          src/Photomosaic/CUDA/PhotomosaicGenerator.cu:201-209
// and, when no repeat penalty is active, findLowestKernel (:175-189) through the argmin epilogue.
// CPU semantics followed: CPUPhotomosaicGenerator.cpp:137-169 (masked, bounded sum; strict <,
// lowest index wins).
//
// Data layout (built by prep_kernels.cu):
//   lib  : float4 (x0,x1,x2,-)  [lib_tile][chunk][TNB][KP]   one contiguous 16 KB block per (tile, chunk)
//          CIEDE2000: stored-scale channels (L/2-25, a/50, b/50, C/50), w = 50, and image PAIRS interleaved for the packed
//          FP32 path: [lib_tile][chunk][TNB/2][2][KP] float4 = (L0,L1,a0,a1) then (b0,b1,C0,C1)
//   cell : float4 (x0,x1,x2,C)  [cell_tile][chunk][TCB][KP]  followed by float w[TCB][KP] -> 20 KB block
//   w = 1 where the (flipped) detail mask is set and the pixel lies inside the cell's detail-space bound,
//   else 0; pixels are stored in the compacted order of the step's active-pixel list, padded with w = 0.
//
// Kernel shape: one CTA = 8 consumer warps + 1 producer warp. The producer lane streams (cell block,
// lib block) pairs into a 3-stage shared-memory ring with cp.async.bulk (TMA bulk copy, UBLKCP in SASS)
// completing on "full" mbarriers; consumers release stages through "empty" mbarriers. Consumer warp w
// owns cell w of the tile, its 32 lanes stride the chunk's pixels (conflict-free LDS.128), and every lane
// keeps TNB running sums, reduced with warp shuffles at the end. Roofline: FP32 + MUFU issue.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "colour_math.cuh"
#include "kernels.h"

namespace mm {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#ifndef MM_PIXEL_UNROLL
#define MM_PIXEL_UNROLL 1
#endif
constexpr int kPixelUnroll = MM_PIXEL_UNROLL;
constexpr int kStages = MM_STAGES;
constexpr int kConsumerWarps = MM_TCB;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr uint32_t kLibBlockBytes = MM_TNB * MM_KP * 16;
constexpr uint32_t kCellBlockBytes = MM_TCB * MM_KP * 20;
constexpr uint32_t kStageBytes = kLibBlockBytes + kCellBlockBytes;

template <int DIFF>
__global__ void __launch_bounds__(kThreads, MM_MIN_CTAS)
diff_sum_kernel(const unsigned char *__restrict__ cells, const unsigned char *__restrict__ lib, float *__restrict__ D,
                unsigned long long *__restrict__ best_key, int n_chunks, int n_lib, int n_lib_pad, int n_cells,
                int n_cell_tiles, int n_lib_tiles, const int *__restrict__ cancel, unsigned long long *__restrict__ progress,
                int nk, int sb_a, int sb_b)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];
    __shared__ int s_cancelled;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cell_tile, lib_tile;
    {
        // super-block raster (kernels.h: Raster): keep a column of sb_a cell tiles and sweep the library in bands of sb_b tiles
        const unsigned id = blockIdx.x;
        const unsigned col_all = (unsigned)sb_a * (unsigned)n_lib_tiles;
        const unsigned cb = id / col_all;
        const unsigned r = id - cb * col_all;
        const int cw = min(sb_a, n_cell_tiles - (int)cb * sb_a);
        const unsigned band_ctas = (unsigned)cw * (unsigned)sb_b;
        const unsigned band = r / band_ctas;
        const unsigned r2 = r - band * band_ctas;
        cell_tile = (int)cb * sb_a + (int)(r2 % cw);
        lib_tile = (int)band * sb_b + (int)(r2 / cw);
    }

    if (threadIdx.x == 0) {
        const int cancelled = cancel ? load_cancel_flag(cancel) : 0;  // device word (L2 hit), in flight during the barrier set-up
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cancelled = cancelled;
    }
    __syncthreads();
    if (s_cancelled)
        return;  // cancel(): nothing has been issued yet, the whole CTA leaves

    if (warp == kConsumerWarps) {
        // ---------------- producer warp: one elected lane drives the TMA ring
        if (lane == 0) {
            const unsigned char *cell_src = cells + (size_t)cell_tile * n_chunks * kCellBlockBytes;
            const unsigned char *lib_src = lib + (size_t)lib_tile * n_chunks * kLibBlockBytes;
            for (int k = 0; k < nk; ++k) {
                const int s = k % kStages;
                if (k >= kStages)
                    mbar_wait(&empty_bar[s], ((k / kStages) - 1) & 1);
                unsigned char *dst = smem + (size_t)s * kStageBytes;
                if (MM_STRESS_SKEW)
                    __nanosleep((unsigned)((k * 131 + blockIdx.x * 17) % 300));
                mbar_expect_tx(&full_bar[s], kStageBytes);
                bulk_g2s(dst, lib_src + (size_t)k * kLibBlockBytes, kLibBlockBytes, &full_bar[s]);
                bulk_g2s(dst + kLibBlockBytes, cell_src + (size_t)k * kCellBlockBytes, kCellBlockBytes, &full_bar[s]);
            }
        }
        return;
    }

    // ---------------- consumer warps: warp = cell within the tile, lanes stride pixels
    float acc[MM_TNB];
#pragma unroll
    for (int i = 0; i < MM_TNB; ++i)
        acc[i] = 0.0f;

    for (int k = 0; k < nk; ++k) {
        const int s = k % kStages;
        mbar_wait(&full_bar[s], (k / kStages) & 1);
        if (MM_STRESS_SKEW)
            __nanosleep((unsigned)((warp * 97 + k * 29 + blockIdx.x * 7) % 400));
        const float4 *lib_s = reinterpret_cast<const float4 *>(smem + (size_t)s * kStageBytes);
        const float4 *cell_s = reinterpret_cast<const float4 *>(smem + (size_t)s * kStageBytes + kLibBlockBytes) + warp * MM_KP;
        const float *w_s = reinterpret_cast<const float *>(smem + (size_t)s * kStageBytes + kLibBlockBytes + MM_TCB * MM_KP * 16) + warp * MM_KP;
#pragma unroll kPixelUnroll
        for (int j = 0; j < MM_KP / 32; ++j) {
            const int p = j * 32 + lane;
            const float4 c = cell_s[p];
            const float w = w_s[p];
            if (DIFF == MM_DIFF_CIEDE2000) {
                // packed FP32: two library images per lane vector. The tile stores image pairs interleaved:
                // [pair][0][p] = (L0, L1, a0, a1), [pair][1][p] = (b0, b1, C0, C1)  (prep_kernels.cu)
#pragma unroll
                for (int i = 0; i < MM_TNB / 2; ++i) {
                    const float4 la = lib_s[(i * 2 + 0) * MM_KP + p];
                    const float4 lb = lib_s[(i * 2 + 1) * MM_KP + p];
                    const mm_f2 d = mm_ciede2000_stored_v<mm_f2>(c.x, c.y, c.z, c.w, mm_f2{la.x, la.y}, mm_f2{la.z, la.w},
                                                               mm_f2{lb.x, lb.y}, mm_f2{lb.z, lb.w});
                    const mm_f2 a2 = v_fma(mm_f2{w, w}, d, mm_f2{acc[2 * i], acc[2 * i + 1]});
                    acc[2 * i] = a2.x;
                    acc[2 * i + 1] = a2.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < MM_TNB; ++i) {
                    const float4 l = lib_s[i * MM_KP + p];
                    acc[i] = fmaf(w, mm_euclid(c.x, c.y, c.z, l.x, l.y, l.z), acc[i]);
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&empty_bar[s]);
    }

    // ---------------- epilogue: warp-shuffle reduction, D store, fused argmin
#pragma unroll
    for (int i = 0; i < MM_TNB; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[i] = v;
    }
    const int cell = cell_tile * MM_TCB + warp;
    if (lane < MM_TNB) {
        float v = 0.0f;
#pragma unroll
        for (int i = 0; i < MM_TNB; ++i)
            v = (lane == i) ? acc[i] : v;
        const int li = lib_tile * MM_TNB + lane;
        if (D)
            D[(size_t)cell * n_lib_pad + li] = v;
        if (best_key && li < n_lib && cell < n_cells) {
            // non-negative floats order like their bit patterns; low word = index -> lowest index wins ties
            const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)li;
            atomicMin(best_key + cell, key);
        }
    }
    if (progress && threadIdx.x == 0)
        atomicAdd(progress, 1ull);  // consumer warp 0 has stored its results; an approximate "tiles done" count is all that is needed
}

template <int DIFF>
static cudaError_t launch(const void *cells, const void *lib, float *D, unsigned long long *best_key, int n_cell_tiles,
                          int n_lib_tiles, int n_chunks, int n_lib, int n_cells, cudaStream_t stream, const int *cancel,
                          unsigned long long *progress, Raster raster, size_t seg_stride)
{
    const size_t smem = (size_t)kStages * kStageBytes;
    // per device (context) attribute: set on every launch, it is cheap
    cudaError_t e = cudaFuncSetAttribute(diff_sum_kernel<DIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    const unsigned grid = (unsigned)n_cell_tiles * (unsigned)n_lib_tiles;  // 1-D, super-block rasterisation (kernels.h)
    // One launch per pixel segment (split-K across LAUNCHES): the kernel itself only learns how many chunks to walk (nk) and gets
    // its tensors pre-offset to the segment's first chunk; the tile stride inside the packed tensors stays n_chunks. Keeping the
    // segment out of the kernel keeps its inner loop's instruction schedule exactly the one measured fastest in round 1 -- ptxas
    // re-orders the 461-instruction loop body on the slightest change of the surrounding code, worth +-1 %.
    for (int seg = 0; seg < raster.n_segs; ++seg) {
        const int k0 = raster.n_segs > 1 ? seg * raster.seg_chunks : 0;
        const int nk = raster.n_segs > 1 ? std::min(n_chunks, k0 + raster.seg_chunks) - k0 : n_chunks;
        diff_sum_kernel<DIFF><<<grid, kThreads, smem, stream>>>(
            (const unsigned char *)cells + (size_t)k0 * kCellBlockBytes, (const unsigned char *)lib + (size_t)k0 * kLibBlockBytes,
            D ? D + (size_t)seg * seg_stride : nullptr, best_key, n_chunks, n_lib, n_lib_tiles * MM_TNB, n_cells, n_cell_tiles,
            n_lib_tiles, cancel, progress, nk, raster.sb_a, raster.sb_b);
        e = cudaGetLastError();
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

cudaError_t launch_diff_sum(int diff_type, const void *cells, const void *lib, float *D, unsigned long long *best_key,
                            int n_cell_tiles, int n_lib_tiles, int n_chunks, int n_lib, int n_cells, cudaStream_t stream,
                            const int *cancel, unsigned long long *progress, Raster raster, size_t seg_stride)
{
    if (n_cell_tiles <= 0 || n_lib_tiles <= 0 || n_chunks <= 0)
        return cudaSuccess;
    if (diff_type != MM_DIFF_CIEDE2000)
        return cudaErrorInvalidValue;  // RGB Euclidean / CIE76 run diff_euclid_kernel (diff_euclid.cu)
    if (raster.n_segs < 1 || raster.sb_a < 1 || raster.sb_b < 1 || (raster.n_segs > 1 && (best_key || !D || raster.seg_chunks < 1)))
        return cudaErrorInvalidValue;
    return launch<MM_DIFF_CIEDE2000>(cells, lib, D, best_key, n_cell_tiles, n_lib_tiles, n_chunks, n_lib, n_cells, stream, cancel, progress,
                                     raster, seg_stride);
}

// ---------------------------------------------------------------- raster choice + segment reduction

Raster choose_raster(int n_cell_tiles, int n_lib_tiles, int n_chunks, bool can_split)
{
    auto env = [](const char *name) -> int {
        const char *v = getenv(name);
        return v ? atoi(v) : 0;
    };
    // Measured on config 4 (ncu dram__bytes_read, profiles/r2_raster_sweep.txt), segments / sb_a x sb_b -> DRAM reads, kernel time:
    //   1 / 16 x 16: 47.4 GB, 893.8 ms     2 / 32 x 8: 30.6 GB, 894.4 ms     3 / 32 x 8: 22.5 GB, 895.2 ms     4 / 32 x 8: 22.0 GB, 895.7 ms
    //   2 / 16 x 16: 42.8 GB               1 / 24 x 12: 102.5 GB             1 / 32 x 8: 233 GB                 2 / 64 x 4: 251 GB
    // i.e. a hot set (sb_a cell tiles x one segment x 20 KB per chunk) of 42 MB survives in the 126 MB L2, 63 MB and more do not,
    // and every extra launch costs ~0.6 ms of tail. Default: segments of <= 64 chunks (8,192 pixels), the widest column whose hot
    // set stays <= 42 MB.
    Raster r{1, n_chunks, kSuperTiles, kSuperTiles};
    if (can_split && n_chunks > 64)
        r.n_segs = (n_chunks + 63) / 64;
    if (env("MM_SPLITK") > 0 && can_split)
        r.n_segs = std::min(env("MM_SPLITK"), std::max(n_chunks, 1));
    r.seg_chunks = (n_chunks + r.n_segs - 1) / r.n_segs;
    r.n_segs = (n_chunks + r.seg_chunks - 1) / std::max(r.seg_chunks, 1);
    const size_t seg_bytes = (size_t)std::max(r.seg_chunks, 1) * kCellBlockBytes;
    int a = (int)((size_t)(42u << 20) / seg_bytes);
    a = a >= 64 ? 64 : (a >= 32 ? 32 : 16);
    if (env("MM_SB_A") > 0)
        a = env("MM_SB_A");
    r.sb_a = std::max(1, std::min(a, n_cell_tiles));
    r.sb_b = std::max(1, 256 / r.sb_a);
    if (env("MM_SB_B") > 0)
        r.sb_b = env("MM_SB_B");
    r.sb_b = std::max(1, std::min(r.sb_b, n_lib_tiles));
    return r;
}

__global__ void sum_segments_kernel(float4 *__restrict__ D, int n_segs, size_t seg_stride4, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = D[i];
        for (int s = 1; s < n_segs; ++s) {  // fixed order: the result does not depend on how the launch was scheduled
            const float4 w = D[(size_t)s * seg_stride4 + i];
            v.x += w.x;
            v.y += w.y;
            v.z += w.z;
            v.w += w.w;
        }
        D[i] = v;
    }
}

cudaError_t launch_sum_segments(float *D, int n_segs, size_t seg_stride, size_t n, cudaStream_t stream)
{
    if (n_segs <= 1 || n == 0)
        return cudaSuccess;
    if ((seg_stride | n) & 3)
        return cudaErrorInvalidValue;  // rows are padded to the library tile (8 floats), so both are multiples of 4
    const size_t n4 = n / 4;
    const unsigned blocks = (unsigned)((n4 + 255) / 256 > 148 * 16 ? 148 * 16 : (n4 + 255) / 256);
    sum_segments_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<float4 *>(D), n_segs, seg_stride / 4, n4);
    return cudaGetLastError();
}

}  // namespace mm
