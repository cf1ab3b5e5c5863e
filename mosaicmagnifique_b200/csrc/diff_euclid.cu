// Fused masked difference-sum kernel for RGB Euclidean / CIE76: D[cell, lib] = sum_p w[cell,p] * |cell[p] - lib[p]|.
//
// Same role as diff_sum_kernel (diff_kernels.cu; reference: CUDA/PhotomosaicGenerator.cu:35-72 + Reduction.cu +
// flattenKernel, CPU semantics CPUPhotomosaicGenerator.cpp:137-169) but a different shape, because the Euclidean
// difference is only ~7 FP32 lane-ops + 1 MUFU.SQRT per pixel pair: operand traffic, not arithmetic, decides.
// The CIEDE2000 kernel's "warp = one cell, lanes = pixels" layout re-reads 4.5 B of shared memory / L2 per pair and
// saturates the shared-memory pipe at ~40 % of the MUFU ceiling (profiles/). This kernel is register-tiled like an
// SGEMM with the inner product replaced by sqrt(sum of squares):
//   * CTA tile = 64 cells x 64 library images, 256 consumer threads, thread tile = 4 cells x 4 images (16 running sums,
//     complete over all pixels -> no cross-lane reduction at the end);
//   * shared-memory chunks are pixel-major: cells [pixel][x0,x1,x2,w][64 cells], library [pixel][x0,x1,x2][64 images]
//     (library values stored NEGATED so that cell - lib is a packed add). One LDS.128 fetches a channel of the thread's
//     4 cells (two packed FP32 pairs) or of its 4 images (broadcast operands); a warp touches 16 + 2 distinct 16-byte
//     words per load pair, i.e. 0.7 shared-memory wavefronts per pixel pair-column instead of 4;
//   * packed FP32 (FADD2/FMUL2/FFMA2): 7 packed ops + 2 MUFU.SQRT per two pixel pairs -> MUFU-bound by design
//     (8 XU cycles vs 7 FMA-pipe cycles per pair);
//   * global -> shared: one 16 KB + one 12 KB cp.async.bulk (TMA bulk copy) per 16-pixel chunk through a 3-stage
//     mbarrier ring driven by a producer warp; 0.44 B of L2/HBM traffic per pixel pair.
#include <cuda_runtime.h>
#include <stdint.h>

#include "colour_math.cuh"
#include "kernels.h"

namespace mm {

namespace {

__device__ __forceinline__ uint32_t smem_u32e(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_e(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32e(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_e(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32e(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_e(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32e(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_e(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32e(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_e(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32e(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32e(bar))
                 : "memory");
}

constexpr int kEStages = MM_ESTAGES;
constexpr int kEConsumerWarps = 8;
constexpr int kEThreads = (kEConsumerWarps + 1) * 32;
constexpr uint32_t kECellBlock = MM_EKP * 4 * MM_ETC * 4;  // [pixel][x0,x1,x2,w][64 cells] f32
constexpr uint32_t kELibBlock = MM_EKP * 3 * MM_ETN * 4;   // [pixel][x0,x1,x2][64 images] f32 (negated)
constexpr uint32_t kEStage = kECellBlock + kELibBlock;

__device__ __forceinline__ float comp(const float4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

}  // namespace

__global__ void __launch_bounds__(kEThreads, 2)
diff_euclid_kernel(const unsigned char *__restrict__ cells, const unsigned char *__restrict__ lib, float *__restrict__ D,
                   unsigned long long *__restrict__ best_key, int n_chunks, int n_lib, int n_lib_pad, int n_cells,
                   int n_cell_tiles, int n_lib_tiles, const int *__restrict__ cancel, unsigned long long *__restrict__ progress)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[kEStages];
    __shared__ uint64_t empty_bar[kEStages];
    __shared__ int s_cancelled;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cell_tile, lib_tile;
    tile_of_block(blockIdx.x, n_cell_tiles, n_lib_tiles, cell_tile, lib_tile);

    if (threadIdx.x == 0) {
        const int cancelled = cancel ? load_cancel_flag(cancel) : 0;  // device word (L2 hit), in flight during the barrier set-up
        for (int s = 0; s < kEStages; ++s) {
            mbar_init_e(&full_bar[s], 1);
            mbar_init_e(&empty_bar[s], kEConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_cancelled = cancelled;
    }
    __syncthreads();
    if (s_cancelled)
        return;  // cancel(): nothing has been issued yet, the whole CTA leaves

    if (warp == kEConsumerWarps) {
        if (lane == 0) {
            const unsigned char *cell_src = cells + (size_t)cell_tile * n_chunks * kECellBlock;
            const unsigned char *lib_src = lib + (size_t)lib_tile * n_chunks * kELibBlock;
            for (int k = 0; k < n_chunks; ++k) {
                const int s = k % kEStages;
                if (k >= kEStages)
                    mbar_wait_e(&empty_bar[s], ((k / kEStages) - 1) & 1);
                unsigned char *dst = smem + (size_t)s * kEStage;
                if (MM_STRESS_SKEW)
                    __nanosleep((unsigned)((k * 131 + blockIdx.x * 17) % 300));
                mbar_expect_tx_e(&full_bar[s], kEStage);
                bulk_g2s_e(dst, cell_src + (size_t)k * kECellBlock, kECellBlock, &full_bar[s]);
                bulk_g2s_e(dst + kECellBlock, lib_src + (size_t)k * kELibBlock, kELibBlock, &full_bar[s]);
            }
        }
        return;
    }

    // consumer thread -> 4 cells x 4 images of the 64 x 64 tile
    const int tc = threadIdx.x & 15;   // cells 4*tc .. 4*tc+3
    const int tl = threadIdx.x >> 4;   // images 4*tl .. 4*tl+3
    // Two-level accumulation: `acc` collects kFlushChunks chunks (1,024 pixels), then moves into `tot`. One FP32 accumulator per
    // pair over a whole 256 px cell (65,536 terms of ~150 into a sum of 1e7, ulp 1) drifted by up to 3e-5 relative; in two
    // levels every addition happens at <= 1/64 of that magnitude and the error stays near 1e-6 (tests/test_gpu_ring_stress.py).
    constexpr int kFlushChunks = 64;
    mm_f2 acc[4][2], tot[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        acc[i][0] = acc[i][1] = tot[i][0] = tot[i][1] = mm_f2{0.0f, 0.0f};

    for (int k = 0; k < n_chunks; ++k) {
        const int s = k % kEStages;
        mbar_wait_e(&full_bar[s], (k / kEStages) & 1);
        if (MM_STRESS_SKEW)
            __nanosleep((unsigned)((warp * 97 + k * 29 + blockIdx.x * 7) % 400));
        const float4 *cs = reinterpret_cast<const float4 *>(smem + (size_t)s * kEStage) + tc;
        const float4 *ls = reinterpret_cast<const float4 *>(smem + (size_t)s * kEStage + kECellBlock) + tl;
#pragma unroll 2
        for (int p = 0; p < MM_EKP; ++p) {
            const float4 c0 = cs[(p * 4 + 0) * (MM_ETC / 4)], c1 = cs[(p * 4 + 1) * (MM_ETC / 4)], c2 = cs[(p * 4 + 2) * (MM_ETC / 4)],
                         cw = cs[(p * 4 + 3) * (MM_ETC / 4)];
            const float4 l0 = ls[(p * 3 + 0) * (MM_ETN / 4)], l1 = ls[(p * 3 + 1) * (MM_ETN / 4)], l2 = ls[(p * 3 + 2) * (MM_ETN / 4)];
            const mm_f2 C0[2] = {{c0.x, c0.y}, {c0.z, c0.w}}, C1[2] = {{c1.x, c1.y}, {c1.z, c1.w}}, C2[2] = {{c2.x, c2.y}, {c2.z, c2.w}},
                        W[2] = {{cw.x, cw.y}, {cw.z, cw.w}};
#pragma unroll
            for (int li = 0; li < 4; ++li) {
                const float n0 = comp(l0, li), n1 = comp(l1, li), n2 = comp(l2, li);  // negated library pixel
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) {
                    const mm_f2 d0 = v_add(C0[cp], mm_f2{n0, n0}), d1 = v_add(C1[cp], mm_f2{n1, n1}), d2 = v_add(C2[cp], mm_f2{n2, n2});
                    const mm_f2 ss = v_fma(d2, d2, v_fma(d1, d1, v_mul(d0, d0)));
                    acc[li][cp] = v_fma(W[cp], v_sqrt(ss), acc[li][cp]);
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive_e(&empty_bar[s]);
        if ((k % kFlushChunks) == kFlushChunks - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) {
                    tot[i][cp] = v_add(tot[i][cp], acc[i][cp]);
                    acc[i][cp] = mm_f2{0.0f, 0.0f};
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int cp = 0; cp < 2; ++cp)
            acc[i][cp] = v_add(tot[i][cp], acc[i][cp]);

    // epilogue: every thread owns complete sums of its 4 x 4 block
    const int li0 = lib_tile * MM_ETN + tl * 4;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        const int cell = cell_tile * MM_ETC + tc * 4 + cc;
        float v[4];
#pragma unroll
        for (int li = 0; li < 4; ++li)
            v[li] = (cc & 1) ? acc[li][cc >> 1].y : acc[li][cc >> 1].x;
        if (D)
            *reinterpret_cast<float4 *>(D + (size_t)cell * n_lib_pad + li0) = make_float4(v[0], v[1], v[2], v[3]);
        if (best_key && cell < n_cells) {
            unsigned long long key = ~0ull;
#pragma unroll
            for (int li = 0; li < 4; ++li)
                if (li0 + li < n_lib) {
                    const unsigned long long k2 = ((unsigned long long)__float_as_uint(v[li]) << 32) | (unsigned)(li0 + li);
                    key = k2 < key ? k2 : key;
                }
            if (key != ~0ull)
                atomicMin(best_key + cell, key);
        }
    }
    if (progress && threadIdx.x == 0)
        atomicAdd(progress, 1ull);
}

cudaError_t launch_diff_euclid(const void *cells, const void *lib, float *D, unsigned long long *best_key, int n_cell_tiles,
                               int n_lib_tiles, int n_chunks, int n_lib, int n_cells, cudaStream_t stream, const int *cancel,
                               unsigned long long *progress)
{
    if (n_cell_tiles <= 0 || n_lib_tiles <= 0 || n_chunks <= 0)
        return cudaSuccess;
    const size_t smem = (size_t)kEStages * kEStage;
    cudaError_t e = cudaFuncSetAttribute(diff_euclid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    const unsigned grid = (unsigned)n_cell_tiles * (unsigned)n_lib_tiles;  // 1-D, super-block rasterisation (kernels.h)
    diff_euclid_kernel<<<grid, kEThreads, smem, stream>>>((const unsigned char *)cells, (const unsigned char *)lib, D, best_key, n_chunks,
                                                          n_lib, n_lib_tiles * MM_ETN, n_cells, n_cell_tiles, n_lib_tiles, cancel, progress);
    return cudaGetLastError();
}

}  // namespace mm
