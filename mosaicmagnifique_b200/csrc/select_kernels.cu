// Selection stage: repeat-penalised argmin in the reference's raster order, as a GPU wavefront.
//
// Replaces the reference's serial <<<1,1>>> kernels
//   calculateRepeats  src/Photomosaic/CUDA/PhotomosaicGenerator.cu:127-172
//   findLowestKernel  src/Photomosaic/CUDA/PhotomosaicGenerator.cu:175-196
// with the CPU generator's semantics (CPUPhotomosaicGenerator.cpp:137-169, 185-225):
//   best(x,y) = argmin_i ( A * #{already chosen cells in the window with value i} + D[cell, i] ),
//   strict <, lowest index wins; window = rows y-r..y-1 x columns x-r..x+r, plus row y columns x-r..x-1.
//
// Wavefront: valid cells are dealt to the CTAs of one co-resident (cooperative) grid in raster order.
// A cell may start once every valid cell left of it in its row is final and rows y-r..y-1 are final up
// to column x+r; rows publish "final up to column" counters (release/acquire through L2). All
// dependencies point backwards in raster order, so the earliest unfinished cell can always run.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "kernels.h"

namespace mm {

__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

cudaError_t launch_fill_u64(unsigned long long *p, size_t n, unsigned long long v, cudaStream_t stream)
{
    if (n == 0)
        return cudaSuccess;
    fill_u64_kernel<<<(unsigned)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0, stream>>>(p, n, v);
    return cudaGetLastError();
}

// rows are laid out cell-major: row(cell, v) = cell * V + v; the minimum lands in row(cell, 0)
__global__ void min_variants_kernel(float *D, int n_cells, int V, int n_lib_pad)
{
    const size_t total = (size_t)n_cells * n_lib_pad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / n_lib_pad, l = i - c * n_lib_pad;
        float *row = D + c * V * (size_t)n_lib_pad + l;
        float m = row[0];
        for (int v = 1; v < V; ++v)
            m = fminf(m, row[(size_t)v * n_lib_pad]);
        row[0] = m;
    }
}

cudaError_t launch_min_variants(float *D, int n_cells, int V, int n_lib_pad, cudaStream_t stream)
{
    const size_t total = (size_t)n_cells * n_lib_pad;
    if (total == 0 || V <= 1)
        return cudaSuccess;
    min_variants_kernel<<<(unsigned)((total + 255) / 256 > 4736 ? 4736 : (total + 255) / 256), 256, 0, stream>>>(D, n_cells, V,
                                                                                                                  n_lib_pad);
    return cudaGetLastError();
}

__global__ void keys_to_grid_kernel(const unsigned long long *best_key, const int *cell_pos, long long *grid, int n_cells,
                                    float *best_score)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells)
        return;
    const unsigned long long k = best_key[c];
    // an untouched key means no library image produced a finite sum: the reference's "should never
    // happen" nullopt (CPUPhotomosaicGenerator.cpp:174-178)
    grid[cell_pos[c]] = (k == ~0ull) ? -1ll : (long long)(k & 0xffffffffull);
    if (best_score)
        best_score[c] = __uint_as_float((unsigned)(k >> 32));
}

cudaError_t launch_keys_to_grid(const unsigned long long *best_key, const int *cell_pos, long long *grid, int n_cells,
                                float *best_score, cudaStream_t stream)
{
    if (n_cells == 0)
        return cudaSuccess;
    keys_to_grid_kernel<<<(n_cells + 255) / 256, 256, 0, stream>>>(best_key, cell_pos, grid, n_cells, best_score);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ top-K candidates per cell

constexpr int kTopkThreads = 256;

// K smallest entries of one row by (score bits, index): 4-pass radix select for the K-th value, then an
// index-ordered compaction so that ties at the threshold keep the lowest indices (the CPU's tie rule).
__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const float *__restrict__ D, int row_stride, int n_lib, int K, float *__restrict__ cand_score,
            int *__restrict__ cand_idx)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining, s_base;
    __shared__ unsigned scan[kTopkThreads];
    const float *row = D + (size_t)blockIdx.x * row_stride;
    const int tid = threadIdx.x;

    unsigned prefix = 0, remaining = (unsigned)K;  // looking for the remaining-th smallest among keys matching prefix
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const unsigned mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < n_lib; i += kTopkThreads) {
            const unsigned b = __float_as_uint(row[i]);
            if ((b & mask_hi) == prefix)
                atomicAdd(&hist[(b >> shift) & 255], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned acc = 0, d = 0;
            for (; d < 256; ++d) {
                if (acc + hist[d] >= remaining)
                    break;
                acc += hist[d];
            }
            s_prefix = prefix | (d << shift);
            s_remaining = remaining - acc;
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        __syncthreads();
    }
    // prefix = bits of the K-th smallest value; `remaining` of the entries equal to it are taken (lowest indices)
    const unsigned thr = prefix;
    if (tid == 0)
        s_base = 0;
    unsigned eq_taken = 0;  // uniform across the block: equal-to-threshold entries accepted so far
    float *out_s = cand_score + (size_t)blockIdx.x * K;
    int *out_i = cand_idx + (size_t)blockIdx.x * K;
    __syncthreads();
    for (int base = 0; base < n_lib; base += kTopkThreads) {
        const int i = base + tid;
        unsigned b = 0xffffffffu;
        if (i < n_lib)
            b = __float_as_uint(row[i]);
        const bool less = i < n_lib && b < thr;
        const bool eq = i < n_lib && b == thr;
        // inclusive scans of both predicates (packed: low 16 bits = less, high 16 bits = eq)
        unsigned v = (less ? 1u : 0u) | (eq ? 0x10000u : 0u);
        scan[tid] = v;
        __syncthreads();
        for (int o = 1; o < kTopkThreads; o <<= 1) {
            const unsigned t = tid >= o ? scan[tid - o] : 0u;
            __syncthreads();
            scan[tid] += t;
            __syncthreads();
        }
        const unsigned incl = scan[tid];
        const unsigned total = scan[kTopkThreads - 1];
        const unsigned eq_before = eq_taken + (incl >> 16) - (eq ? 1u : 0u);
        const bool take_eq = eq && eq_before < remaining;
        // number of accepted entries before this one inside the chunk
        const unsigned less_before = (incl & 0xffffu) - (less ? 1u : 0u);
        const unsigned eq_acc_before = min(eq_before, remaining) - min(eq_taken, remaining);
        if (less || take_eq) {
            const unsigned pos = s_base + less_before + eq_acc_before;
            out_s[pos] = __uint_as_float(b);
            out_i[pos] = i;
        }
        __syncthreads();
        const unsigned eq_total = total >> 16;
        const unsigned eq_acc_total = min(eq_taken + eq_total, remaining) - min(eq_taken, remaining);
        if (tid == 0)
            s_base += (total & 0xffffu) + eq_acc_total;
        eq_taken += eq_total;
        __syncthreads();
    }
}

cudaError_t launch_topk(const float *D, int row_stride, int n_lib, int n_cells, int K, float *cand_score, int *cand_idx,
                        cudaStream_t stream)
{
    if (n_cells == 0 || K <= 0)
        return cudaSuccess;
    if (K > n_lib)
        return cudaErrorInvalidValue;
    topk_kernel<<<n_cells, kTopkThreads, 0, stream>>>(D, row_stride, n_lib, K, cand_score, cand_idx);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ wavefront selection

constexpr int kSelThreads = 256;

struct SelArgs {
    long long *grid;
    const int *cell_pos;   // [n_cells] y * cols + x of each valid cell, raster order
    const int *next_x;     // [n_cells] column of the next valid cell in the same row (cols if none)
    int n_cells, rows, cols;
    const float *scores;   // per cell M entries, row stride M_stride
    const int *idx;        // per cell M entries (stride M) or nullptr: entry j is library image j
    int M, M_stride, n_lib;
    int range, addition;
    int *row_progress;     // [rows], initialised to the column of the first valid cell (cols if none)
    int *counts;           // [gridDim.x][n_lib] zeroed scratch: occurrences of each library image in the window
    float *margins;        // optional [n_cells][2]: best and second-best PENALISED score (tie-band reporting)
    // all-gathered candidate blocks (multi-GPU): cell c lives in block c / rows_per_block at row c % rows_per_block; blocks are
    // block_stride 4-byte elements apart (scores and indices alike). rows_per_block == 0: one plain array.
    long long rows_per_block, block_stride;
};

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(SelArgs a)
{
    __shared__ double s_val[kSelThreads / 32];
    __shared__ int s_id[kSelThreads / 32];
    __shared__ double s_second[kSelThreads / 32];
    int *cnt = a.counts + (size_t)blockIdx.x * a.n_lib;
    const int tid = threadIdx.x;

    for (int c = blockIdx.x; c < a.n_cells; c += gridDim.x) {
        const int pos = a.cell_pos[c];
        const int y = pos / a.cols, x = pos - y * a.cols;
        // clamped window (CPUPhotomosaicGenerator.cpp:188-192)
        const int y0 = min(max(y - a.range, 0), a.rows);
        const int x0 = min(max(x - a.range, 0), a.cols);
        const int x1 = min(max(x + a.range, 0), a.cols - 1);
        const bool penalise = a.range > 0 && a.addition != 0;

        // ---- wait for the dependencies
        if (penalise) {
            // row y final up to x (exclusive); rows y0..y-1 final up to x1 (inclusive)
            for (int ry = y0 + tid; ry <= y; ry += kSelThreads) {
                const int need = ry == y ? x : x1 + 1;
                while (ld_acquire(a.row_progress + ry) < need)
                    __nanosleep(64);
            }
        }
        __syncthreads();

        // ---- count the window's library images
        const int ww = x1 - x0 + 1;
        const int n_above = (y - y0) * ww;
        const int n_win = penalise ? n_above + (x - x0) : 0;
        for (int j = tid; j < n_win; j += kSelThreads) {
            int ry, rx;
            if (j < n_above) {
                ry = y0 + j / ww;
                rx = x0 + j % ww;
            } else {
                ry = y;
                rx = x0 + (j - n_above);
            }
            const long long v = __ldcg(a.grid + (size_t)ry * a.cols + rx);
            if (v >= 0)
                atomicAdd(cnt + v, 1);
        }
        __syncthreads();

        // ---- penalised argmin over this cell's entries
        size_t row = (size_t)c, base = 0;
        if (a.rows_per_block > 0) {
            base = (size_t)(c / a.rows_per_block) * (size_t)a.block_stride;
            row = (size_t)(c % a.rows_per_block);
        }
        const float *sc = a.scores + base + row * a.M_stride;
        const int *ids = a.idx ? a.idx + base + row * a.M : nullptr;
        double best = DBL_MAX, second = DBL_MAX;
        int best_id = 0x7fffffff;
        for (int j = tid; j < a.M; j += kSelThreads) {
            const int id = ids ? ids[j] : j;
            const float s = sc[j];
            const int k = penalise ? __ldcg(cnt + id) : 0;
            const double v = (double)s + (double)a.addition * (double)k;
            if (v < best || (v == best && id < best_id)) {
                second = best;
                best = v;
                best_id = id;
            } else if (v < second) {
                second = v;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
            const double os = __shfl_xor_sync(0xffffffffu, second, o);
            if (ov < best || (ov == best && oi < best_id)) {
                second = fmin(best, os);
                best = ov;
                best_id = oi;
            } else {
                second = fmin(second, ov);
            }
        }
        if ((tid & 31) == 0) {
            s_val[tid >> 5] = best;
            s_id[tid >> 5] = best_id;
            s_second[tid >> 5] = second;
        }
        __syncthreads();

        // ---- undo the counts (every thread revisits its own window entries)
        for (int j = tid; j < n_win; j += kSelThreads) {
            int ry, rx;
            if (j < n_above) {
                ry = y0 + j / ww;
                rx = x0 + j % ww;
            } else {
                ry = y;
                rx = x0 + (j - n_above);
            }
            const long long v = __ldcg(a.grid + (size_t)ry * a.cols + rx);
            if (v >= 0)
                cnt[v] = 0;
        }

        if (tid == 0) {
            for (int w = 1; w < kSelThreads / 32; ++w)
                if (s_val[w] < best || (s_val[w] == best && s_id[w] < best_id)) {
                    second = fmin(best, s_second[w]);
                    best = s_val[w];
                    best_id = s_id[w];
                } else {
                    second = fmin(second, s_val[w]);
                }
            // DBL_MAX start + strict < as in the reference: NaN / inf rows leave the cell unset (nullopt)
            const long long result = (best < DBL_MAX && best_id != 0x7fffffff) ? (long long)best_id : -1ll;
            __stcg(a.grid + pos, result);
            if (a.margins) {
                a.margins[2 * c] = (float)best;
                a.margins[2 * c + 1] = (float)second;
            }
        }
        __syncthreads();  // count resets and the result are complete before the row counter moves
        if (tid == 0) {
            __threadfence();
            st_release(a.row_progress + y, a.next_x[c]);
        }
    }
}

cudaError_t launch_select(long long *grid, const int *cell_pos, const int *next_x, int n_cells, int rows, int cols,
                          const float *scores, const int *idx, int M, int M_stride, int n_lib, int repeat_range,
                          int repeat_addition, int *row_progress, int *counts, int n_ctas, float *margins,
                          cudaStream_t stream, long long rows_per_block, long long block_stride)
{
    if (n_cells == 0)
        return cudaSuccess;
    SelArgs a{grid, cell_pos, next_x, n_cells, rows, cols, scores, idx, M, M_stride, n_lib, repeat_range, repeat_addition,
              row_progress, counts, margins, rows_per_block, block_stride};
    void *args[] = {&a};
    // cooperative launch: fails instead of deadlocking if the CTAs could not all be resident
    return cudaLaunchCooperativeKernel((void *)select_kernel, dim3(n_ctas), dim3(kSelThreads), args, 0, stream);
}

int select_max_ctas(int device)
{
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_kernel, kSelThreads, 0);
    return sms * (per_sm > 0 ? 1 : 0);
}

}  // namespace mm
