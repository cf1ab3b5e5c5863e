// Pipe-rate micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not carry
// ("SM count / MUFU width must be confirmed by a micro-benchmark"):
//   FP32 FFMA lane-ops/s, packed FFMA2 (fma.rn.f32x2) lane-ops/s, MUFU rsq and ex2 ops/s, and the
//   SM clock that was actually sustained while they ran.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mm {

constexpr int kIters = 4096;
constexpr int kChains = 16;

__global__ void __launch_bounds__(256) ffma_kernel(float *out, float a, float b, long long *clk)
{
    float x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i)
        x[i] = (float)(threadIdx.x + i);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i)
            x[i] = fmaf(x[i], a, b);
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i)
        s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *clk = t1 - t0;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float *out, float a, float b, long long *clk)
{
    unsigned long long x[kChains / 2];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) {
        const float lo = (float)(threadIdx.x + i), hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains / 2; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains / 2; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *clk = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(256) mufu_kernel(float *out, long long *clk)
{
    float x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i)
        x[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (OP == 0)
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            else
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i)
        s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *clk = t1 - t0;
}

// mixed loop in the CIEDE2000 kernel's proportions: per "pixel pair" 40 packed FP32 ops (79 lane-ops + rounding) + 9 MUFU + 5 ALU ops
template <int N_F2, int N_MUFU, int N_ALU>
__global__ void __launch_bounds__(256) mixed_kernel(float *out, float a, float b, long long *clk)
{
    unsigned long long x[8];
    float m[4];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float lo = (float)(threadIdx.x + i), hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        m[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
    float sel = b;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters / 4; ++it) {
#pragma unroll
        for (int j = 0; j < N_F2; ++j) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[j % 8]) : "l"(av), "l"(bv));
            if (j * N_MUFU / N_F2 != (j + 1) * N_MUFU / N_F2)
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[j % 4]));
            if (j * N_ALU / N_F2 != (j + 1) * N_ALU / N_F2)
                asm volatile("max.f32 %0, %0, %1;" : "+f"(sel) : "f"(m[(j + 1) % 4]));
        }
    }
    const long long t1 = clock64();
    float s = sel;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *clk = t1 - t0;
}

// packed and scalar FP32 in one loop: N2 fma.rn.f32x2 + N1 scalar fma per iteration on independent chains. If scalar FFMA could
// issue to a second FP32 pipe while FFMA2 holds the first, the combined lane-op rate would exceed either pure rate.
template <int N2, int N1>
__global__ void __launch_bounds__(256) mix_f2_f1_kernel(float *out, float a, float b, long long *clk)
{
    unsigned long long x[N2];
    float y[N1];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < N2; ++i) {
        const float lo = (float)(threadIdx.x + i), hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
#pragma unroll
    for (int i = 0; i < N1; ++i)
        y[i] = (float)(threadIdx.x + 3 * i);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < (N2 > N1 ? N2 : N1); ++i) {
            if (i < N2)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
            if (i < N1)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(a), "f"(b));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < N2; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi;
    }
#pragma unroll
    for (int i = 0; i < N1; ++i)
        s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *clk = t1 - t0;
}

cudaError_t run_microbench(double *out, int n_out, cudaStream_t stream)
{
    if (n_out < 6)
        return cudaErrorInvalidValue;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    float *buf = nullptr;
    long long *clk = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)blocks * threads * sizeof(float));
    if (e != cudaSuccess)
        return e;
    e = cudaMalloc(&clk, sizeof(long long));
    if (e != cudaSuccess) {
        cudaFree(buf);
        return e;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double lane_ops = (double)blocks * threads * kIters * kChains;
    for (int which = 0; which < 7; ++which) {
        if (which >= 4 && n_out < 6 + which - 3)
            break;
        float best_ms = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0, stream);
            switch (which) {
            case 0: ffma_kernel<<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk); break;
            case 1: ffma2_kernel<<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk); break;
            case 2: mufu_kernel<0><<<blocks, threads, 0, stream>>>(buf, clk); break;
            case 3: mufu_kernel<1><<<blocks, threads, 0, stream>>>(buf, clk); break;
            case 4: mixed_kernel<40, 9, 5><<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk); break;
            case 5: mixed_kernel<40, 9, 0><<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk); break;
            default: mixed_kernel<40, 0, 0><<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk); break;
            }
            cudaEventRecord(e1, stream);
            e = cudaEventSynchronize(e1);
            if (e != cudaSuccess)
                goto done;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms)
                best_ms = ms;
        }
        if (which >= 4) {
            // "pixel pairs" per second: one inner iteration of mixed_kernel per thread = one pair
            out[6 + which - 4] = (double)blocks * threads * (kIters / 4) / (best_ms * 1e-3);
            continue;
        }
        out[which] = lane_ops / (best_ms * 1e-3);
        if (which == 0) {
            // CTA 0 of an 8-waves-per-SM launch: its loop cycles / (kernel time / 8 waves ... ) is not exact;
            // report cycles per iteration instead and let the caller combine with nvidia-smi clocks
            long long c = 0;
            cudaMemcpyAsync(&c, clk, sizeof c, cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            out[5] = (double)c / kIters;  // cycles per loop iteration (kChains FFMA per thread, 8 warps/CTA resident mix)
        }
    }
    out[4] = (double)sms;
    // [9], [10]: FP32 lane-ops/s of 8 FFMA2 + 8 FFMA and of 8 FFMA2 + 16 FFMA per iteration
    for (int which = 0; which < 2 && n_out >= 11; ++which) {
        float best_ms = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0, stream);
            if (which == 0)
                mix_f2_f1_kernel<8, 8><<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk);
            else
                mix_f2_f1_kernel<8, 16><<<blocks, threads, 0, stream>>>(buf, 1.0001f, 0.5f, clk);
            cudaEventRecord(e1, stream);
            e = cudaEventSynchronize(e1);
            if (e != cudaSuccess)
                goto done;
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms)
                best_ms = ms;
        }
        out[9 + which] = (double)blocks * threads * kIters * (which == 0 ? 24.0 : 32.0) / (best_ms * 1e-3);
    }

done:
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(clk);
    return e;
}

}  // namespace mm
