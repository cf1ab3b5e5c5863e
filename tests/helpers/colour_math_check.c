/* Test-only CPU build of mosaicmagnifique_b200/csrc/colour_math.cuh (the FP32 formulas the CUDA
 * kernels run), so the algebra can be checked against the f64 oracle without a GPU.
 * Built by tests/test_colour_math.py into tests/helpers/_build/. Never part of the product. */
#include "../../mosaicmagnifique_b200/csrc/colour_math.cuh"

void cm_euclid_batch(const float *a, const float *b, long n, float *out)
{
    for (long i = 0; i < n; ++i)
        out[i] = mm_euclid(a[3 * i], a[3 * i + 1], a[3 * i + 2], b[3 * i], b[3 * i + 1], b[3 * i + 2]);
}

void cm_ciede2000_batch(const float *a, const float *b, long n, float *out)
{
    for (long i = 0; i < n; ++i) {
        const float *p = a + 3 * i, *q = b + 3 * i;
        /* chroma as the prep kernels compute it: f64 sqrt of the f32 values, rounded to f32 */
        const float c1 = (float)sqrt((double)p[1] * p[1] + (double)p[2] * p[2]);
        const float c2 = (float)sqrt((double)q[1] * q[1] + (double)q[2] * q[2]);
        out[i] = mm_ciede2000(p[0], p[1], p[2], c1, q[0], q[1], q[2], c2);
    }
}
