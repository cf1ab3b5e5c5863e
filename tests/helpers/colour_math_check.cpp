/* Test-only CPU build of mosaicmagnifique_b200/csrc/colour_math.cuh (the FP32 formulas the CUDA
 * kernels run), so the algebra can be checked against the f64 oracle without a GPU.
 * Built with g++ by tests/test_colour_math.py into tests/helpers/_build/. Never part of the product. */
#include "../../mosaicmagnifique_b200/csrc/colour_math.cuh"

extern "C" {
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

/* the packed (two library pixels per call) form must reproduce the scalar form bit for bit */
void cm_ciede2000_batch_x2(const float *a, const float *b, long n, float *out)
{
    for (long i = 0; i + 1 < n; i += 2) {
        const float *p = a + 3 * i, *q0 = b + 3 * i, *q1 = b + 3 * (i + 1);
        const float c1 = (float)sqrt((double)p[1] * p[1] + (double)p[2] * p[2]);
        const float c20 = (float)sqrt((double)q0[1] * q0[1] + (double)q0[2] * q0[2]);
        const float c21 = (float)sqrt((double)q1[1] * q1[1] + (double)q1[2] * q1[2]);
        const float sc = MM_CIEDE_AB_SCALE;
        const mm_f2 L2{fmaf(0.5f, q0[0], -25.0f), fmaf(0.5f, q1[0], -25.0f)}, a2{sc * q0[1], sc * q1[1]}, b2{sc * q0[2], sc * q1[2]},
            C2{sc * c20, sc * c21};
        const mm_f2 r = mm_ciede2000_stored_v<mm_f2>(fmaf(0.5f, p[0], -25.0f), sc * p[1], sc * p[2], sc * c1, L2, a2, b2, C2);
        out[i] = MM_CIEDE_WEIGHT * r.x;
        out[i + 1] = MM_CIEDE_WEIGHT * r.y;
    }
}
}
