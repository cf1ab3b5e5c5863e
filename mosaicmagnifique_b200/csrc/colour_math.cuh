// Per-pixel colour differences in FP32 for the fused difference-sum kernels.
//
// Replaces the reference's per-pixel device functions
//   euclideanDifference   src/Photomosaic/CUDA/ColourDifference.cuh:10-15
//   CIEDE2000Difference   src/Photomosaic/CUDA/ColourDifference.cuh:24-139
// (CPU originals: src/Photomosaic/ColourDifference.cpp:28-33, 42-158).
//
// The CIEDE2000 here is NOT a transcription of the reference's formula: it is an algebraically
// equivalent trig-free form built for the B200's pipe balance (MUFU 16 lanes/clk/SM vs FP32 128).
// The reference evaluates 2 atan2, 4 cos, 2 sin, 1 exp, 9 sqrt and 9 divisions (27 special-function
// ops) per pixel pair in f64; this form needs 11 MUFU ops (rsq/rcp/ex2/sqrt) and ~125 FP32 ops:
//   * dH' = 2 sqrt(C1'C2') sin(dh/2) is computed from the dot and cross products of the (a', b)
//     vectors:  dH' = sqrt2 * cross / sqrt(C1'C2' + dot)   (dot > 0, well conditioned for close hues)
//                   = sign(cross) * sqrt(2 (C1'C2' - dot))  (dot <= 0),
//     which also carries the sign the R_T cross term needs.
//   * the mean hue enters T only through cos/sin of 1..4 times the angle, so the unit bisector
//     (C2' v1 + C1' v2)/|..| and angle-addition recurrences replace atan2 + 4 cos. The reference's
//     special cases (h' = 0 for achromatic colours, mean = sum when C1'C2' = 0, the +-pi folds)
//     are exactly the statement "mean hue = direction of the sum of the unit vectors".
//   * dTheta = 30deg exp(-((hbar-275deg)/25deg)^2) needs the angle itself: rotate the bisector by
//     -275deg, half-angle tangent, odd minimax atan polynomial. For |hbar-275deg| > 90deg the
//     Gaussian is < 2.4e-6 and is dropped (changes dE by < 1e-5 relative; documented in DESIGN.md).
//   * x/(x+k) square roots are x*rsq(x*(x+k)), the three final divisions share one rcp.
// Accuracy against the f64 reference formula: see tests/test_colour_math.py (max relative error of
// the per-pixel value ~1e-5 on random Lab pairs, golden Sharma vectors within the reference's 1e-4).
//
// The functions are __host__ __device__ so the same source is unit-tested on the CPU
// (tests build a tiny checker from this header); the product only ever runs the device code.
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
#else
#define MM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
// approximate MUFU forms: one XU op each (rsq.approx, rcp.approx, sqrt.approx, ex2.approx)
MM_HD float mm_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MM_HD float mm_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MM_HD float mm_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
MM_HD float mm_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
MM_HD float mm_rsq(float x) { return 1.0f / sqrtf(x); }
MM_HD float mm_rcp(float x) { return 1.0f / x; }
MM_HD float mm_sqrt(float x) { return sqrtf(x); }
MM_HD float mm_ex2(float x) { return exp2f(x); }
#endif

// Euclidean distance of two 3-vectors (RGB Euclidean on BGR 0..255 floats, CIE76 on Lab;
// ColourDifference.cpp:28-33, ColourDifference.h:41).
MM_HD float mm_euclid(float x0, float x1, float x2, float y0, float y1, float y2)
{
    const float d0 = x0 - y0, d1 = x1 - y1, d2 = x2 - y2;
    return mm_sqrt(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
}

// CIEDE2000 (kL = kC = kH = 1). (L, a, b, C) per colour, C = sqrt(a*a + b*b) precomputed per pixel
// (ColourDifference.cpp:48-49 hoisted out of the pair loop).
MM_HD float mm_ciede2000(float L1, float a1, float b1, float C1, float L2, float a2, float b2, float C2)
{
    const float k25_7 = 6103515625.0f;  // 25^7, ColourDifference.cpp:46
    const float tiny = 1e-30f;

    // G and a' (:50-57): sqrt(x/(x+k)) = x * rsq(x*(x+k))
    const float cbar = 0.5f * (C1 + C2);
    const float cb2 = cbar * cbar;
    const float cb4 = cb2 * cb2;
    const float cb7 = cb4 * cb2 * cbar;
    const float qg = cb7 * mm_rsq(fmaf(cb7, cb7 + k25_7, tiny));
    const float g1 = fmaf(-0.5f, qg, 1.5f);  // 1 + G
    const float a1p = g1 * a1, a2p = g1 * a2;

    // C' (:58-59)
    const float c1p = mm_sqrt(fmaf(a1p, a1p, b1 * b1));
    const float c2p = mm_sqrt(fmaf(a2p, a2p, b2 * b2));

    // deltas (:87-102)
    const float dL = L2 - L1;
    const float dC = c2p - c1p;
    const float P = c1p * c2p;
    const float dot = fmaf(a1p, a2p, b1 * b2);
    const float cross = fmaf(a1p, b2, -(a2p * b1));
    const bool pos = dot > 0.0f;
    const float argA = P + dot;
    const float argB = 2.0f * (P - dot);
    const float arg = fmaxf(pos ? argA : argB, tiny);
    const float rs = mm_rsq(arg);
    const float dHa = 1.41421356237f * cross * rs;
    const float dHb = copysignf(arg * rs, cross);
    float dH = pos ? dHa : dHb;
    dH = (P == 0.0f) ? 0.0f : dH;  // dh' = 0 when C1'C2' == 0 (:90)

    // unit vector of the mean hue (:104-122)
    float wx = fmaf(c2p, a1p, c1p * a2p);
    float wy = fmaf(c2p, b1, c1p * b2);
    if (P == 0.0f) { wx = a1p + a2p; wy = b1 + b2; }  // mean = h1' + h2' with one of them 0 (:109-110)
    if (argA < 1e-4f * P) {
        // hues within ~0.8deg of opposite: the sum of the unit vectors cancels. Its direction is also
        // perp(C2' v1 - C1' v2) signed by the cross product, which stays well conditioned. (Exactly at
        // 180deg the reference's own mean hue flips with the last bit of atan2, :111-121.)
        const float dx = fmaf(c2p, a1p, -(c1p * a2p));
        const float dy = fmaf(c2p, b1, -(c1p * b2));
        wx = copysignf(1.0f, cross) * -dy;
        wy = copysignf(1.0f, cross) * dx;
    }
    const float n2 = fmaf(wx, wx, wy * wy);
    const float rn = mm_rsq(fmaxf(n2, tiny));
    const float ch = wx * rn, sh = wy * rn;

    // T (:124-127) through angle-addition recurrences
    const float ch2x = ch + ch;
    const float c2 = fmaf(ch2x, ch, -1.0f);
    const float s2 = ch2x * sh;
    const float c3 = fmaf(ch, c2, -(sh * s2));
    const float s3 = fmaf(sh, c2, ch * s2);
    const float c2x2 = c2 + c2;
    const float c4 = fmaf(c2x2, c2, -1.0f);
    const float s4 = c2x2 * s2;
    float T = 1.0f;
    T = fmaf(-0.17f * 0.86602540378f, ch, T);   // -0.17 cos(h - 30)
    T = fmaf(-0.17f * 0.5f, sh, T);
    T = fmaf(0.24f, c2, T);                     // +0.24 cos(2h)
    T = fmaf(0.32f * 0.99452189536f, c3, T);    // +0.32 cos(3h + 6)
    T = fmaf(-0.32f * 0.10452846326f, s3, T);
    T = fmaf(-0.20f * 0.45399049974f, c4, T);   // -0.20 cos(4h - 63)
    T = fmaf(-0.20f * 0.89100652418f, s4, T);

    // dTheta (:129-134): angle phi = hbar - 275deg from the rotated unit vector
    const float cphi = fmaf(ch, 0.08715574275f, -(sh * 0.99619469809f));   // cos275 = 0.0871557, sin275 = -0.9961947
    const float sphi = fmaf(sh, 0.08715574275f, ch * 0.99619469809f);
    const float th = sphi * mm_rcp(fmaxf(1.0f + cphi, 0.5f));               // tan(phi/2), |.| <= 1 when cphi >= 0
    const float th2 = th * th;
    // atan(x)/x on [-1,1] as a degree-6 polynomial in x^2 (Lawson/minimax fit, |err| < 7e-7)
    float pa = 0.008249403913f;
    pa = fmaf(pa, th2, -0.03821812122f);
    pa = fmaf(pa, th2, 0.08530285112f);
    pa = fmaf(pa, th2, -0.1356754763f);
    pa = fmaf(pa, th2, 0.1990285901f);
    pa = fmaf(pa, th2, -0.3332884304f);
    pa = fmaf(pa, th2, 1.0f);
    const float half_phi = th * pa;
    // exp(-(phi/25deg)^2) = 2^(-(4 log2(e) / (25deg)^2) * half_phi^2)
    const float kexp = -4.0f * 1.44269504089f / (0.43633231299f * 0.43633231299f);
    float gauss = mm_ex2(kexp * half_phi * half_phi);
    gauss = (cphi < 0.0f) ? 0.0f : gauss;
    // sin(2 dTheta), 2 dTheta = (pi/3) * gauss in [0, 1.0472]
    const float xs = 1.0471975512f * gauss;
    const float xs2 = xs * xs;
    float ps = 2.7557319e-6f;
    ps = fmaf(ps, xs2, -1.9841270e-4f);
    ps = fmaf(ps, xs2, 8.3333333e-3f);
    ps = fmaf(ps, xs2, -1.6666667e-1f);
    ps = fmaf(ps, xs2, 1.0f);
    const float sin2dt = xs * ps;

    // R_C, R_T (:135-136, 146)
    const float cpbar = 0.5f * (c1p + c2p);
    const float cp2 = cpbar * cpbar;
    const float cp4 = cp2 * cp2;
    const float cp7 = cp4 * cp2 * cpbar;
    const float rc = 2.0f * cp7 * mm_rsq(fmaf(cp7, cp7 + k25_7, tiny));
    const float rt = -sin2dt * rc;

    // S_L, S_C, S_H (:138-144)
    const float lm = fmaf(0.5f, L1 + L2, -50.0f);
    const float ql = lm * lm;
    const float sl = fmaf(0.015f * ql, mm_rsq(20.0f + ql), 1.0f);
    const float sc = fmaf(0.045f, cpbar, 1.0f);
    const float shh = fmaf(0.015f * cpbar, T, 1.0f);

    // dE (:153-157), the three divisions share one reciprocal
    const float scsh = sc * shh;
    const float inv = mm_rcp(sl * scsh);
    const float x = dL * (inv * scsh);
    const float y = dC * (inv * (sl * shh));
    const float z = dH * (inv * (sl * sc));
    const float s = fmaf(z, z, fmaf(y, fmaf(rt, z, y), x * x));
    return mm_sqrt(fmaxf(s, 0.0f));
}
