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
#error "colour_math.cuh is C++ (the CPU test harness builds it with g++)"
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
// CPU emulation for the unit tests (tests/helpers/colour_math_check.cpp); the product never runs this
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

// ---------------------------------------------------------------------------------------------------
// Lane vectors: the CIEDE2000 body below is written once, for V = float (one pixel pair) and V = mm_f2 (two pixel
// pairs: the same cell pixel against two library images). On sm_100a the mm_f2 arithmetic is the packed FP32 pipe
// (fma/add/mul.rn.f32x2 -> FFMA2 / FADD2 / FMUL2 in SASS): one issue slot for two lanes, which matters because the
// scalar kernel is issue-bound (profiles/r1_diff_sum_ciede2000_v2.txt). MUFU ops and selects stay per lane.
// ---------------------------------------------------------------------------------------------------
struct mm_f2 {
    float x, y;
};
struct mm_b2 {
    bool x, y;
};

MM_HD float v_splat(float a, float) { return a; }
MM_HD mm_f2 v_splat(float a, mm_f2) { return mm_f2{a, a}; }

// ---- scalar lane ops
MM_HD float v_add(float a, float b) { return a + b; }
MM_HD float v_mul(float a, float b) { return a * b; }
MM_HD float v_fma(float a, float b, float c) { return fmaf(a, b, c); }
MM_HD float v_rsq(float a) { return mm_rsq(a); }
MM_HD float v_rcp(float a) { return mm_rcp(a); }
MM_HD float v_sqrt(float a) { return mm_sqrt(a); }
MM_HD float v_ex2(float a) { return mm_ex2(a); }
MM_HD bool v_gt0(float a) { return a > 0.0f; }
MM_HD bool v_eq(float a, float b) { return a == b; }
MM_HD float v_sel(bool m, float a, float b) { return m ? a : b; }
MM_HD float v_abs(float a) { return fabsf(a); }
MM_HD float v_add_abs(float a, float b) { return a + fabsf(b); }
MM_HD float v_copysign(float a, float s) { return copysignf(a, s); }
MM_HD float v_min(float a, float b) { return fminf(a, b); }
MM_HD float v_max(float a, float b) { return fmaxf(a, b); }

// a*b + c*d with each product rounded on its own. ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with
// --fmad=false), so the packed version falls back to scalar __fmul_rn / __fadd_rn, which are never fused.
#if defined(__CUDA_ARCH__)
MM_HD float v_sum_of_products_unfused(float a, float b, float c, float d) { return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
#else
MM_HD float v_sum_of_products_unfused(float a, float b, float c, float d)
{
    volatile float p = a * b, q = c * d;  // volatile: no contraction on the host either
    return p + q;
}
#endif

// ---- packed lane ops
#if defined(__CUDA_ARCH__)
MM_HD mm_f2 v_add(mm_f2 a, mm_f2 b) { const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return mm_f2{r.x, r.y}; }
MM_HD mm_f2 v_mul(mm_f2 a, mm_f2 b) { const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return mm_f2{r.x, r.y}; }
MM_HD mm_f2 v_fma(mm_f2 a, mm_f2 b, mm_f2 c)
{
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return mm_f2{r.x, r.y};
}
#else
MM_HD mm_f2 v_add(mm_f2 a, mm_f2 b) { return mm_f2{a.x + b.x, a.y + b.y}; }
MM_HD mm_f2 v_mul(mm_f2 a, mm_f2 b) { return mm_f2{a.x * b.x, a.y * b.y}; }
MM_HD mm_f2 v_fma(mm_f2 a, mm_f2 b, mm_f2 c) { return mm_f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif
MM_HD mm_f2 v_sum_of_products_unfused(mm_f2 a, mm_f2 b, mm_f2 c, mm_f2 d)
{
    return mm_f2{v_sum_of_products_unfused(a.x, b.x, c.x, d.x), v_sum_of_products_unfused(a.y, b.y, c.y, d.y)};
}
MM_HD mm_f2 v_rsq(mm_f2 a) { return mm_f2{mm_rsq(a.x), mm_rsq(a.y)}; }
MM_HD mm_f2 v_rcp(mm_f2 a) { return mm_f2{mm_rcp(a.x), mm_rcp(a.y)}; }
MM_HD mm_f2 v_sqrt(mm_f2 a) { return mm_f2{mm_sqrt(a.x), mm_sqrt(a.y)}; }
MM_HD mm_f2 v_ex2(mm_f2 a) { return mm_f2{mm_ex2(a.x), mm_ex2(a.y)}; }
MM_HD mm_b2 v_gt0(mm_f2 a) { return mm_b2{a.x > 0.0f, a.y > 0.0f}; }
MM_HD mm_b2 v_eq(mm_f2 a, mm_f2 b) { return mm_b2{a.x == b.x, a.y == b.y}; }
MM_HD mm_f2 v_sel(mm_b2 m, mm_f2 a, mm_f2 b) { return mm_f2{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
MM_HD mm_f2 v_abs(mm_f2 a) { return mm_f2{fabsf(a.x), fabsf(a.y)}; }
MM_HD mm_f2 v_add_abs(mm_f2 a, mm_f2 b) { return mm_f2{a.x + fabsf(b.x), a.y + fabsf(b.y)}; }
MM_HD mm_f2 v_copysign(mm_f2 a, mm_f2 s) { return mm_f2{copysignf(a.x, s.x), copysignf(a.y, s.y)}; }
MM_HD mm_f2 v_min(mm_f2 a, mm_f2 b) { return mm_f2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
MM_HD mm_f2 v_max(mm_f2 a, mm_f2 b) { return mm_f2{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }

// ---------------------------------------------------------------------------------------------------
// CIEDE2000 (kL = kC = kH = 1) on the STORED per-pixel channels, returning dE / MM_CIEDE_WEIGHT (= dE / 50).
//
// Stored per pixel (prep_kernels.cu):  Lh = L/2 - 25,  as = a/50,  bs = b/50,  Cs = sqrt(a*a + b*b)/50.
//   * Lh: the sum of two stored values is Lbar - 50 and their difference is dL'/2.
//   * (a, b) plane: everything there is homogeneous of degree one, so with 1/50-scale inputs the SUM of two chromas is
//     the mean chroma over 25 -- exactly the x of the two (x^7 / (x^7 + 25^7)) terms (G and R_C), which become
//     t/(t + 1) with t = sum^7 -- and dC', dH' come out as (dC'/2)/25, (dH'/2)/25. The 25 is folded into S_L
//     (sl25 = 25 S_L multiplies both), the constants of S_C and T, and the result's overall 1/25 (plus the 1/2 of the
//     half-scale deltas) into the per-pixel weight: w = MM_CIEDE_WEIGHT for an active pixel.
//
// Work per pixel pair: 9 MUFU + 79 FP32 lane-ops + 5 ALU ops (reference: 27 special-function ops + ~110 flop in f64). Besides the
// algebra described at the top of this file:
//   * T = P4(cos h) + sin h * Q3(cos h): the four cosines of T are Chebyshev polynomials of cos h and sin h times
//     Chebyshev-U, collected into one quartic and one cubic (8 FMA instead of recurrences);
//   * mean hue = hue(v1) + dh/2: e^(i dh/2) is (P + dot, cross) for dot > 0 and (|cross|, +-(P - dot)) otherwise
//     (both well conditioned), rotated by v1 and normalised once;
//   * dTheta needs only (hbar - 275deg)^2: v times a Gaussian-weighted cubic in v = 1 - cos(hbar - 275deg) (|error| of
//     R_T < 1.1e-6; for |hbar - 275deg| >= 90deg the Gaussian is < 2.4e-6 whatever the polynomial gives); sin(2 dTheta)
//     is an odd polynomial (1.6e-8);
//   * S_L = 1 + 0.015 q rsq(20 + q) as two nested FMAs around the rsq;
//   * the three divisions AND the final square root are one rsq: dE = sqrt(N)/den = N rsq(N den^2) over the common
//     denominator den = S_L S_C S_H;
//   * 1/S_H carries the sqrt(2) of dH' and all constant factors are folded into polynomial coefficients;
//   * subtractions are written as FMAs with -1 / negated hoisted scalars so that they pack.
// The first colour (the cell pixel) is scalar and shared by all lanes; the second colour is a lane vector.
// ---------------------------------------------------------------------------------------------------
#define MM_CIEDE_WEIGHT 50.0f
#define MM_CIEDE_AB_SCALE 0.02f  // 1/50 (a, b and C channels)

template <typename V>
MM_HD V mm_ciede2000_stored_v(float L1, float a1, float b1, float C1, V L2, V a2, V b2, V C2)
{
    const V tag = L2;
#define K(c) v_splat((c), tag)
    const float tiny = 1e-30f;

    // ---- G and a' (:50-57): sqrt(t/(t+1)) = t * rsq(t t + t) with t = (Cbar/25)^7 = (C1s + C2s)^7 (25^7 of :46 is the
    //      channel scale); tiny rides in the last FMA of the power so that t > 0 and the rsq stays finite
    const V cbar = v_add(K(C1), C2);
    const V cb2 = v_mul(cbar, cbar);
    const V cb4 = v_mul(cb2, cb2);
    const V cb7 = v_fma(v_mul(cb4, cb2), cbar, K(tiny));
    const V qg = v_mul(cb7, v_rsq(v_fma(cb7, cb7, cb7)));
    const V g1 = v_fma(K(-0.5f), qg, K(1.5f));  // 1 + G
    const V a1p = v_mul(g1, K(a1)), a2p = v_mul(g1, a2);

    // ---- C' (:58-59), deltas (:87-89)
    const V c1p = v_sqrt(v_fma(a1p, a1p, K(b1 * b1)));
    const V c2p = v_sqrt(v_fma(a2p, a2p, v_mul(b2, b2)));
    const V dL = v_add(L2, K(-L1));
    const V dC = v_fma(c1p, K(-1.0f), c2p);
    const V cpbar = v_add(c1p, c2p);  // mean C' / 25

    // ---- dH' (:90-102) from dot / cross; dHq * sqrt2 = (dH'/2)/25
    const V P = v_fma(c1p, c2p, K(tiny));
    const V dot = v_fma(a1p, a2p, v_mul(K(b1), b2));
    // two separately rounded products, never an FMA: identical colours must give cross == 0 and therefore dE == 0
    const V cross = v_sum_of_products_unfused(a1p, b2, a2p, K(-b1));
    // dot > 0: arg = P + dot, half-angle vector (arg, cross); dot <= 0: arg = P - dot, (|cross|, +-arg). Either way
    // arg = P + |dot|: one scalar FADD with an |.| source modifier per lane instead of two packed ops and a select
    const auto pos = v_gt0(dot);
    const V arg = v_add_abs(P, dot);
    const V rs = v_rsq(arg);
    const V qy = v_sel(pos, cross, v_copysign(arg, cross));
    const V dHq = v_mul(rs, qy);

    // ---- mean hue (:104-122): v1 rotated by dh/2, normalised
    // C1'C2' == 0 (:109-110, mean = h1' + h2'): no special case is needed. There dot = cross = 0, so dHq = rsq(tiny) *
    // (+-tiny) ~ 1e-15 and the mean hue only reaches the result through T in z = dHq / S_H and R_T y z, both ~ 0;
    // (mx, my) is then (0, 0) or v1 rotated by 90deg, rn stays finite (tiny under the rsq), so no NaN can form.
    const V qx = v_sel(pos, arg, v_abs(cross));
    const V mx = v_fma(a1p, qx, v_mul(K(-b1), qy));
    const V my = v_fma(K(b1), qx, v_mul(a1p, qy));
    const V rn = v_rsq(v_fma(mx, mx, v_fma(my, my, K(tiny))));
    const V ch = v_mul(mx, rn), sh = v_mul(my, rn);

    // ---- T (:124-127) as quartic + sin * cubic, pre-scaled by kT so that shn = S_H / sqrt2 = 1/sqrt2 + cpbar * Ts
    //      a1 = -0.17 cos30, b1 = -0.17 sin30, a2 = 0.24, a3 = 0.32 cos6, b3 = -0.32 sin6, a4 = -0.2 cos63, b4 = -0.2 sin63
    const float kA1 = -0.17f * 0.86602540378f, kB1 = -0.17f * 0.5f, kA2 = 0.24f, kA3 = 0.32f * 0.99452189536f,
                kB3 = -0.32f * 0.10452846326f, kA4 = -0.20f * 0.45399049974f, kB4 = -0.20f * 0.89100652418f;
    const float kT = 25.0f * 0.015f * 0.70710678118f;
    V tp = K(kT * (8.0f * kA4));
    tp = v_fma(tp, ch, K(kT * (4.0f * kA3)));
    tp = v_fma(tp, ch, K(kT * (2.0f * kA2 - 8.0f * kA4)));
    tp = v_fma(tp, ch, K(kT * (kA1 - 3.0f * kA3)));
    tp = v_fma(tp, ch, K(kT * (1.0f - kA2 + kA4)));
    V tq = K(kT * (8.0f * kB4));
    tq = v_fma(tq, ch, K(kT * (4.0f * kB3)));
    tq = v_fma(tq, ch, K(kT * (-4.0f * kB4)));
    tq = v_fma(tq, ch, K(kT * (kB1 - kB3)));
    const V Ts = v_fma(sh, tq, tp);
    const V shn = v_fma(cpbar, Ts, K(0.70710678118f));

    // ---- dTheta (:129-134): v = 1 - cos(hbar - 275deg); exponent = -(log2 e / (25deg)^2) (hbar - 275deg)^2 = v * poly(v)
    // v in [0, 2]; beyond v = 1 (|hbar - 275deg| > 90deg) the quartic under-estimates the square but stays monotonic, so
    // the Gaussian is < 2^-18.7 = 2.4e-6 there whatever it evaluates to (the true value is smaller still): no clamp
    const V v = v_fma(K(0.99619469809f), sh, v_fma(K(-0.08715574275f), ch, K(1.0f)));
    const float kE = -1.44269504089f / (0.43633231299f * 0.43633231299f);
    // (hbar - 275deg)^2 / v as a cubic in v, fitted with the Gaussian as weight: what matters is the error of R_T, and
    // that is < 1.1e-6 of its range (+-1.73) over the whole circle (the unweighted quartic it replaces gave 1.3e-7)
    V pe = K(kE * 0.04221360751201737f);
    pe = v_fma(pe, v, K(kE * 0.08423404788568702f));
    pe = v_fma(pe, v, K(kE * 0.3338742216190624f));
    pe = v_fma(pe, v, K(kE * 1.9999825857294062f));
    const V gauss = v_ex2(v_mul(pe, v));
    // -2 sin(2 dTheta), 2 dTheta = 60deg * gauss: odd polynomial in gauss (|err| < 1.6e-8); sign and the 2 of R_C folded
    const V g2 = v_mul(gauss, gauss);
    V ps = K(2.0f * 2.647660919774267e-4f);
    ps = v_fma(ps, g2, K(2.0f * -0.01048760534946101f));
    ps = v_fma(ps, g2, K(2.0f * 0.19139485959145958f));
    ps = v_fma(ps, g2, K(2.0f * -1.0471974082480757f));
    const V nsin2dt2 = v_mul(ps, gauss);

    // ---- R_C (:135-136) = 2 sqrt(t/(t+1)), t = (mean C'/25)^7 = cpbar^7; R_T = -sin(2 dTheta) R_C
    const V cp2 = v_mul(cpbar, cpbar);
    const V cp4 = v_mul(cp2, cp2);
    const V cp7 = v_fma(v_mul(cp4, cp2), cpbar, K(tiny));
    const V rt = v_mul(nsin2dt2, v_mul(cp7, v_rsq(v_fma(cp7, cp7, cp7))));

    // ---- S_L, S_C (:138-142); L1 + L2 = Lbar - 50.  S_L = 1 + 0.015 q rsq(20 + q) with q20 = 20 + q is
    //      rsq(q20) (0.015 q20 - 0.3) + 1 (the cancellation in 0.015 q20 - 0.3 costs < 5e-9 absolute on a value >= 1);
    //      sl25 = 25 S_L undoes the 1/25 of dC and dHq
    const V lm = v_add(K(L1), L2);
    const V q20 = v_fma(lm, lm, K(20.0f));
    const V sl25 = v_fma(v_rsq(q20), v_fma(K(25.0f * 0.015f), q20, K(25.0f * -0.3f)), K(25.0f));
    const V sc = v_fma(K(25.0f * 0.045f), cpbar, K(1.0f));

    // ---- dE (:153-157) over the common denominator den = S_L S_C S_H:  sqrt(N) / den = N rsq(N den^2), which replaces
    //      the reciprocal AND the final square root by one rsq. N is a positive-definite form (|R_T| < 1.74 < 2:
    //      N >= 0.13 (Y^2 + Z^2) + X^2), so rounding cannot make it negative; identical colours give N = 0 -> exactly 0.
    //      X, Y, Z carry the true half-scale deltas, den25 = 25 den: the result is (dE/2)/25.
    const V scsh = v_mul(sc, shn);
    const V den25 = v_mul(sl25, scsh);
    const V X = v_mul(dL, scsh);
    const V Y = v_mul(dC, v_mul(sl25, shn));
    const V Z = v_mul(dHq, v_mul(sl25, sc));
    const V N = v_fma(Z, Z, v_fma(Y, v_fma(rt, Z, Y), v_mul(X, X)));
    return v_mul(N, v_rsq(v_fma(N, v_mul(den25, den25), K(tiny))));
#undef K
}

// Full-scale convenience form (tests): (L, a, b, C) per colour, returns dE.
MM_HD float mm_ciede2000(float L1, float a1, float b1, float C1, float L2, float a2, float b2, float C2)
{
    const float s = MM_CIEDE_AB_SCALE;
    return MM_CIEDE_WEIGHT * mm_ciede2000_stored_v<float>(fmaf(0.5f, L1, -25.0f), s * a1, s * b1, s * C1, fmaf(0.5f, L2, -25.0f),
                                                          s * a2, s * b2, s * C2);
}
