// Lean fp64 elementary functions for the evolution kernel.  The CUDA math library's versions carry
// special-case handling (denormals, infinities, huge-argument Payne-Hanek calls) whose branches and
// constant moves cost more than the arithmetic in this latency-bound kernel; the ranges here are known.
// All stay at full double accuracy (errors quoted from a 2e6-point sweep against libm).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD inline
#endif

namespace rc {

// Polynomial coefficients live in constant memory on the device: an FP64 FMA can take a constant-bank
// operand directly, whereas a literal double costs two move instructions per use (ncu: 744 UMOV per
// evaluation before this change).
#if defined(__CUDACC__)
#define RC_COEF static __constant__ double
#else
#define RC_COEF static const double
#endif
RC_COEF RC_SIN_C[8] = {1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0,
                       1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0};
RC_COEF RC_COS_C[9] = {1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0,
                       1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5, 1.0};
RC_COEF RC_LOG_C[9] = {1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01,   // Lg6 Lg4 Lg2
                       1.479819860511658591e-01, 1.818357216161805012e-01, 2.857142874366239149e-01,   // Lg7 Lg5 Lg3
                       6.666666666666735130e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10};  // Lg1 ln2_hi ln2_lo
RC_COEF RC_PIO2_C[4] = {6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050630396597660e-11,
                        2.02226624871116645580e-21};

// 1/sqrt(h), normal positive h: hardware seed (MUFU.RSQ64H, ~2^-22) + one cubic correction
// y1 = y0 (1 + e/2 + 3 e^2/8), e = 1 - h y0^2  (error ~ e^3).
RC_HD double rc_rsqrt(double h) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(h));
    const double t = h * y;
    const double e = fma(-t, y, 1.0);
    const double q = e * fma(0.375, e, 0.5);
    return fma(y, q, y);
#else
    return 1.0 / sqrt(h);
#endif
}

// Same value to within rounding, one level less on the dependent chain: (y e) (1/2 + 3 e / 8) + y.
RC_HD double rc_rsqrt_short(double h) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(h));
    const double e = fma(-(h * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
#else
    return 1.0 / sqrt(h);
#endif
}

// sqrt(x) for x >= 0 (x = 0 gives 0): x * rsqrt(x) plus one Heron correction (<= 1 ulp).
RC_HD double rc_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    const double y = rc_rsqrt(x + 1e-300);
    const double r = x * y;
    return fma(fma(-r, r, x), 0.5 * y, r);
#else
    return sqrt(x);
#endif
}

// 1/t to ~2^-44 (seed + one Newton step): only for the Wilkinson shift, whose accuracy affects the
// convergence rate but never the result.
RC_HD double rc_rcp_approx(double t) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    return fma(y, fma(-t, y, 1.0), y);
#else
    return 1.0 / t;
#endif
}

// log(x) for normal x in (0,1) (the Box-Muller radius): fdlibm's e_log.c algorithm (exponent split,
// s = f/(2+f), degree-7 minimax in s^2; published constants) without its special cases; the
// division is an approximate reciprocal plus one residual correction.  <= 1 ulp on (0,1).
RC_HD double rc_log01(double x) {
#if defined(__CUDA_ARCH__)
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                 // mantissa in [1,2)
    if ((hi & 0x000fffff) >= 0x6a09f) { hi -= 0x00100000; k += 1; }   // > sqrt(2): halve
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double den = 2.0 + f;
    const double rcp = rc_rcp_approx(den);
    double s = f * rcp;
    s = fma(fma(-s, den, f), rcp, s);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, RC_LOG_C[0], RC_LOG_C[1]), RC_LOG_C[2]);
    const double t2 = z * fma(w, fma(w, fma(w, RC_LOG_C[3], RC_LOG_C[4]), RC_LOG_C[5]), RC_LOG_C[6]);
    const double R = t1 + t2, hfsq = 0.5 * f * f, dk = (double)k;
    return fma(dk, RC_LOG_C[7], -((hfsq - fma(s, hfsq + R, dk * RC_LOG_C[8])) - f));
#else
    return log(x);
#endif
}

// sin and cos of r in [-pi/4, pi/4] (Taylor to r^17 / r^16: truncation < 5e-17), rotated by quadrant q.
RC_HD void rc_sincos_quadrant(double r, int q, double* sn, double* cs) {
    const double r2 = r * r;
    double s = RC_SIN_C[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) s = fma(s, r2, RC_SIN_C[k]);
    s = fma(s * r2, r, r);
    double c = RC_COS_C[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) c = fma(c, r2, RC_COS_C[k]);
    const double a = (q & 1) ? c : s;
    const double b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// sincos(x) for |x| < 1e5 (|lambda T| is a few hundred here): three-constant Cody-Waite reduction by
// pi/2 (fdlibm split, exact for |k| < 2^20) + the polynomials above; max error 1.2e-16 on [-1000,1000].
// Larger arguments take the library path.
RC_HD void rc_sincos(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
    if (!(fabs(x) < 1.0e5)) { sincos(x, sn, cs); return; }
    const double k = rint(x * RC_PIO2_C[0]);
    double r = fma(-k, RC_PIO2_C[1], x);
    r = fma(-k, RC_PIO2_C[2], r);
    r = fma(-k, RC_PIO2_C[3], r);
    rc_sincos_quadrant(r, (int)k, sn, cs);
#else
    sincos(x, sn, cs);
#endif
}

// Table-driven sincos for the phase sum (7 per evaluation at N = 7): reduce by pi/32 instead of pi/2 (same
// three-constant Cody-Waite split scaled by 1/16: exact for |k| < 2^20, i.e. |x| < 1e5), look up
// (sin, cos)(j pi/32) — 64 correctly rounded pairs, 1 KB, L1 resident — and finish with degree-7 / degree-8
// polynomials on |r| <= pi/64 and the angle-addition formulas: 17 FP64 instructions and no quadrant selects
// instead of 23 + selects.  Max error 1.6e-16 on [-1e5, 1e5] (tools/sincos_check.cpp).
struct RcSinCos { double s, c; };
#if defined(__CUDACC__)
static __device__ const RcSinCos RC_SC_TAB[64] = {
#include "rc_sincos_table.inc"
};
#endif
static const RcSinCos RC_SC_TAB_HOST[64] = {
#include "rc_sincos_table.inc"
};
RC_COEF RC_PIO32_C[4] = {6.36619772367581382433e-01 * 16.0, 1.57079632673412561417e+00 / 16.0,
                         6.07710050630396597660e-11 / 16.0, 2.02226624871116645580e-21 / 16.0};
RC_COEF RC_SCP_C[7] = {-1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0,               // sin: r + r^3 (c3 + r^2 (c5 + r^2 c7))
                       1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5};       // cos: 1 + r^2 (c2 + r^2 (c4 + r^2 (c6 + r^2 c8)))

RC_HD void rc_sincos_tab_core(double x, const RcSinCos* tab, double* sn, double* cs) {
    const double k = rint(x * RC_PIO32_C[0]);
    double r = fma(-k, RC_PIO32_C[1], x);
    r = fma(-k, RC_PIO32_C[2], r);
    r = fma(-k, RC_PIO32_C[3], r);
    const RcSinCos t = tab[(int)k & 63];
    const double r2 = r * r;
    const double ps = fma(r2, fma(r2, RC_SCP_C[0], RC_SCP_C[1]), RC_SCP_C[2]);
    const double s = fma(r * r2, ps, r);
    const double pc = fma(r2, fma(r2, fma(r2, RC_SCP_C[3], RC_SCP_C[4]), RC_SCP_C[5]), RC_SCP_C[6]);
    const double c = fma(r2, pc, 1.0);
    *sn = fma(t.s, c, t.c * s);
    *cs = fma(t.c, c, -(t.s * s));
}

RC_HD void rc_sincos_tab(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
    if (!(fabs(x) < 1.0e5)) { sincos(x, sn, cs); return; }
    rc_sincos_tab_core(x, RC_SC_TAB, sn, cs);
#else
    sincos(x, sn, cs);
#endif
}

// sincos(2 pi u) for u in [0,1): exact reduction (k = rint(4u), r = (2u - k/2) pi).
RC_HD void rc_sincos_2pi(double u, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
    const double x = u + u;
    const double k = rint(x + x);
    const double r = fma(-0.5, k, x) * 3.14159265358979311600e+00;
    rc_sincos_quadrant(r, (int)k, sn, cs);
#else
    sincos(2.0 * 3.14159265358979323846 * u, sn, cs);
#endif
}

}  // namespace rc
