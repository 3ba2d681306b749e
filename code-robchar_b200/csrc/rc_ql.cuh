// Transfer amplitude <out| exp(-i H T) |in> of a real symmetric tridiagonal H, lane-private.
//
// Replaces the reference's dense complex `scipy.linalg.expm(-1j*T*H)` followed by the
// [out,in] element (noise_model.py:105-109, qnewton.py:397-400).  The reference Hamiltonian is
// complex Hermitian tridiagonal (noise_model.py:135-147); a diagonal phase gauge maps it to the
// real symmetric tridiagonal (d_i, b_i = |1 + nn_i + i nn2_i|) and |U[out,in]|^2 is invariant.
//
// Algorithm: implicit-shift QL with Wilkinson shifts (Givens chase from the bottom of the
// unreduced block), accumulating the plane rotations ONLY into the two rows `in` and `out` of
// the eigenvector matrix.  Then amp = sum_k V[out,k] V[in,k] exp(-i lambda_k T).
// Backward stable => |fid - expm| ~ 1e-14 (tests pin 1e-10).
//
// Register-resident variant (QlReg<N>): all loops over matrix positions are compile-time
// unrolled so d/e/zi/zo live in registers; the eigenvalue index l is a compile-time constant
// (template recursion) and only the per-eigenvalue sweep count is data dependent.  A warp
// evaluates 32 noise draws of the SAME controller, so sweep counts are strongly correlated
// across lanes and divergence stays low.
#pragma once
#ifndef RC_QL_RMIN
#define RC_QL_RMIN 1
#endif
#include <math.h>
#include <float.h>
#include <string.h>
#include "rc_math.cuh"

#if defined(__CUDACC__)
#define RC_D __device__ __forceinline__
#else
#define RC_D inline
#endif

namespace rc {

constexpr int QL_MAX_SWEEPS = 40;  // per eigenvalue (EISPACK uses 30)

#ifdef RC_QL_STATS
struct QlStats { int sweeps_per_l[64]; int total_sweeps; int rotations; int irregular; };
#define RC_STAT(x) x
#else
#define RC_STAT(x)
#endif

// |x| < 2^-20-granular threshold test on the high words: (hi(x) & 0x7fffffff) < hi(tol).  Slightly
// stricter than |x| <= tol (values whose high word equals tol's are kept); zero always passes.
RC_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
RC_HD int threshold_hi(double tol) { int h = hi_word(tol); return h < 1 ? 1 : h; }
RC_HD bool negligible_hi(double x, int tolhi) { return (hi_word(x) & 0x7fffffff) < tolhi; }

RC_HD double wilkinson_g(double dl, double dl1, double el, double dm) {
    const double delta = 0.5 * (dl1 - dl);
    const double e2 = el * el;
    const double q = fma(delta, delta, e2);          // > 0: el is not negligible inside the block
    const double t = delta + copysign(q * rc_rsqrt(q), delta);
    return (dm - dl) + e2 * rc_rcp_approx(t);
}

// Strided-memory solver (defined below); tolhi_given != 0: continue a started solve under its threshold.
RC_HD void amplitude_strided(double* d, double* e, double* zi, double* zo, int ld, int n, double T, int* fail,
                             double& re_out, double& im_out, int tolhi_given = 0);

// One implicit QL sweep on the unreduced block [L, m] (m found by the caller), L compile time.
// e[i] couples sites i and i+1; e[N-1] is a scratch slot.
template <int N, int L>
struct QlSweep {
    static RC_HD void run(double (&d)[N], double (&e)[N], double (&zi)[N], double (&zo)[N], int m, double dm,
                          int mend, double tiny, int& rmin) {
        // Wilkinson shift from the leading 2x2 of the block, single-division form:
        // mu = d[L] - e^2 / (delta + sign(delta) sqrt(delta^2 + e^2)),  g = d[m] - mu
        double g = wilkinson_g(d[L], d[L + 1], e[L], dm);
        double r;
        double s = 1.0, c = 1.0, p = 0.0;
#pragma unroll
        for (int i = N - 2; i >= L; --i) {
            if (i < m) {
                double f = s * e[i];
                double b = c * e[i];
                // h >= tol^2-ish inside an unreduced block; the tiny offset only keeps the (measure-zero)
                // total-cancellation case finite instead of branching on it in the hot loop
                double h = fma(f, f, fma(g, g, tiny));
                double rinv = rc_rsqrt(h);
                r = h * rinv;
                e[i + 1] = r;
                rmin = hi_word(r) < rmin ? hi_word(r) : rmin;   // smallest new coupling of this sweep (r >= 0)
                s = f * rinv;
                c = g * rinv;
                g = d[i + 1] - p;
                r = fma(d[i] - g, s, (2.0 * c) * b);
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                double t = zi[i + 1];
                zi[i + 1] = fma(s, zi[i], c * t);
                zi[i] = fma(c, zi[i], -(s * t));
                t = zo[i + 1];
                zo[i + 1] = fma(s, zo[i], c * t);
                zo[i] = fma(c, zo[i], -(s * t));
            }
        }
        d[L] -= p;
        e[L] = g;
        // the first rotation wrote r into e[m]; that slot is scratch when m is the end of the active
        // block (never read), so the zeroing select chain only runs for a genuine interior split
        if (m != mend) {
#pragma unroll
            for (int i = L + 1; i < N; ++i)
                if (i == m) e[i] = 0.0;
        }
    }
};

template <int N, int L>
struct QlLevel {
    static RC_HD int run(double (&d)[N], double (&e)[N], double (&zi)[N], double (&zo)[N], double tol
#ifdef RC_QL_STATS
                         , QlStats* st
#endif
    ) {
        int fail = 0;
        if constexpr (L < N - 1) {
            int it = 0;
            while (true) {
                // first negligible off-diagonal at or after L (descending scan: smallest index wins)
                int m = N - 1;
                double dm = d[N - 1];
#pragma unroll
                for (int i = N - 2; i >= L; --i) {
                    if (fabs(e[i]) <= tol) { m = i; dm = d[i]; }
                }
                if (m == L) break;
                if (++it > QL_MAX_SWEEPS) { fail = 1; break; }
                int rmin_unused = 0x7fffffff;
                QlSweep<N, L>::run(d, e, zi, zo, m, dm, -1, 1e-280, rmin_unused);
                RC_STAT(st->sweeps_per_l[L]++; st->total_sweeps++; st->rotations += m - L;)
            }
            fail |= QlLevel<N, L + 1>::run(d, e, zi, zo, tol
#ifdef RC_QL_STATS
                                           , st
#endif
            );
        }
        return fail;
    }
};

// amp = sum_k zo[k] zi[k] exp(-i d[k] T)
template <int N>
RC_HD void phase_sum(const double (&d)[N], const double (&zi)[N], const double (&zo)[N], double T, double& re,
                     double& im) {
    re = 0.0; im = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double sn, cs;
        sincos(d[k] * T, &sn, &cs);
        double w = zo[k] * zi[k];
        re = fma(w, cs, re);
        im = fma(-w, sn, im);
    }
}

// Full register-resident evaluation.  d[0..N-1] diagonal, e[0..N-2] off-diagonal (e[N-1] ignored).
// Returns fidelity; *fail set to 1 when QL did not converge (result NaN).
template <int N>
RC_HD double fidelity_reg(double (&d)[N], double (&e)[N], int in, int out, double T, int* fail
#ifdef RC_QL_STATS
                          , QlStats* st
#endif
) {
    double zi[N], zo[N];
    double anorm = 0.0, chk = 0.0;
    e[N - 1] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        zi[k] = (k == in) ? 1.0 : 0.0;
        zo[k] = (k == out) ? 1.0 : 0.0;
        anorm = fmax(anorm, fabs(d[k]) + fabs(e[k]));
        chk += d[k] + e[k];
    }
    chk += T;
    if (!(fabs(chk) <= DBL_MAX)) {  // NaN / Inf controller or draw -> NaN fidelity (mcsim.py:369-374)
        *fail = 0;
        return NAN;
    }
    double tol = DBL_EPSILON * anorm;
    int f = QlLevel<N, 0>::run(d, e, zi, zo, tol
#ifdef RC_QL_STATS
                               , st
#endif
    );
    *fail = f;
    if (f) return NAN;
    double re, im;
    phase_sum<N>(d, zi, zo, T, re, im);
    return fma(re, re, im * im);
}

// ---------------------------------------------------------------------------------------------
// Compact register-resident variant (the one the kernels use).
//
// The fully recursive version above instantiates one sweep body per eigenvalue index (N(N-1)/2
// rotation bodies: 90 KB of SASS at N=7, which thrashes the instruction cache — ncu showed
// `no_instruction` as the top stall).  Here the active block always starts at position 0: when
// e[0] becomes negligible, eigenvalue d[0] and its weight zi[0]*zo[0] are written to a per-lane
// scratch row and all four arrays are shifted down by one (register moves).  There is a single
// sweep body (N-1 predicated rotation slots, compile-time indices), every lane runs its own
// sequence of sweeps, and a warp iterates max-over-lanes(total sweeps) times.
// scratch: 2N doubles per lane, element k at scratch[k * sstride].
// ---------------------------------------------------------------------------------------------
#ifndef RC_QL_PINNED_END
#define RC_QL_PINNED_END 1
#endif
#ifndef RC_QL_SHIFT_AT_TOP
#define RC_QL_SHIFT_AT_TOP 1
#endif
#if RC_QL_PINNED_END
// ---------------------------------------------------------------------------------------------
// Pinned-end form (round 2, what the kernels use).  The block-at-index-0 form below pays for its
// static shift end with data movement: every deflation shifts the four register arrays down by one
// (48 moves at N=7), the chase starts at a dynamic position (a select chain for d[m], a predicate per
// rotation slot that skips the slots above m) and finished eigenpairs are stored through a dynamic
// index.  Here nothing moves: the active block is [L, N-1] with the END pinned at N-1, so the chase
// always starts with the static slot N-2 (s = c = 1, p = 0 folded in: two multiplications and the
// dead store of the first r less per sweep) and LEAVES after the slot that rotates (L, L+1).  The chase
// value g travels in e[i] and the pending "- p" in d[i] (each slot stores d[i] - p, which is what the next
// slot would compute first and is the final diagonal element if there is no next slot), so a slot has
// no epilogue that depends on being the last one.  The slot that was the last one (static index again)
// hands the three numbers of the next Wilkinson shift — of block L, or of block L+1 when e[L] has just
// become negligible — to the common tail; deflated eigenpairs simply stay where they are.
// Every arithmetic operation and its order equal the form below: results are bit-identical.
//
// The price is generality: a sweep cannot stop short of N-1.  A negligible INTERIOR coupling (seen as
// in the other form by `rmin`, the smallest coupling the sweep wrote) would need exactly that, so such
// evaluations (measured: a few in 1e5 at the paper's noise levels, plus exactly decoupled chains) leave
// for ql_irregular: the state goes to a lane-private local-memory array and the strided solver finishes it
// from where the sweeps stand, out of line.
// ---------------------------------------------------------------------------------------------
struct QlTail { double dl, dl1, el; bool defl; };

template <int N, int I>
RC_HD void ql_rows(double (&zi)[N], double (&zo)[N], double s, double c) {
    double t = zi[I + 1];
    zi[I + 1] = fma(s, zi[I], c * t);
    zi[I] = fma(c, zi[I], -(s * t));
    t = zo[I + 1];
    zo[I + 1] = fma(s, zo[I], c * t);
    zo[I] = fma(c, zo[I], -(s * t));
}

// operands of the next shift after slot I was the last of the sweep
template <int N, int I>
RC_HD void ql_tail(const double (&d)[N], const double (&e)[N], int tolhi, QlTail& t) {
    const bool ng = negligible_hi(e[I], tolhi);
    t.defl = ng;
    if constexpr (I + 2 <= N - 1) {
        t.dl = ng ? d[I + 1] : d[I];
        t.dl1 = ng ? d[I + 2] : d[I + 1];
        t.el = ng ? e[I + 1] : e[I];
    } else {   // block (N-2, N-1): deflation ends the solve, the shift is not used
        t.dl = d[I]; t.dl1 = d[I + 1]; t.el = e[I];
    }
}

template <int N, int I>
struct QlChain {
    static RC_HD void run(double (&d)[N], double (&e)[N], double (&zi)[N], double (&zo)[N], double s, double c,
                          int L, double tiny, int tolhi, int& rmin, QlTail& t) {
        const double g = e[I + 1];
        const double f = s * e[I];
        const double b = c * e[I];
        const double h = fma(f, f, fma(g, g, tiny));
        const double rinv = rc_rsqrt(h);
        const double r = h * rinv;
        e[I + 1] = r;
        rmin = hi_word(r) < rmin ? hi_word(r) : rmin;   // smallest interior coupling of this sweep (r >= 0)
        s = f * rinv;
        c = g * rinv;
        const double gg = d[I + 1];                      // already d - p of the slot before
        const double r2 = fma(d[I] - gg, s, (2.0 * c) * b);
        const double p = s * r2;
        d[I + 1] = gg + p;
        e[I] = c * r2 - b;
        d[I] = d[I] - p;
        ql_rows<N, I>(zi, zo, s, c);
        if constexpr (I == 0) {
            ql_tail<N, 0>(d, e, tolhi, t);
        } else {
            if (L == I) ql_tail<N, I>(d, e, tolhi, t);
            else QlChain<N, I - 1>::run(d, e, zi, zo, s, c, L, tiny, tolhi, rmin, t);
        }
    }
};

// Out-of-line continuation of an evaluation the pinned-end sweeps cannot finish (negligible interior
// coupling).  st = d[N], e[N], zi[N], zo[N] of the lane; the strided solver skips the deflated head
// (its couplings are negligible under the same threshold) and goes on exactly as the block-at-0 form would.
template <int N>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
void ql_irregular(double* st, double T, int tolhi, int* fail, double* reim) {
    double re, im;
    amplitude_strided(st, st + N, st + 2 * N, st + 3 * N, 1, N, T, fail, re, im, tolhi);
    reim[0] = re; reim[1] = im;
}

template <int N, bool AMP = false>
RC_HD double fidelity_reg_compact(double (&d)[N], double (&e)[N], int in, int out, double T, double* scratch,
                                  int sstride, int* fail
#ifdef RC_QL_STATS
                                  , QlStats* st
#endif
                                  , double* amp = nullptr
) {
    double zi[N], zo[N];
    double anorm = 0.0, chk = T;
    if (AMP) { amp[0] = NAN; amp[1] = NAN; }
    e[N - 1] = 0.0;
    int emin = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        zi[k] = (k == in) ? 1.0 : 0.0;
        zo[k] = (k == out) ? 1.0 : 0.0;
        anorm = fmax(anorm, fabs(d[k]) + fabs(e[k]));
        chk += d[k] + e[k];
        if (k < N - 1) { const int h = hi_word(e[k]) & 0x7fffffff; emin = h < emin ? h : emin; }
    }
    *fail = 0;
    if (!(fabs(chk) <= DBL_MAX)) return NAN;  // NaN / Inf controller or draw (mcsim.py:369-374)
    const double tol = DBL_EPSILON * anorm;
    const int tolhi = threshold_hi(tol);
    const double tiny = fmin(tol, 1e-280);  // == 1e-280, kept in a register instead of re-materialised per rotation
    double re = 0.0, im = 0.0;
    // L = first site of the active block; N = "left for ql_irregular".  Non-convergence (more than
    // QL_MAX_SWEEPS sweeps on one eigenvalue) leaves the loop through its condition with L < N - 1.
    int L = emin < tolhi ? N : 0, it = 0;
    if constexpr (N >= 2) {
#if RC_QL_SHIFT_AT_TOP
        QlTail t;
        t.dl = d[0]; t.dl1 = d[1]; t.el = e[0]; t.defl = false;
        while (L < N - 1 && it <= QL_MAX_SWEEPS) {
            // the shift of this sweep from the three numbers the previous sweep's last slot handed over (computed here
            // rather than at the end of the trip: the trip that deflates the last pair does not pay for a shift nobody uses)
            const double g = wilkinson_g(t.dl, t.dl1, t.el, d[N - 1]);
#else
        double g = wilkinson_g(d[0], d[1], e[0], d[N - 1]);
        while (L < N - 1 && it <= QL_MAX_SWEEPS) {
            QlTail t;
#endif
            ++it;
            int rmin = 0x7fffffff;
            {   // slot N-2: first rotation of every sweep, s = c = 1 and p = 0 folded in, its r is not stored
                constexpr int I = N - 2;
                const double b = e[I];
                const double h = fma(b, b, fma(g, g, tiny));
                const double rinv = rc_rsqrt(h);
                const double s = b * rinv;
                const double c = g * rinv;
                const double gg = d[I + 1];
                const double r2 = fma(d[I] - gg, s, (2.0 * c) * b);
                const double p = s * r2;
                d[I + 1] = gg + p;
                e[I] = c * r2 - b;
                d[I] = d[I] - p;
                ql_rows<N, I>(zi, zo, s, c);
                if constexpr (I == 0) {
                    ql_tail<N, 0>(d, e, tolhi, t);
                } else {
                    if (L == I) ql_tail<N, I>(d, e, tolhi, t);
                    else QlChain<N, I - 1>::run(d, e, zi, zo, s, c, L, tiny, tolhi, rmin, t);
                }
            }
            RC_STAT(st->total_sweeps++; st->rotations += N - 1 - L;)
            if (t.defl) { ++L; it = 0; }
            if (rmin < tolhi) L = N;     // negligible interior coupling: ql_irregular takes over
#if !RC_QL_SHIFT_AT_TOP
            g = wilkinson_g(t.dl, t.dl1, t.el, d[N - 1]);
#endif
        }
    }
    if (L == N) {
        RC_STAT(st->irregular++;)
        double buf[4 * N], reim[2];
#pragma unroll
        for (int k = 0; k < N; ++k) { buf[k] = d[k]; buf[N + k] = e[k]; buf[2 * N + k] = zi[k]; buf[3 * N + k] = zo[k]; }
        ql_irregular<N>(buf, T, tolhi, fail, reim);
        re = reim[0]; im = reim[1];
        if (*fail) return NAN;
    } else {
        if (L < N - 1) { *fail = 1; return NAN; }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            scratch[(size_t)k * sstride] = d[k];
            scratch[(size_t)(N + k) * sstride] = zi[k] * zo[k];
        }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int k = 0; k < N; ++k) {
            double sn, cs;
            rc_sincos_tab(scratch[(size_t)k * sstride] * T, &sn, &cs);
            const double w = scratch[(size_t)(N + k) * sstride];
            re = fma(w, cs, re);
            im = fma(-w, sn, im);
        }
    }
    if (AMP) { amp[0] = re; amp[1] = im; }
    return fma(re, re, im * im);
}
#else
// AMP: also stores the transfer amplitude <out| exp(-iHT) |in> as amp[0] + i amp[1] (NaN where the fidelity is NaN).
template <int N, bool AMP = false>
RC_HD double fidelity_reg_compact(double (&d)[N], double (&e)[N], int in, int out, double T, double* scratch,
                                  int sstride, int* fail
#ifdef RC_QL_STATS
                                  , QlStats* st
#endif
                                  , double* amp = nullptr
) {
    double zi[N], zo[N];
    double anorm = 0.0, chk = T;
    if (AMP) { amp[0] = NAN; amp[1] = NAN; }
    e[N - 1] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        zi[k] = (k == in) ? 1.0 : 0.0;
        zo[k] = (k == out) ? 1.0 : 0.0;
        anorm = fmax(anorm, fabs(d[k]) + fabs(e[k]));
        chk += d[k] + e[k];
    }
    *fail = 0;
    if (!(fabs(chk) <= DBL_MAX)) return NAN;  // NaN / Inf controller or draw (mcsim.py:369-374)
    const double tol = DBL_EPSILON * anorm;
    const int tolhi = threshold_hi(tol);
    const double tiny = fmin(tol, 1e-280);  // == 1e-280, kept in a register instead of re-materialised per rotation
#if RC_QL_RMIN
    // rmin = smallest high word among the couplings the last sweep produced (a full-block sweep rewrites
    // every interior coupling of the active block), so "no interior split" is ONE integer compare per
    // trip; the descending scan for the split position only runs when that compare fails (first trip,
    // and in the rare genuine split).  The block handed to the sweep must be unreduced: a zero interior
    // coupling would drive g to exactly 0 and break the next rotation.
    // Non-convergence (more than QL_MAX_SWEEPS sweeps on one eigenvalue) leaves the loop through its
    // condition with nact > 1 — no `break`, no flag to keep live across the loop body.
    int nact = N, ndone = 0, it = 0, rmin = 0;
    while (nact > 1 && it <= QL_MAX_SWEEPS) {
        if (negligible_hi(e[0], tolhi)) {
            // deflate: eigenvalue d[0] with weight V[in,k] V[out,k]
            scratch[(size_t)ndone * sstride] = d[0];
            scratch[(size_t)(N + ndone) * sstride] = zi[0] * zo[0];
            ++ndone; --nact; it = 0;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) { d[i] = d[i + 1]; e[i] = e[i + 1]; zi[i] = zi[i + 1]; zo[i] = zo[i + 1]; }
        }
        // sweep in the same trip (lanes that deflated stay converged with the lanes that did not)
        // unless the new leading off-diagonal is negligible as well
        if (nact > 1 && !negligible_hi(e[0], tolhi)) {
            int m = nact - 1;
            bool split = false;
            if (rmin < tolhi) {
                // first negligible off-diagonal inside the active block (descending scan: smallest wins);
                // compares the high words as integers (ALU pipe) instead of fp64 compares
#pragma unroll
                for (int i = N - 2; i >= 1; --i)
                    if (i < nact - 1 && negligible_hi(e[i], tolhi)) m = i;
                split = m != nact - 1;
            }
            double dm = d[N - 1];
#pragma unroll
            for (int i = N - 2; i >= 1; --i)
                if (i == m) dm = d[i];
            ++it;
            rmin = 0x7fffffff;
            QlSweep<N, 0>::run(d, e, zi, zo, m, dm, nact - 1, tiny, rmin);
            if (split) rmin = 0;   // e[m] = 0 stays inside the block until [0, m] is deflated: keep scanning
            RC_STAT(st->total_sweeps++; st->rotations += m;)
        }
    }
    const int bad = nact > 1;
#else
    int nact = N, ndone = 0, it = 0;
    while (nact > 1 && it <= QL_MAX_SWEEPS) {
        if (negligible_hi(e[0], tolhi)) {
            // deflate: eigenvalue d[0] with weight V[in,k] V[out,k]
            scratch[(size_t)ndone * sstride] = d[0];
            scratch[(size_t)(N + ndone) * sstride] = zi[0] * zo[0];
            ++ndone; --nact; it = 0;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) { d[i] = d[i + 1]; e[i] = e[i + 1]; zi[i] = zi[i + 1]; zo[i] = zo[i + 1]; }
        }
        // sweep in the same trip unless the new leading off-diagonal is negligible as well
        if (nact > 1 && !negligible_hi(e[0], tolhi)) {
            // first negligible off-diagonal inside the active block (descending scan: smallest wins).
            // The block handed to the sweep must be unreduced: a zero interior coupling would drive
            // g to exactly 0 and break the next rotation, so this scan runs every trip; it compares
            // the high words as integers (ALU pipe) instead of fp64 compares.
            int m = nact - 1;
#pragma unroll
            for (int i = N - 2; i >= 1; --i)
                if (i < nact - 1 && negligible_hi(e[i], tolhi)) m = i;
            double dm = d[N - 1];
#pragma unroll
            for (int i = N - 2; i >= 1; --i)
                if (i == m) dm = d[i];
            ++it;
            int rmin_unused = 0x7fffffff;
            QlSweep<N, 0>::run(d, e, zi, zo, m, dm, nact - 1, tiny, rmin_unused);
            RC_STAT(st->total_sweeps++; st->rotations += m;)
        }
    }
    const int bad = nact > 1;
#endif
    if (bad) { *fail = 1; return NAN; }
    scratch[(size_t)ndone * sstride] = d[0];
    scratch[(size_t)(N + ndone) * sstride] = zi[0] * zo[0];
    double re = 0.0, im = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < N; ++k) {
        double sn, cs;
        rc_sincos_tab(scratch[(size_t)k * sstride] * T, &sn, &cs);
        const double w = scratch[(size_t)(N + k) * sstride];
        re = fma(w, cs, re);
        im = fma(-w, sn, im);
    }
    if (AMP) { amp[0] = re; amp[1] = im; }
    return fma(re, re, im * im);
}

#endif  // RC_QL_PINNED_END

// ---------------------------------------------------------------------------------------------
// Strided-memory variant for large N: arrays live in shared (or any) memory with element stride
// `ld` between consecutive matrix positions (column = this lane), dynamic loop bounds.
// Holds the "i+1" elements in registers while chasing upwards to halve the memory traffic.
// ---------------------------------------------------------------------------------------------
RC_HD void amplitude_strided(double* d, double* e, double* zi, double* zo, int ld, int n, double T, int* fail,
                             double& re_out, double& im_out, int tolhi_given) {
#define AT(a, i) a[(size_t)(i) * ld]
    double anorm = 0.0, chk = T;
    AT(e, n - 1) = 0.0;
    for (int k = 0; k < n; ++k) {
        anorm = fmax(anorm, fabs(AT(d, k)) + fabs(AT(e, k)));
        chk += AT(d, k) + AT(e, k);
    }
    re_out = NAN; im_out = NAN;
    if (!(fabs(chk) <= DBL_MAX)) { *fail = 0; return; }
    const double tol = DBL_EPSILON * anorm;
    const int tolhi = tolhi_given ? tolhi_given : threshold_hi(tol);   // given: continuation of a started solve (ql_irregular)
    const double tiny = fmin(tol, 1e-280);
    int bad = 0;
    // rmin: smallest high word among the couplings the last sweep wrote (see fidelity_reg_compact) — the
    // scan for an interior split (a dependent chain of shared-memory loads) only runs when it says so.
    int rmin = 0;
    // One flat loop: every lane walks its own sequence (deflate while e[l] is negligible, else sweep), so a
    // warp runs max-over-lanes(total sweeps) trips instead of the sum over l of max-over-lanes(sweeps at l).
    // No `break` / `continue` in the body: structured exits keep the warp reconverging every trip (the
    // register kernel gained 18 % from removing its non-convergence `break`).  Non-convergence leaves through
    // the loop condition with l < n - 1.
    int l = 0, it = 0;
    while (l < n - 1 && it <= QL_MAX_SWEEPS) {
        if (negligible_hi(AT(e, l), tolhi)) { ++l; it = 0; }
        if (l < n - 1 && !negligible_hi(AT(e, l), tolhi)) {
        int m = n - 1;
        if (rmin < tolhi) {
            m = l + 1;
            while (m < n - 1 && !negligible_hi(AT(e, m), tolhi)) ++m;
        }
        const bool split = m != n - 1;
        ++it;
        double g = wilkinson_g(AT(d, l), AT(d, l + 1), AT(e, l), AT(d, m));
        double r;
        double s = 1.0, c = 1.0, p = 0.0;
        double d_up = AT(d, m), zi_up = AT(zi, m), zo_up = AT(zo, m);  // values at i+1
        // running pointers to position i of each array (one subtraction per array per rotation
        // instead of an index multiply per access); the operands of rotation i-1 are loaded while
        // rotation i computes (they do not depend on the chase), so the shared-memory latency stays
        // off the dependent chain
        double* pe = e + (size_t)(m - 1) * ld;
        double* pd = d + (size_t)(m - 1) * ld;
        double* pzi = zi + (size_t)(m - 1) * ld;
        double* pzo = zo + (size_t)(m - 1) * ld;
        double ei = *pe, di = *pd, zii = *pzi, zoi = *pzo;
        rmin = 0x7fffffff;
        for (int i = m - 1; i >= l; --i) {
            const int back = i > l ? ld : 0;   // clamp: the last prefetch re-reads position l
            const double ein = *(pe - back), din = *(pd - back), ziin = *(pzi - back), zoin = *(pzo - back);
            double f = s * ei, b = c * ei;
            double h = fma(f, f, fma(g, g, tiny));
            double rinv = rc_rsqrt(h);
            r = h * rinv;
            pe[ld] = r;
            rmin = hi_word(r) < rmin ? hi_word(r) : rmin;
            s = f * rinv;
            c = g * rinv;
            g = d_up - p;
            r = fma(di - g, s, (2.0 * c) * b);
            p = s * r;
            pd[ld] = g + p;
            g = c * r - b;
            pzi[ld] = fma(s, zii, c * zi_up);
            zi_up = fma(c, zii, -(s * zi_up));
            pzo[ld] = fma(s, zoi, c * zo_up);
            zo_up = fma(c, zoi, -(s * zo_up));
            d_up = di;
            ei = ein; di = din; zii = ziin; zoi = zoin;
            pe -= ld; pd -= ld; pzi -= ld; pzo -= ld;
        }
        AT(zi, l) = zi_up;
        AT(zo, l) = zo_up;
        AT(d, l) = d_up - p;
        AT(e, l) = g;
        AT(e, m) = 0.0;
        if (split) rmin = 0;   // the zero stays inside [l, n-1] until [l, m] is deflated: keep scanning
        }
    }
    bad = l < n - 1;
    *fail = bad;
    if (bad) return;
    double re = 0.0, im = 0.0;
    for (int k = 0; k < n; ++k) {
        double sn, cs;
        rc_sincos_tab(AT(d, k) * T, &sn, &cs);
        double w = AT(zo, k) * AT(zi, k);
        re = fma(w, cs, re);
        im = fma(-w, sn, im);
    }
#undef AT
    re_out = re; im_out = im;
}

RC_HD double fidelity_strided(double* d, double* e, double* zi, double* zo, int ld, int n, double T, int* fail) {
    double re, im;
    amplitude_strided(d, e, zi, zo, ld, n, T, fail, re, im);
    return fma(re, re, im * im);
}

}  // namespace rc
