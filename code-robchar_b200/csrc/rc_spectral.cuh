// Transfer amplitude from eigenvalues alone ("spectral weights"), lane-private, strided storage.
//
// Same quantity as rc_ql.cuh — <out| exp(-i H T) |in> of the real symmetric tridiagonal H that replaces the
// reference's dense `scipy.linalg.expm(-1j*T*H)[out, in]` (noise_model.py:105-109, qnewton.py:397-400) — but
// without accumulating eigenvector rows.  For an unreduced symmetric tridiagonal matrix with diagonal d and
// couplings b the residue of the resolvent gives, for a <= b,
//
//     V[a,k] V[b,k] = chi_{0..a-1}(lambda_k) * (b_a ... b_{b-1}) * chi_{b+1..N-1}(lambda_k) / chi'(lambda_k),
//     chi'(lambda_k) = prod_{j != k} (lambda_k - lambda_j),
//
// chi_{p..q} = characteristic polynomial of the principal block on sites p..q (three-term recurrence, empty
// block = 1).  So the evaluation is: implicit-shift QL for the EIGENVALUES only (21 instead of 29 FP64
// instructions per rotation, and two [N] arrays per lane instead of four: twice the resident lanes per SM in
// the shared-memory kernels), N(N-1) differences and products, and the phase sum.  For the end-to-end transfer
// (in = 0, out = N-1: BASELINE configs 3-5) there are no minors at all.
//
// Accuracy.  The weights only see DIFFERENCES of computed eigenvalues; a pair at distance g carries a relative
// error ~ eps*|H|/g, identical for both members of the pair, which multiplies that pair's contribution to the
// amplitude.  An a-posteriori estimate (per eigenvalue: |w_k| (N-1) 4 eps |H| / min_j |lambda_k - lambda_j|,
// plus 8 eps times the uncancelled magnitude of the minors) is accumulated next to the amplitude; when it
// exceeds SPEC_EST_THR (or is not finite: exactly coincident computed eigenvalues), the caller recomputes that
// evaluation with the eigenvector-accumulating QL of rc_ql.cuh.  Measured against that QL on 5e6 evaluations of
// the reference's controllers (N = 4..7, sigma <= 0.1): max |delta fid| 2.5e-13, no fallback taken; mirror-
// symmetric / double-well chains up to N = 32 (near-degenerate pairs with O(1) weights): <= 3e-14 where the
// estimate accepts (tools/spectral_sim.cpp).
#pragma once
#include "rc_ql.cuh"

namespace rc {

constexpr double SPEC_EST_THR = 1e-11;   // accepted a-posteriori error estimate of the amplitude

#define RC_AT(a, i) a[(size_t)(i) * ld]

// Eigenvalues of the symmetric tridiagonal (d, e), in place in d (unordered); e is destroyed.  Same flat
// deflate / sweep loop and the same `rmin` split test as amplitude_strided, minus the eigenvector rows.
// Returns 1 when an eigenvalue needed more than QL_MAX_SWEEPS sweeps.
// LD > 0: the column stride is the compile-time constant LD (it becomes an immediate offset of the shared-memory
// loads / stores of the chase: no address arithmetic per rotation); LD = 0: run-time stride ld_rt.
// The operands of rotation i-1 are fetched unconditionally while rotation i computes: for i = l = 0 that reads one
// row BEFORE the arrays, so the caller keeps one pad row in front of d (e follows d, so its row -1 is d's last row).
template <int LD = 0>
RC_HD int ql_eigenvalues_strided(double* d, double* e, int ld_rt, int n, int tolhi, double tiny) {
    const int ld = LD ? LD : ld_rt;
    int rmin = 0, l = 0, it = 0;
    while (l < n - 1 && it <= QL_MAX_SWEEPS) {
        if (negligible_hi(RC_AT(e, l), tolhi)) { ++l; it = 0; }
        if (l < n - 1 && !negligible_hi(RC_AT(e, l), tolhi)) {
            int m = n - 1;
            if (rmin < tolhi) {
                m = l + 1;
                while (m < n - 1 && !negligible_hi(RC_AT(e, m), tolhi)) ++m;
            }
            const bool split = m != n - 1;
            ++it;
            double g = wilkinson_g(RC_AT(d, l), RC_AT(d, l + 1), RC_AT(e, l), RC_AT(d, m));
            double r, s = 1.0, c = 1.0, p = 0.0;
            double d_up = RC_AT(d, m);
            double* pe = e + (size_t)(m - 1) * ld;
            double* pd = d + (size_t)(m - 1) * ld;
            double ei = *pe, di = *pd;
            rmin = 0x7fffffff;
            for (int i = m - 1; i >= l; --i) {
                const double ein = *(pe - ld), din = *(pd - ld);   // operands of rotation i-1 (pad row when i = 0)
                const double f = s * ei, b = c * ei;
                const double h = fma(f, f, fma(g, g, tiny));
                // Two levels off the dependent chain at the same operation count: 2b is ready before c, and the
                // cubic correction of 1/sqrt(h) multiplies y e and (1/2 + 3e/8) side by side (rc_rsqrt_short).
                // Measured on B200 (profiles/README.md, r02x): +0.4 % at N = 16 (six warps per scheduler hide the
                // chain), +2.4 % at N = 28 / 32 (three).  Taking (di - g') f + 2 g b out of the chain as well costs
                // one more multiplication: +2.9 % at N = 32 but -0.7 % at N = 16 — not taken.
                const double b2 = b + b;
                const double gp = d_up - p, dg = di - gp;
                const double rinv = rc_rsqrt_short(h);
                r = h * rinv;
                pe[ld] = r;
                rmin = hi_word(r) < rmin ? hi_word(r) : rmin;
                s = f * rinv;
                c = g * rinv;
                r = fma(dg, s, c * b2);
                p = s * r;
                pd[ld] = gp + p;
                g = fma(c, r, -b);
                d_up = di;
                ei = ein; di = din;
                pe -= ld; pd -= ld;
            }
            RC_AT(d, l) = d_up - p;
            RC_AT(e, l) = g;
            RC_AT(e, m) = 0.0;
            if (split) rmin = 0;
        }
    }
    return l < n - 1;
}

// 1/x to full precision for finite non-zero x (seed 2^-22 + two Newton steps); x = 0 gives inf.
RC_HD double rc_rcp_full(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    return fma(y, fma(-x, y, 1.0), y);
#else
    return 1.0 / x;
#endif
}
// 1/x to ~20 bits (error-estimate arithmetic only)
RC_HD double rc_rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return 1.0 / x;
#endif
}
RC_HD double hi_as_double(int hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, 0);
#else
    long long b = (long long)(unsigned)hi << 32;
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

// Original entries of the principal blocks before `a = min(in,out)` and after `b = max(in,out)`, kept aside
// before the QL destroys (d, e):  xd[0..na)   = d[0..a),      xe[j] couples xd[j], xd[j+1]   (j < na-1)
//                                 xd[na..na+nb) = d[b+1..n),  xe[na+j] couples xd[na+j], xd[na+j+1]
struct SpecBlocks { const double* xd; const double* xe; int na, nb; };

// amp = sum_k w_k exp(-i lambda_k T) from the eigenvalues in d[0..n); pb = product of the couplings between
// the two sites; anorm = max_i(|d_i| + |e_i|) of the original matrix.  *est receives the error estimate.
template <int LD = 0>
RC_HD void spectral_phase_sum(const double* d, int ld_rt, int n, double T, double pb, double anorm, const SpecBlocks& xb,
                              double& re_out, double& im_out, double* est_out) {
    const int ld = LD ? LD : ld_rt;
    double re = 0.0, im = 0.0, est = 0.0;
    const double cgap = (double)(n - 1) * 4.0 * DBL_EPSILON * anorm;
    const bool minors = (xb.na + xb.nb) > 0;
    // four eigenvalues per pass over the spectrum: four independent product chains, one load per four pairs.
    // Code size matters (the kernel shares the instruction cache between warps that are in different phases):
    // the pair loops and the per-eigenvalue epilogue are deliberately NOT unrolled.
    for (int k0 = 0; k0 < n; k0 += 4) {
        double lam[4], P[4];
        int mh[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + q < n ? k0 + q : n - 1;
            lam[q] = RC_AT(d, k);
            P[q] = 1.0;
            mh[q] = 0x7fffffff;
        }
        const int kend = k0 + 4 < n ? k0 + 4 : n;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < n; ++j) {
            const double lj = RC_AT(d, j);
            const bool inblock = j >= k0 && j < kend;      // warp uniform
            if (!inblock) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double df = lam[q] - lj;
                    P[q] *= df;
                    const int h = hi_word(df) & 0x7fffffff;
                    mh[q] = h < mh[q] ? h : mh[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {               // the block itself: skip j == k (and the clamped duplicates)
                    const double df = lam[q] - lj;
                    const bool self = (k0 + q == j) || (k0 + q >= n);
                    P[q] *= self ? 1.0 : df;
                    const int h = self ? 0x7fffffff : (hi_word(df) & 0x7fffffff);
                    mh[q] = h < mh[q] ? h : mh[q];
                }
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int q = 0; q < kend - k0; ++q) {
            // q is a run-time index: select from the named registers instead of indexing the arrays
            double l0 = lam[0], Pq = P[0];
            int mq = mh[0];
            if (q == 1) { l0 = lam[1]; Pq = P[1]; mq = mh[1]; }
            if (q == 2) { l0 = lam[2]; Pq = P[2]; mq = mh[2]; }
            if (q == 3) { l0 = lam[3]; Pq = P[3]; mq = mh[3]; }
            double num = pb, mag = fabs(pb);
            if (minors) {
                if (xb.na > 0) {
                    double p0 = 1.0, p1 = l0 - RC_AT(xb.xd, 0), a0 = 1.0, a1 = fabs(p1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                    for (int j = 1; j < xb.na; ++j) {
                        const double x = l0 - RC_AT(xb.xd, j), b2 = RC_AT(xb.xe, j - 1) * RC_AT(xb.xe, j - 1);
                        const double t = fma(x, p1, -(b2 * p0)), at = fma(fabs(x), a1, b2 * a0);
                        p0 = p1; p1 = t; a0 = a1; a1 = at;
                    }
                    num *= p1; mag *= a1;
                }
                if (xb.nb > 0) {
                    const int top = xb.na + xb.nb - 1;
                    double p0 = 1.0, p1 = l0 - RC_AT(xb.xd, top), a0 = 1.0, a1 = fabs(p1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                    for (int j = top - 1; j >= xb.na; --j) {
                        const double x = l0 - RC_AT(xb.xd, j), b2 = RC_AT(xb.xe, j) * RC_AT(xb.xe, j);
                        const double t = fma(x, p1, -(b2 * p0)), at = fma(fabs(x), a1, b2 * a0);
                        p0 = p1; p1 = t; a0 = a1; a1 = at;
                    }
                    num *= p1; mag *= a1;
                }
            }
            const double y = rc_rcp_full(Pq);
            double w = num * y;
            w = fma(fma(-w, Pq, num), y, w);
            est = fma(fabs(w) * cgap, rc_rcp_seed(hi_as_double(mq)), est);
            if (minors) est = fma(8.0 * DBL_EPSILON * mag, fabs(y), est);
            double sn, cs;
            rc_sincos_tab(l0 * T, &sn, &cs);
            re = fma(w, cs, re);
            im = fma(-w, sn, im);
        }
    }
    re_out = re; im_out = im; *est_out = est;
}

// Whole evaluation on strided storage (one pad row in front of d, see ql_eigenvalues_strided).  On entry d[0..n), e[0..n-1) hold the matrix; xb the blocks outside
// [a, b] (already copied aside by the caller), pb the coupling product.  Returns true when the result is
// accepted; false = recompute with amplitude_strided (non-finite estimate, estimate above threshold, or
// eigenvalue non-convergence).  NaN / Inf input: accepted with NaN output, like amplitude_strided.
template <int LD = 0>
RC_HD bool amplitude_spectral_strided(double* d, double* e, int ld_rt, int n, double T, double pb, const SpecBlocks& xb,
                                      double& re_out, double& im_out) {
    const int ld = LD ? LD : ld_rt;
    double anorm = 0.0, chk = T;
    RC_AT(e, n - 1) = 0.0;
    for (int k = 0; k < n; ++k) {
        anorm = fmax(anorm, fabs(RC_AT(d, k)) + fabs(RC_AT(e, k)));
        chk += RC_AT(d, k) + RC_AT(e, k);
    }
    re_out = NAN; im_out = NAN;
    if (!(fabs(chk) <= DBL_MAX)) return true;
    const double tol = DBL_EPSILON * anorm;
    if (ql_eigenvalues_strided<LD>(d, e, ld, n, threshold_hi(tol), fmin(tol, 1e-280))) return false;
    double est;
    spectral_phase_sum<LD>(d, ld, n, T, pb, anorm, xb, re_out, im_out, &est);
    return est <= SPEC_EST_THR;   // false for NaN / inf as well
}

#undef RC_AT

// ---------------------------------------------------------------------------------------------
// Register-resident variant for the short chains (tuning builds, RC_REG_SPECTRAL): the solver of
// fidelity_reg_compact (pinned-end chase by default, block-at-0 form with RC_QL_PINNED_END=0), without the two
// eigenvector rows — 21 instead of 29 FP64 per rotation slot, half the shift moves, 2N fewer live
// registers — then the spectral weights from the eigenvalues parked in the lane's scratch row.
// scratch: >= 3N doubles per lane, element k at scratch[k * sstride]:
//   [0, N)     eigenvalues (written as they deflate)
//   [N, 2N)    original diagonal, [2N, 3N-1) original couplings (for the minors and for the recomputation)
// Returns false when the caller must recompute with fidelity_reg_compact (estimate rejected / QL failure).
// ---------------------------------------------------------------------------------------------
template <int N>
struct QlSweepValues {
    static RC_HD void run(double (&d)[N], double (&e)[N], int m, double dm, int mend, double tiny, int& rmin) {
        double g = wilkinson_g(d[0], d[1], e[0], dm);
        double r;
        double s = 1.0, c = 1.0, p = 0.0;
#pragma unroll
        for (int i = N - 2; i >= 0; --i) {
            if (i < m) {
                const double f = s * e[i];
                const double b = c * e[i];
                const double h = fma(f, f, fma(g, g, tiny));
                const double rinv = rc_rsqrt(h);
                r = h * rinv;
                e[i + 1] = r;
                rmin = hi_word(r) < rmin ? hi_word(r) : rmin;
                s = f * rinv;
                c = g * rinv;
                g = d[i + 1] - p;
                r = fma(d[i] - g, s, (2.0 * c) * b);
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
            }
        }
        d[0] -= p;
        e[0] = g;
        if (m != mend) {
#pragma unroll
            for (int i = 1; i < N; ++i)
                if (i == m) e[i] = 0.0;
        }
    }
};

#if RC_QL_PINNED_END
// QlChain of rc_ql.cuh without the eigenvector rows (21 FP64 per slot)
template <int N, int I>
struct QlChainValues {
    static RC_HD void run(double (&d)[N], double (&e)[N], double s, double c, int L, double tiny, int tolhi, int& rmin,
                          QlTail& t) {
        const double g = e[I + 1];
        const double f = s * e[I];
        const double b = c * e[I];
        const double h = fma(f, f, fma(g, g, tiny));
        const double rinv = rc_rsqrt(h);
        const double r = h * rinv;
        e[I + 1] = r;
        rmin = hi_word(r) < rmin ? hi_word(r) : rmin;
        s = f * rinv;
        c = g * rinv;
        const double gg = d[I + 1];
        const double r2 = fma(d[I] - gg, s, (2.0 * c) * b);
        const double p = s * r2;
        d[I + 1] = gg + p;
        e[I] = c * r2 - b;
        d[I] = d[I] - p;
        if constexpr (I == 0) {
            ql_tail<N, 0>(d, e, tolhi, t);
        } else {
            if (L == I) ql_tail<N, I>(d, e, tolhi, t);
            else QlChainValues<N, I - 1>::run(d, e, s, c, L, tiny, tolhi, rmin, t);
        }
    }
};
#endif

template <int N>
RC_HD bool amplitude_reg_spectral(double (&d)[N], double (&e)[N], int in, int out, double T, double* scratch, int sstride,
                                  double& re_out, double& im_out) {
#define RC_S(k) scratch[(size_t)(k) * sstride]
    double anorm = 0.0, chk = T;
    e[N - 1] = 0.0;
    const int a = in < out ? in : out, b = in < out ? out : in;
    double pb = 1.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        anorm = fmax(anorm, fabs(d[k]) + fabs(e[k]));
        chk += d[k] + e[k];
        RC_S(N + k) = d[k];
        if (k < N - 1) {
            RC_S(2 * N + k) = e[k];
            pb *= (k >= a && k < b) ? e[k] : 1.0;
        }
    }
    re_out = NAN; im_out = NAN;
    if (!(fabs(chk) <= DBL_MAX)) return true;     // NaN / Inf controller or draw (mcsim.py:369-374)
    const double tol = DBL_EPSILON * anorm;
    const int tolhi = threshold_hi(tol);
    const double tiny = fmin(tol, 1e-280);
#if RC_QL_PINNED_END
    // eigenvalues by the pinned-end chase of rc_ql.cuh (active block [L, N-1], nothing moves on deflation),
    // without the eigenvector rows; an evaluation with a negligible interior coupling is left to the caller's
    // recomputation (fidelity_reg_compact handles it), like a rejected estimate
    {
        int emin = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < N - 1; ++k) { const int h = hi_word(e[k]) & 0x7fffffff; emin = h < emin ? h : emin; }
        int L = emin < tolhi ? N : 0, it = 0;
        double g = wilkinson_g(d[0], d[1], e[0], d[N - 1]);
        while (L < N - 1 && it <= QL_MAX_SWEEPS) {
            ++it;
            int rmin = 0x7fffffff;
            QlTail t;
            {
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
                if constexpr (I == 0) {
                    ql_tail<N, 0>(d, e, tolhi, t);
                } else {
                    if (L == I) ql_tail<N, I>(d, e, tolhi, t);
                    else QlChainValues<N, I - 1>::run(d, e, s, c, L, tiny, tolhi, rmin, t);
                }
            }
            if (t.defl) { ++L; it = 0; }
            if (rmin < tolhi) L = N;
            g = wilkinson_g(t.dl, t.dl1, t.el, d[N - 1]);
        }
        if (L != N - 1) return false;
#pragma unroll
        for (int k = 0; k < N; ++k) RC_S(k) = d[k];
    }
#else
    int nact = N, ndone = 0, it = 0, rmin = 0;
    while (nact > 1 && it <= QL_MAX_SWEEPS) {
        if (negligible_hi(e[0], tolhi)) {
            RC_S(ndone) = d[0];
            ++ndone; --nact; it = 0;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) { d[i] = d[i + 1]; e[i] = e[i + 1]; }
        }
        if (nact > 1 && !negligible_hi(e[0], tolhi)) {
            int m = nact - 1;
            bool split = false;
            if (rmin < tolhi) {
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
            QlSweepValues<N>::run(d, e, m, dm, nact - 1, tiny, rmin);
            if (split) rmin = 0;
        }
    }
    if (nact > 1) return false;
    RC_S(ndone) = d[0];
#endif
    // weights: loops, not unrolled (the hot code of the register kernels has to stay inside the instruction cache)
    const double cgap = (double)(N - 1) * 4.0 * DBL_EPSILON * anorm;
    const int na = a, nb = N - 1 - b;
    double re = 0.0, im = 0.0, est = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < N; ++k) {
        const double lk = RC_S(k);
        double P0 = 1.0, P1 = 1.0;
        int mh = 0x7fffffff;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int j = 0; j < N; ++j) {
            const double df = lk - RC_S(j);
            const bool self = j == k;
            const int h = self ? 0x7fffffff : (hi_word(df) & 0x7fffffff);
            mh = h < mh ? h : mh;
            const double fct = self ? 1.0 : df;
            if (j & 1) P1 *= fct; else P0 *= fct;
        }
        const double Pk = P0 * P1;
        double num = pb, mag = fabs(pb);
        if (na > 0) {
            double p0 = 1.0, p1 = lk - RC_S(N), a0 = 1.0, a1 = fabs(p1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int j = 1; j < na; ++j) {
                const double x = lk - RC_S(N + j), b2 = RC_S(2 * N + j - 1) * RC_S(2 * N + j - 1);
                const double t = fma(x, p1, -(b2 * p0)), at = fma(fabs(x), a1, b2 * a0);
                p0 = p1; p1 = t; a0 = a1; a1 = at;
            }
            num *= p1; mag *= a1;
        }
        if (nb > 0) {
            double p0 = 1.0, p1 = lk - RC_S(N + N - 1), a0 = 1.0, a1 = fabs(p1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int j = N - 2; j > b; --j) {
                const double x = lk - RC_S(N + j), b2 = RC_S(2 * N + j) * RC_S(2 * N + j);
                const double t = fma(x, p1, -(b2 * p0)), at = fma(fabs(x), a1, b2 * a0);
                p0 = p1; p1 = t; a0 = a1; a1 = at;
            }
            num *= p1; mag *= a1;
        }
        const double y = rc_rcp_full(Pk);
        double w = num * y;
        w = fma(fma(-w, Pk, num), y, w);
        est = fma(fabs(w) * cgap, rc_rcp_seed(hi_as_double(mh)), est);
        est = fma(8.0 * DBL_EPSILON * mag, fabs(y), est);
        double sn, cs;
        rc_sincos_tab(lk * T, &sn, &cs);
        re = fma(w, cs, re);
        im = fma(-w, sn, im);
    }
    re_out = re; im_out = im;
    return est <= SPEC_EST_THR;
#undef RC_S
}

// Recomputation of a rejected evaluation with the eigenvector-accumulating solver, OUT OF LINE (its 4N live doubles
// must not set the register allocation of the spectral path): d / e are restored from the scratch row.
template <int N>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
double fidelity_reg_recompute(double* scratch, int sstride, int in, int out, double T, int* fail) {
    double d[N], e[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        d[k] = scratch[(size_t)(N + k) * sstride];
        e[k] = k < N - 1 ? scratch[(size_t)(2 * N + k) * sstride] : 0.0;
    }
    return fidelity_reg_compact<N>(d, e, in, out, T, scratch, sstride, fail);
}

template <int N>
RC_HD double fidelity_reg_spectral(double (&d)[N], double (&e)[N], int in, int out, double T, double* scratch, int sstride,
                                   int* fail, int* recomputed) {
    double re, im;
    *fail = 0;
    if (amplitude_reg_spectral<N>(d, e, in, out, T, scratch, sstride, re, im)) return fma(re, re, im * im);
    *recomputed = 1;
    return fidelity_reg_recompute<N>(scratch, sstride, in, out, T, fail);
}

}  // namespace rc
