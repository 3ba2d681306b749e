// Shared pieces of the statistics stage (sort-based rc_stats and the streaming fused path).
#pragma once
#include "rc_ql.cuh"

namespace rc {

// Row order of the [15][nseg] statistics tensor == key order of the reference's .mcm dict
// (mcsim.py:178-183 x {"", " upper", " lower"}, mcsim.py:496-498).
enum StatRow { ST_W = 0, ST_Q95 = 3, ST_Q98 = 6, ST_STD = 9, ST_WC = 12, ST_ROWS = 15 };

// np.clip(x, 0, 1) with NaN propagation (mcsim.py:484-485).
RC_HD double clip01(double x) { return (x != x) ? x : fmin(fmax(x, 0.0), 1.0); }

// Streaming partial of one chunk of a segment: count, then per variant (centre, upper, lower)
// mean / M2 (Chan et al. pairwise merge), sum(1 - v), #v>=0.95, #v>=0.98; and the raw minimum.
constexpr int PART_DOUBLES = 17;
struct Moments {
    double n, mean[3], m2[3], s1[3], c95[3], c98[3], mn;
};

RC_HD void moments_init(Moments& m) {
    m.n = 0.0; m.mn = INFINITY;
    for (int k = 0; k < 3; ++k) { m.mean[k] = 0.0; m.m2[k] = 0.0; m.s1[k] = 0.0; m.c95[k] = 0.0; m.c98[k] = 0.0; }
}

// a <- a merged with b (order matters only at rounding level; callers use a fixed tree).
RC_HD void moments_merge(Moments& a, const Moments& b) {
    if (b.n == 0.0) return;
    if (a.n == 0.0) { a = b; return; }
    const double n = a.n + b.n;
    for (int k = 0; k < 3; ++k) {
        const double dlt = b.mean[k] - a.mean[k];
        a.mean[k] = a.mean[k] + dlt * (b.n / n);
        a.m2[k] = a.m2[k] + b.m2[k] + dlt * dlt * (a.n * b.n / n);
        a.s1[k] += b.s1[k];
        a.c95[k] += b.c95[k];
        a.c98[k] += b.c98[k];
    }
    a.mn = (a.mn != a.mn || b.mn != b.mn) ? NAN : fmin(a.mn, b.mn);
    a.n = n;
}

}  // namespace rc
