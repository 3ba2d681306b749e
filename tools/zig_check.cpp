// Development aid (NOT product): host build of the device noise generator (rc_philox.cuh) for a
// statistical check of the ziggurat without a GPU.
//   g++ -O2 -I code-robchar_b200/csrc tools/zig_check.cpp -o /tmp/zig_check && /tmp/zig_check 20000000
// Prints moments, a 256-bin equiprobable chi-square (df 255), tail counts against expectation, and the
// observed fast-path miss rate.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rc_philox.cuh"
#define RC_ZIG_QUAL static const
#include "rc_zig_table.inc"

int main(int argc, char** argv) {
    const long long nevals = argc > 1 ? atoll(argv[1]) : 1000000;
    const int nc = argc > 2 ? atoi(argv[2]) : 19;   // N = 7 complex model: 3N - 2 draws
    const uint32_t seed = argc > 3 ? (uint32_t)atoll(argv[3]) : 12345u;
    rc::ZigTables t{rc_zig_kw, rc_zig_y};
    const int NB = 256;
    std::vector<long long> bins(NB, 0);
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    long long n = 0, parked = 0, t3 = 0, t4 = 0, t45 = 0, t5 = 0;
    double lag = 0, prev = 0, mx = 0;
    std::vector<double> row(nc);
    for (long long ev = 0; ev < nevals; ++ev) {
        rc::NoiseKey key{seed, 678u, (uint32_t)(ev % 11), (uint64_t)(ev / 1100), (uint64_t)(ev % 100)};
        rc::normals_fill(key, nc, t, [&](int j) -> double& { return row[j]; });
        for (int j = 0; j < nc; ++j) {
            const double x = row[j];
            if (!(fabs(x) < 40.0)) { printf("bad value %g at ev %lld j %d\n", x, ev, j); return 1; }
            s1 += x; s2 += x * x; s3 += x * x * x; s4 += x * x * x * x;
            lag += x * prev; prev = x;
            const double ax = fabs(x);
            if (ax > mx) mx = ax;
            t3 += ax > 3; t4 += ax > 4; t45 += ax > 4.5; t5 += ax > 5;
            int b = (int)(0.5 * erfc(-x / sqrt(2.0)) * NB);
            if (b >= NB) b = NB - 1;
            bins[b]++; ++n;
        }
    }
    // miss rate measured directly on the fast path
    long long tries = 0;
    for (uint32_t p = 0; p < 2000000; ++p) {
        rc::NoiseKey key{1u, 2u, 3u, p, 7u};
        rc::Philox4 r = rc::philox_block(key, 0);
        bool miss; rc::zig_try(r.x, r.y, rc_zig_kw, &miss); parked += miss; ++tries;
    }
    double chi = 0, e = (double)n / NB;
    for (int b = 0; b < NB; ++b) chi += (bins[b] - e) * (bins[b] - e) / e;
    const double m = s1 / n, var = s2 / n - m * m;
    printf("{\"n\": %lld, \"mean\": %.6e, \"var\": %.8f, \"skew\": %.6e, \"kurt\": %.6f, \"lag1\": %.6e, \"chi2_255\": %.2f, "
           "\"max\": %.4f, \"gt3\": %lld, \"exp3\": %.1f, \"gt4\": %lld, \"exp4\": %.1f, \"gt45\": %lld, \"exp45\": %.2f, "
           "\"gt5\": %lld, \"exp5\": %.3f, \"miss_rate\": %.6f}\n",
           n, m, var, s3 / n, s4 / n, lag / n, chi, mx, t3, n * erfc(3 / sqrt(2.0)), t4, n * erfc(4 / sqrt(2.0)), t45,
           n * erfc(4.5 / sqrt(2.0)), t5, n * erfc(5 / sqrt(2.0)), (double)parked / tries);
    return 0;
}
