// Development aid (NOT product, NOT used by tests): compiles rc_spectral.cuh for the host and compares the
// eigenvalue-only "spectral weights" evaluation with the eigenvector-accumulating QL of rc_ql.cuh.
//   g++ -O2 -I code-robchar_b200/csrc tools/spectral_sim.cpp -o /tmp/spectral_sim
//   /tmp/spectral_sim in.bin      in: int32 N,in,out,count ; then count*(N d, N-1 e, 1 T) doubles
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "rc_spectral.cuh"
using namespace rc;

template <int N>
static void run_reg(FILE* fi, int in, int out, int count) {
    std::vector<double> buf(2 * N);
    double maxdiff = 0;
    long nre = 0;
    for (int k = 0; k < count; ++k) {
        if (fread(buf.data(), sizeof(double), 2 * N, fi) != (size_t)(2 * N)) exit(1);
        double d[N], e[N], d2[N], e2[N], scr[3 * N], scr2[3 * N];
        for (int i = 0; i < N; ++i) { d[i] = d2[i] = buf[i]; e[i] = e2[i] = i < N - 1 ? buf[N + i] : 0.0; }
        const double T = buf[2 * N - 1];
        int fail = 0, re = 0;
        const double f = fidelity_reg_spectral<N>(d, e, in, out, T, scr, 1, &fail, &re);
        int fail2 = 0;
        const double fref = fidelity_reg_compact<N>(d2, e2, in, out, T, scr2, 1, &fail2);
        nre += re;
        const double diff = fabs(f - fref);
        if (!(diff <= maxdiff)) maxdiff = diff;
    }
    printf("REG N=%d %d->%d count=%d maxdiff=%.3e recomputed=%.5f\n", N, in, out, count, maxdiff, (double)nre / count);
}

int main(int argc, char** argv) {
    FILE* fi = fopen(argv[1], "rb");
    int hdr[4];
    if (!fi || fread(hdr, sizeof(int), 4, fi) != 4) return 1;
    const int N = hdr[0], in = hdr[1], out = hdr[2], count = hdr[3];
    if (argc > 2 && !strcmp(argv[2], "reg")) {
        switch (N) {
            case 4: run_reg<4>(fi, in, out, count); break;
            case 5: run_reg<5>(fi, in, out, count); break;
            case 6: run_reg<6>(fi, in, out, count); break;
            case 7: run_reg<7>(fi, in, out, count); break;
            case 8: run_reg<8>(fi, in, out, count); break;
            case 9: run_reg<9>(fi, in, out, count); break;
            case 10: run_reg<10>(fi, in, out, count); break;
            case 12: run_reg<12>(fi, in, out, count); break;
            case 13: run_reg<13>(fi, in, out, count); break;
            case 16: run_reg<16>(fi, in, out, count); break;
            default: fprintf(stderr, "reg variant: N = 4..10, 12, 13, 16\n"); return 2;
        }
        return 0;
    }
    std::vector<double> buf(2 * N);
    double maxdiff = 0, maxdiff_all = 0;
    long nfallback = 0, nbad = 0;
    const int a = std::min(in, out), b = std::max(in, out);
    for (int k = 0; k < count; ++k) {
        if (fread(buf.data(), sizeof(double), 2 * N, fi) != (size_t)(2 * N)) return 1;
        std::vector<double> d(buf.begin(), buf.begin() + N), e(N, 0.0), zi(N, 0.0), zo(N, 0.0);
        for (int i = 0; i < N - 1; ++i) e[i] = buf[N + i];
        const double T = buf[2 * N - 1];
        std::vector<double> de(2 * N + 1, 0.0), xd(N, 0.0), xe(N, 0.0);   // [pad][d][e] contiguous: the chase prefetches row -1
        double* d2p = de.data() + 1;
        double* e2p = d2p + N;
        for (int i = 0; i < N; ++i) { d2p[i] = d[i]; e2p[i] = e[i]; }
        SpecBlocks xb;
        xb.na = a; xb.nb = N - 1 - b; xb.xd = xd.data(); xb.xe = xe.data();
        for (int j = 0; j < a; ++j) { xd[j] = d[j]; xe[j] = e[j]; }
        for (int j = 0; j < xb.nb; ++j) { xd[a + j] = d[b + 1 + j]; xe[a + j] = e[b + 1 + j]; }
        double pb = 1.0;
        for (int i = a; i < b; ++i) pb *= e[i];
        double re, im;
        const bool ok = amplitude_spectral_strided(d2p, e2p, 1, N, T, pb, xb, re, im);
        zi[in] = 1; zo[out] = 1;
        int fail = 0;
        const double fref = fidelity_strided(d.data(), e.data(), zi.data(), zo.data(), 1, N, T, &fail);
        const double diff = fabs(re * re + im * im - fref);
        if (!(diff <= maxdiff_all)) maxdiff_all = diff;
        if (!ok) { ++nfallback; continue; }
        if (diff > maxdiff) maxdiff = diff;
        if (diff > 1e-11) ++nbad;
    }
    printf("N=%d %d->%d count=%d accepted: maxdiff=%.3e nbad(>1e-11)=%ld  fallback_rate=%.5f  (maxdiff incl. rejected %.3e)\n",
           N, in, out, count, maxdiff, nbad, (double)nfallback / count, maxdiff_all);
    return 0;
}
