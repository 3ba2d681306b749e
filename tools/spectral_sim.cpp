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

int main(int argc, char** argv) {
    FILE* fi = fopen(argv[1], "rb");
    int hdr[4];
    if (!fi || fread(hdr, sizeof(int), 4, fi) != 4) return 1;
    const int N = hdr[0], in = hdr[1], out = hdr[2], count = hdr[3];
    std::vector<double> buf(2 * N);
    double maxdiff = 0, maxdiff_all = 0;
    long nfallback = 0, nbad = 0;
    const int a = std::min(in, out), b = std::max(in, out);
    for (int k = 0; k < count; ++k) {
        if (fread(buf.data(), sizeof(double), 2 * N, fi) != (size_t)(2 * N)) return 1;
        std::vector<double> d(buf.begin(), buf.begin() + N), e(N, 0.0), zi(N, 0.0), zo(N, 0.0);
        for (int i = 0; i < N - 1; ++i) e[i] = buf[N + i];
        const double T = buf[2 * N - 1];
        std::vector<double> d2 = d, e2 = e, xd(N, 0.0), xe(N, 0.0);
        SpecBlocks xb;
        xb.na = a; xb.nb = N - 1 - b; xb.xd = xd.data(); xb.xe = xe.data();
        for (int j = 0; j < a; ++j) { xd[j] = d[j]; xe[j] = e[j]; }
        for (int j = 0; j < xb.nb; ++j) { xd[a + j] = d[b + 1 + j]; xe[a + j] = e[b + 1 + j]; }
        double pb = 1.0;
        for (int i = a; i < b; ++i) pb *= e[i];
        double re, im;
        const bool ok = amplitude_spectral_strided(d2.data(), e2.data(), 1, N, T, pb, xb, re, im);
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
