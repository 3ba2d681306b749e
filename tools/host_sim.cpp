// Development aid (NOT product, NOT used by tests): compiles the device eigensolver header for the
// host so numerics and sweep statistics can be studied without a GPU.
//   g++ -O2 -DRC_QL_STATS -I code-robchar_b200/csrc tools/host_sim.cpp -o /tmp/host_sim
//   /tmp/host_sim in.bin out.bin     in: int32 N,in,out,count ; then count*(N d, N-1 e, 1 T) doubles
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "rc_ql.cuh"

template <int N>
void run(FILE* fi, FILE* fo, int in, int out, int count, int strided) {
    std::vector<double> buf(2 * N);
    long hist[64] = {0}; long totsweeps = 0, totrot = 0, irregular = 0; int maxsweeps = 0; long perl[64] = {0};
    for (int k = 0; k < count; ++k) {
        if (fread(buf.data(), sizeof(double), 2 * N, fi) != (size_t)(2 * N)) { fprintf(stderr, "short read\n"); exit(1); }
        double d[N], e[N], T = buf[2 * N - 1];
        for (int i = 0; i < N; ++i) d[i] = buf[i];
        for (int i = 0; i < N - 1; ++i) e[i] = buf[N + i];
        e[N - 1] = 0;
        int fail = 0;
        double f;
        if (strided == 1) {
            double zi[N], zo[N];
            for (int i = 0; i < N; ++i) { zi[i] = (i == in); zo[i] = (i == out); }
            f = rc::fidelity_strided(d, e, zi, zo, 1, N, T, &fail);
        } else {
            rc::QlStats st; memset(&st, 0, sizeof st);
            if (strided == 2) { double scr[2 * N]; f = rc::fidelity_reg_compact<N>(d, e, in, out, T, scr, 1, &fail, &st); }
            else f = rc::fidelity_reg<N>(d, e, in, out, T, &fail, &st);
            totsweeps += st.total_sweeps; totrot += st.rotations; irregular += st.irregular;
            if (st.total_sweeps > maxsweeps) maxsweeps = st.total_sweeps;
            hist[st.total_sweeps < 63 ? st.total_sweeps : 63]++;
            for (int l = 0; l < N; ++l) perl[l] += st.sweeps_per_l[l];
        }
        if (fail) fprintf(stderr, "nonconvergence at %d\n", k);
        fwrite(&f, sizeof(double), 1, fo);
    }
    if (strided != 1) {
        fprintf(stderr, "N=%d mean sweeps %.3f max %d mean rotations %.2f irregular %ld of %d\n per-l mean:", N, (double)totsweeps / count, maxsweeps, (double)totrot / count, irregular, count);
        for (int l = 0; l < N; ++l) fprintf(stderr, " %.2f", (double)perl[l] / count);
        fprintf(stderr, "\n");
    }
}

int main(int argc, char** argv) {
    FILE* fi = fopen(argv[1], "rb"); FILE* fo = fopen(argv[2], "wb");
    int strided = argc > 3 ? atoi(argv[3]) : 0;
    int hdr[4];
    if (fread(hdr, sizeof(int), 4, fi) != 4) return 1;
    int N = hdr[0], in = hdr[1], out = hdr[2], count = hdr[3];
    switch (N) {
#define C(n) case n: run<n>(fi, fo, in, out, count, strided); break;
        C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(10) C(16) C(32)
        default: fprintf(stderr, "N unsupported\n"); return 2;
    }
    fclose(fo);
    return 0;
}
