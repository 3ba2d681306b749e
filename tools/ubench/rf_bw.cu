// Development micro-benchmark: does operand (register-file) bandwidth limit FP64 issue when all three
// DFMA sources are distinct registers (no operand reuse)?  Compare with the 2-cycle pipe occupancy.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = 0.999 + i * 1e-4; c[i] = 1e-3 * i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b[i], c[i]);                 // 3 distinct 64-bit sources
            if (MODE == 1) a[i] = fma(a[i], b[0], c[0]);                 // two shared sources (reuse cache)
            if (MODE == 2) a[i] = a[i] * b[i];                           // 2 distinct sources
            if (MODE == 3) { a[i] = fma(a[i], b[i], c[i]); b[i] = b[i] + c[(i + 1) & 7]; }   // DFMA + DADD
            if (MODE == 4) { a[i] = fma(a[i], b[i], c[i]); int v = __double2hiint(b[i]); v = v * 3 + 1;
                             b[i] = __hiloint2double(v & 0x3fffffff | 0x3ff00000, __double2loint(b[i])); }  // DFMA + int
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, int fp64_per_iter) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 2048;
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters);
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM=%2d : %.2f cycles per FP64 warp-instr per SMSP\n", name, warps, (double)h / iters / fp64_per_iter / (warps / 4));
}

int main() {
    run<0>("DFMA 3 distinct", 16, 8); run<0>("DFMA 3 distinct", 32, 8);
    run<1>("DFMA shared b,c", 16, 8);
    run<2>("DMUL 2 distinct", 16, 8);
    run<3>("DFMA + DADD", 16, 16); run<3>("DFMA + DADD", 32, 16);
    run<4>("DFMA + int ops", 16, 8); run<4>("DFMA + int ops", 32, 8);
    return 0;
}
