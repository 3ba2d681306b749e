// Development micro-benchmark: dependent-chain latency and multi-warp throughput of FP64 ops on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int OP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) x[c] = fma(x[c], a, b);
            else if (OP == 1) x[c] = x[c] * a;
            else if (OP == 2) x[c] = x[c] + b;
            else if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[c])); x[c] = y + 1.0; }
            else if (OP == 4) { x[c] = (x[c] > b) ? x[c] * a : x[c] + a; }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS, int OP>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 4096;
    chain<CHAINS, OP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-7);
    chain<CHAINS, OP><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-7);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s chains=%d warps/SM=%2d : %.2f cycles per op per chain-step, %.2f cycles/warp-instr/SMSP\n", name, CHAINS, warps,
           (double)h / iters, (double)h / iters / CHAINS / ((warps + 3) / 4));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, 0>("DFMA", 1); run<2, 0>("DFMA", 1); run<4, 0>("DFMA", 1); run<8, 0>("DFMA", 1);
    run<1, 0>("DFMA", 4); run<1, 0>("DFMA", 16); run<1, 0>("DFMA", 32); run<4, 0>("DFMA", 16); run<8, 0>("DFMA", 32);
    run<1, 1>("DMUL", 1); run<1, 2>("DADD", 1); run<1, 3>("RSQ+DADD", 1); run<4, 3>("RSQ+DADD", 1); run<1, 4>("SETP+SEL", 1);
    return 0;
}
