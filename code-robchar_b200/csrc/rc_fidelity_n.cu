// One translation unit per chain length: compiled with -DRC_NSPIN=<N> for N = 2..REG_MAX_N so the
// fully unrolled register-resident eigensolvers build in parallel.
#include <stdlib.h>
#include "rc_fidelity.cuh"

#ifndef RC_NSPIN
#error "compile with -DRC_NSPIN=<N>"
#endif
#define RC_CAT2(a, b) a##b
#define RC_CAT(a, b) RC_CAT2(a, b)

namespace rc {

template <int MODEL, bool REPLAY>
static cudaError_t launch_reg(const FidArgs& a, int sm_count, cudaStream_t st) {
    constexpr int N = RC_NSPIN;
    const int threads = reg_threads_runtime(N, REPLAY);
    // one private row per lane (+ the ziggurat fast-path table in Philox mode)
    size_t smem = (size_t)threads * reg_row_doubles(MODEL, N) * sizeof(double) + (REPLAY ? 0 : sizeof(ZigEntry) * ZIG_LAYERS);
    auto kern = fidelity_reg_kernel<N, MODEL, REPLAY>;
    if (const char* e = getenv("RC_FID_SMEM_PAD")) smem += (size_t)atoi(e) * 1024;  // tuning: limits CTAs/SM
    cudaError_t err;
    if (smem > 40 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long total = (long long)a.S * a.C * a.B;
    long long ntiles = (total + threads - 1) / threads;
    long long grid = (long long)sm_count * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, threads, smem, st>>>(a); rc::note_launch();
    return cudaGetLastError();
}

// Philox mode: warp-autonomous items (no CTA barriers), same CTA shape and shared-memory layout as the plain kernel.
template <int MODEL>
static cudaError_t launch_fused_reg_warp(const FusedArgs& g, int sm_count, cudaStream_t st) {
    constexpr int N = RC_NSPIN;
    const int threads = reg_threads_runtime(N, false);
    size_t smem = (size_t)threads * (reg_row_doubles(MODEL, N) + (fused_lane_acc(MODEL, N) ? LACC_DOUBLES : 0)) * sizeof(double) +
                  sizeof(ZigEntry) * ZIG_LAYERS;
    auto kern = fidelity_stats_reg_warp_kernel<N, MODEL>;
    cudaError_t err;
    if (smem > 40 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    const long long wpc = threads / 32;
    const long long nitems = (long long)g.f.S * g.f.C * g.nchunks;
    long long grid = (long long)sm_count * occ;
    const long long need = (nitems + wpc - 1) / wpc;
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, threads, smem, st>>>(g); rc::note_launch();
    return cudaGetLastError();
}

template <int MODEL, bool REPLAY>
static cudaError_t launch_fused_reg(const FusedArgs& g, int threads, int sm_count, cudaStream_t st) {
    constexpr int N = RC_NSPIN;
    size_t smem = (size_t)threads * reg_row_doubles(MODEL, N) * sizeof(double) + (REPLAY ? 0 : sizeof(ZigEntry) * ZIG_LAYERS);
    auto kern = fidelity_stats_reg_kernel<N, MODEL, REPLAY>;
    cudaError_t err;
    if (smem > 40 * 1024) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long nitems = (long long)g.f.S * g.f.C * g.nchunks;
    long long grid = (long long)sm_count * occ;
    if (grid > nitems) grid = nitems;
    if (grid < 1) return cudaSuccess;
    kern<<<(unsigned)grid, threads, smem, st>>>(g); rc::note_launch();
    return cudaGetLastError();
}

cudaError_t RC_CAT(launch_fused_reg_, RC_NSPIN)(const FusedArgs& g, int threads, int sm_count, cudaStream_t st) {
    const bool replay = g.f.replay != nullptr;
    if (!replay)
        return g.f.model == MODEL_COMPLEX3 ? launch_fused_reg_warp<MODEL_COMPLEX3>(g, sm_count, st)
                                           : launch_fused_reg_warp<MODEL_REAL2>(g, sm_count, st);
    return g.f.model == MODEL_COMPLEX3 ? launch_fused_reg<MODEL_COMPLEX3, true>(g, threads, sm_count, st)
                                       : launch_fused_reg<MODEL_REAL2, true>(g, threads, sm_count, st);
}

cudaError_t RC_CAT(launch_fid_reg_, RC_NSPIN)(const FidArgs& a, int sm_count, cudaStream_t st) {
    const bool replay = a.replay != nullptr;
    if (a.model == MODEL_COMPLEX3)
        return replay ? launch_reg<MODEL_COMPLEX3, true>(a, sm_count, st) : launch_reg<MODEL_COMPLEX3, false>(a, sm_count, st);
    return replay ? launch_reg<MODEL_REAL2, true>(a, sm_count, st) : launch_reg<MODEL_REAL2, false>(a, sm_count, st);
}

}  // namespace rc
