// Evolution + fidelity kernels: one lane per (sigma level, controller, noise draw).
//
// Reference path replaced: MCDataSim.get_algo_fid_dist's triple loop (mcsim.py:422-460) around
// structured_perturbation.evaluate_noisy_fidelity (noise_model.py:98-147), and the optimiser-side
// variant LBFGS.structured_perturabation / fidelity_ss (qnewton.py:366-423).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rc_ql.cuh"
#include "rc_spectral.cuh"
#include "rc_philox.cuh"
#include "rc_stats.cuh"

namespace rc {

constexpr int MODEL_COMPLEX3 = 0;  // noise_model.py:122-147: (z_ii, nn_i, nn2_i) per site
constexpr int MODEL_REAL2 = 1;     // qnewton.py:366-379:     (z_ii, nn_i) per site
constexpr int REG_MAX_N = 16;      // register-resident eigensolver compiled up to here
constexpr int REG_DEFAULT_N = 12;  // measured crossover on B200 (round 2, pinned-end QL): registers win for N <= 12, shared memory above
constexpr int MAX_N = 32;

struct FidArgs {
    const double* ctrl;     // [C][N+1] biases then time (device)
    const double* sigma;    // [S] simulation noise levels (device)
    const double* replay;   // [S][C][B][K] standard normals in reference draw order, or nullptr (Philox)
    double* fids;           // [S][C][B]
    unsigned long long* nonconv;  // device counter of QL non-convergences (may be nullptr)
    unsigned long long* respec;   // device counter of spectral-path evaluations recomputed with eigenvector rows (may be nullptr)
    long long C, B;
    int S, N, in, out, model, zz;
    uint32_t seed_lo, seed_hi;
    long long c_offset, b_offset;  // global index of this shard's first controller / draw (Philox counters)
    ZigTables zig;                 // ziggurat tables in global memory (Philox mode), see zig_tables_device()
    int s_offset;                  // global index of this launch's first sigma level (sigma-chunked host sweep)
    double* amps;                  // optional [S][C][B][2]: complex transfer amplitude (re, im) per evaluation (served by
                                   // the shared-memory kernel's AMPS instantiation for every N; small batches only).
                                   // The Hamiltonian must be real symmetric (RC_MODEL_REAL2): the complex model's
                                   // phase gauge leaves |amp|^2 invariant but not amp
};

__host__ __device__ constexpr int draws_per_site(int model) { return model == MODEL_COMPLEX3 ? 3 : 2; }

// Algorithm of the register-resident family: 0 = QL accumulating the in / out eigenvector rows (default),
// 1 = eigenvalues + spectral weights (rc_spectral.cuh, amplitude_reg_spectral) with the former as out-of-line
// recomputation.  Measured on B200 twice (kernel only, 2.0e7 evaluations), and the spectral variant LOSES at every
// short chain both times.  Block-at-0 solvers (first half of round 2): N=4 10.2e9 vs 11.0e9, N=5 7.14 vs 7.54,
// N=6 5.37 vs 5.70, N=7 4.18 vs 4.29, N=8 3.40 vs 3.41 evals/s.  Pinned-end solvers (second half): N=4 10.8e9 vs
// 11.9e9, N=5 7.67 vs 8.49, N=6 5.96 vs 6.56, N=7 4.67 vs 5.09, N=8 3.85 vs 4.06, N=9 3.04 vs 3.15, N=10 2.62 vs
// 2.67: the N(N-1) weight products and the per-eigenvalue reciprocal / error estimate (kept as loops to protect the
// instruction cache) cost more than the 8 FP64 per rotation they save, and above N = 12 the 3N-double shared-memory
// row caps the CTA at 512 lanes, where the eigenvector form fits as well.  Kept for tuning builds only.
#ifndef RC_REG_SPECTRAL
#define RC_REG_SPECTRAL 0
#endif
// Doubles of the private shared-memory row of one lane in the register kernels: the K normals of the evaluation,
// reused as eigensolver scratch (2N for the eigenvector solver, 3N - 1 for the spectral one); odd => conflict-free.
__host__ __device__ constexpr int reg_row_doubles(int model, int n) {
    return (RC_REG_SPECTRAL && draws_per_site(model) * n < 3 * n ? 3 * n : draws_per_site(model) * n) | 1;
}

// Heisenberg / Z diagonal of qnewton.py:148-150 for the open chain: t_i = (N-1)/2 - deg_i.
RC_HD double zz_diag(int i, int n) { return 0.5 * (n - 1) - ((i == 0 || i == n - 1) ? 1.0 : 2.0); }

struct EvalIndex { long long s, c, b; };
RC_HD EvalIndex decode_eval(long long e, long long C, long long B) {
    EvalIndex r;
    long long sc = e / B;
    r.b = e - sc * B;
    r.s = sc / C;
    r.c = sc - r.s * C;
    return r;
}

#if defined(__CUDACC__)
// d/e of the gauge-transformed real symmetric tridiagonal Hamiltonian from one row of draws.
// get(j) returns the j-th STANDARD normal of this evaluation in reference draw order.
template <int N, int MODEL, class Get>
__device__ __forceinline__ void build_tridiagonal(const double* __restrict__ x, double sigma, int zz, Get&& get,
                                                  double (&d)[N], double (&e)[N]) {
    constexpr int P = draws_per_site(MODEL);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double base = zz ? zz_diag(i, N) : 0.0;
        // reference order: (HH_ii + z_ii) + x_i  with z_ii = sigma * normal  (noise_model.py:100-104)
        d[i] = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, get(P * i))), x[i]);
        if (i >= 1) {
            double a = __dadd_rn(1.0, __dmul_rn(sigma, get(P * i + 1)));
            if (MODEL == MODEL_COMPLEX3) {
                double bb = __dmul_rn(sigma, get(P * i + 2));
                e[i - 1] = rc_sqrt(fma(a, a, bb * bb));  // |1 + nn + i nn2|
            } else {
                e[i - 1] = a;
            }
        }
    }
    e[N - 1] = 0.0;
}

RC_HD NoiseKey noise_key(const FidArgs& a, long long s, long long c, long long b) {
    NoiseKey k;
    k.seed_lo = a.seed_lo; k.seed_hi = a.seed_hi; k.sidx = (uint32_t)(s + a.s_offset);
    k.cidx = (uint64_t)(c + a.c_offset); k.bidx = (uint64_t)(b + a.b_offset);
    return k;
}

// Philox mode: this lane writes the standard normals of its evaluation into its own staged row, in
// reference draw order (the discarded site-0 coupling slots are left untouched and never read).
// `kw` = the CTA's shared-memory copy of the ziggurat fast-path table.
template <int N, int MODEL>
__device__ __forceinline__ void philox_fill_row(const FidArgs& a, long long s, long long c, long long b, double* row,
                                                const ZigEntry* kw) {
    constexpr int P = draws_per_site(MODEL);
    constexpr int NC = P * N - (P - 1);  // compact count: reference order minus discarded draws
    ZigTables t;
    t.kw = kw; t.y = a.zig.y;
    normals_fill(noise_key(a, s, c, b), NC, t, [&](int jc) -> double& { return row[jc == 0 ? 0 : jc + (P - 1)]; });
}

// Cooperative copy of the 16 KB fast-path table into shared memory (once per persistent CTA).
__device__ __forceinline__ void load_zig_table(const ZigEntry* __restrict__ g, ZigEntry* s) {
    const double2* src = reinterpret_cast<const double2*>(g);
    double2* dst = reinterpret_cast<double2*>(s);
    for (int i = threadIdx.x; i < ZIG_LAYERS; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// Coalesced staging of `nvalid` consecutive replay rows (K doubles each) into padded shared rows.
template <int K, int KP>
__device__ __forceinline__ void stage_replay_rows(const double* __restrict__ src, int nvalid, double* stage) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < nvalid * K; idx += blockDim.x) {
        int row = idx / K, j = idx - row * K;
        stage[row * KP + j] = __ldcs(src + idx);
    }
    __syncthreads();
}

// One evaluation, register resident.  `row` = this lane's private shared-memory row (K|1 doubles):
// holds the standard normals on entry (staged replay or Philox) and is reused as the scratch of the
// eigensolver (reg_row_doubles).
template <int N, int MODEL, bool REPLAY>
__device__ __forceinline__ double eval_reg(const FidArgs& a, long long s, long long c, long long b, double* row,
                                           const ZigEntry* kw) {
    const double* x = a.ctrl + c * (N + 1);
    double xr[N + 1];
#pragma unroll
    for (int i = 0; i <= N; ++i) xr[i] = __ldg(x + i);
    const double sigma = __ldg(a.sigma + s);
    if (!REPLAY) philox_fill_row<N, MODEL>(a, s, c, b, row, kw);
    double d[N], ee[N];
    build_tridiagonal<N, MODEL>(xr, sigma, a.zz, [&](int j) { return row[j]; }, d, ee);
    int fail = 0;
#if RC_REG_SPECTRAL
    int recomputed = 0;
    double f = fidelity_reg_spectral<N>(d, ee, a.in, a.out, fabs(xr[N]), row, 1, &fail, &recomputed);
    if (recomputed && a.respec) atomicAdd(a.respec, 1ull);
#else
    double f = fidelity_reg_compact<N>(d, ee, a.in, a.out, fabs(xr[N]), row, 1, &fail);
#endif
    if (fail && a.nonconv) atomicAdd(a.nonconv, 1ull);
    return f;
}

// CTA shape of the register-resident kernels, measured on B200 at N = 4..8 (profiles/README.md, r01h):
// Philox mode runs fastest as ONE 24-warp CTA per SM with the compiler held to 80 registers (a few
// spilled doubles): 3.58e9 evaluations/s at N=7 against 3.1e9 for four 4-warp CTAs at 126 registers;
// warp counts that are not a multiple of the four schedulers lose 5-10 %; the short chains have registers to
// spare and run 2 % faster with 32 (N <= 4) / 28 (N = 5) warps (r01k).  Replay mode stages its
// rows behind CTA barriers and prefers three 8-warp CTAs.  N > 8 (non-default, RC_REG_MAX_N) keeps
// 4-warp CTAs: the eigensolver state alone needs more than 80 registers there.
// RC_REG_THREADS / RC_REG_MIN_BLOCKS (compile time) and RC_FID_THREADS (environment) are tuning overrides.
#ifndef RC_REG_MIN_BLOCKS
#define RC_REG_MIN_BLOCKS 0
#endif
#ifndef RC_REG_THREADS
#define RC_REG_THREADS 0
#endif
#ifndef RC_TILE_SYNC
#define RC_TILE_SYNC 0
#endif
__host__ __device__ constexpr int reg_cta_threads(int n, bool replay) {
    return RC_REG_THREADS ? RC_REG_THREADS
                          : (n > REG_DEFAULT_N ? 128 : (n > 8 ? (replay ? 256 : 512) : (replay ? 256 : (n <= 4 ? 1024 : (n == 5 ? 896 : 768)))));
}
__host__ __device__ constexpr int reg_cta_min_blocks(int n, bool replay) {
    return RC_REG_MIN_BLOCKS ? RC_REG_MIN_BLOCKS : (n > REG_DEFAULT_N ? 1 : (n > 8 ? (replay ? 2 : 1) : (replay ? 3 : 1)));
}
constexpr int MAX_CTA_WARPS = 32;
constexpr int SMEM_MAX_THREADS = 768;         // launch bound of the shared-memory evolution kernel
constexpr int SMEM_WIDE_THREADS = 512;        // ... of its long-chain instantiation (N >= 23: fewer lanes fit, 128 registers each)
constexpr int SMEM_FUSED_MAX_THREADS = 512;   // ... and of its fused-statistics variant (more live registers)

template <int N, int MODEL, bool REPLAY>
__global__ void __launch_bounds__(reg_cta_threads(N, REPLAY), reg_cta_min_blocks(N, REPLAY)) fidelity_reg_kernel(FidArgs a) {
    constexpr int K = draws_per_site(MODEL) * N;
    constexpr int KP = reg_row_doubles(MODEL, N);
    extern __shared__ __align__(16) double smem_raw[];
    // Philox mode: [ziggurat table 16 KB][one row per lane]; replay mode: rows only
    const ZigEntry* kw = reinterpret_cast<const ZigEntry*>(smem_raw);
    double* stage = smem_raw + (REPLAY ? 0 : 2 * ZIG_LAYERS);
    if (!REPLAY) load_zig_table(a.zig.kw, reinterpret_cast<ZigEntry*>(smem_raw));
    const long long total = (long long)a.S * a.C * a.B;
    const long long ntiles = (total + blockDim.x - 1) / blockDim.x;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e0 = tile * blockDim.x;
        const long long e = e0 + threadIdx.x;
        if (REPLAY) {
            long long nvalid = total - e0 < (long long)blockDim.x ? total - e0 : (long long)blockDim.x;
            stage_replay_rows<K, KP>(a.replay + e0 * K, (int)nvalid, stage);
        }
#if RC_TILE_SYNC
        if (!REPLAY) __syncthreads();   // keeps the CTA's warps in the same phase of the code
#endif
        if (e >= total) continue;
        EvalIndex ix = decode_eval(e, a.C, a.B);
        a.fids[e] = eval_reg<N, MODEL, REPLAY>(a, ix.s, ix.c, ix.b, stage + threadIdx.x * KP, kw);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused evolution + streaming statistics (no fidelity tensor).  Work item = (segment, chunk of
// `chunk` draws); the CTA writes one Moments partial per item; a finalize kernel merges the
// partials of each segment in chunk order (deterministic).
// ---------------------------------------------------------------------------------------------
struct FusedArgs {
    FidArgs f;
    double eps;          // DKW shift
    long long chunk;     // draws per work item (multiple of blockDim)
    long long nchunks;   // chunks per segment
    double* partials;    // [S*C][nchunks][PART_DOUBLES]
};

__device__ __forceinline__ void moments_add(Moments& m, double f, double eps, double (&shift)[3], bool first) {
    const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (first) shift[k] = v[k];
        const double y = v[k] - shift[k];   // shifted sums: mean[] holds sum(y), m2[] holds sum(y^2) until finish
        m.mean[k] += y;
        m.m2[k] = fma(y, y, m.m2[k]);
        m.s1[k] += 1.0 - v[k];
        m.c95[k] += (v[k] >= 0.95) ? 1.0 : 0.0;
        m.c98[k] += (v[k] >= 0.98) ? 1.0 : 0.0;
    }
    m.mn = (f != f || m.mn != m.mn) ? NAN : fmin(m.mn, f);
    m.n += 1.0;
}

__device__ __forceinline__ void moments_finish_thread(Moments& m, const double (&shift)[3]) {
    if (m.n == 0.0) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sy = m.mean[k], syy = m.m2[k];
        m.mean[k] = shift[k] + sy / m.n;
        m.m2[k] = fmax(syy - sy * sy / m.n, 0.0);
        if (sy != sy || syy != syy) { m.mean[k] = NAN; m.m2[k] = NAN; }
    }
}

__device__ __forceinline__ Moments moments_shfl_down(const Moments& m, int o) {
    Moments r;
    r.n = __shfl_down_sync(0xffffffffu, m.n, o);
    r.mn = __shfl_down_sync(0xffffffffu, m.mn, o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.mean[k] = __shfl_down_sync(0xffffffffu, m.mean[k], o);
        r.m2[k] = __shfl_down_sync(0xffffffffu, m.m2[k], o);
        r.s1[k] = __shfl_down_sync(0xffffffffu, m.s1[k], o);
        r.c95[k] = __shfl_down_sync(0xffffffffu, m.c95[k], o);
        r.c98[k] = __shfl_down_sync(0xffffffffu, m.c98[k], o);
    }
    return r;
}

__device__ __forceinline__ void moments_store(const Moments& m, double* p) {
    p[0] = m.n; p[16] = m.mn;
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[1 + k] = m.mean[k]; p[4 + k] = m.m2[k]; p[7 + k] = m.s1[k]; p[10 + k] = m.c95[k]; p[13 + k] = m.c98[k]; }
}
__device__ __forceinline__ void moments_load(Moments& m, const double* p) {
    m.n = p[0]; m.mn = p[16];
#pragma unroll
    for (int k = 0; k < 3; ++k) { m.mean[k] = p[1 + k]; m.m2[k] = p[4 + k]; m.s1[k] = p[7 + k]; m.c95[k] = p[10 + k]; m.c98[k] = p[13 + k]; }
}

// Fixed shuffle tree over the warp; result valid in lane 0.
__device__ __forceinline__ void moments_warp_merge(Moments& m) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        Moments other = moments_shfl_down(m, o);
        if (lane + o < 32) moments_merge(m, other);
    }
}

// CTA-wide deterministic merge of per-thread Moments; result valid in thread 0.
__device__ __forceinline__ void moments_block_merge(Moments& m, double* scratch /* [nwarp][PART_DOUBLES] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Moments other = moments_shfl_down(m, o);
        if (lane + o < 32) moments_merge(m, other);
    }
    __syncthreads();
    if (lane == 0) moments_store(m, scratch + warp * PART_DOUBLES);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nwarp; ++w) {
            Moments o2;
            moments_load(o2, scratch + w * PART_DOUBLES);
            moments_merge(m, o2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous fused path (Philox mode).  Work item = (segment, 512 draws) processed by ONE warp in 16 passes: no CTA barrier anywhere (a barrier per item keeps re-aligning the 24 warps of the CTA
// to the slowest one).  Nothing statistical stays live in registers across an evaluation (the eigensolver
// already fills the register budget; 20 extra live registers cost more in spills than the statistics cost in
// instructions): after every pass the warp reduces its 32 fidelities by shuffles / ballots and lane 0 adds the
// pass totals to the warp's 16-double accumulator in shared memory.  Sums are taken about a shift (the item's
// first sample) so the variance does not cancel; at the end of the item the accumulator becomes a Moments
// partial; fused_finalize_kernel merges the partials of a segment in chunk order as before.  Fixed order
// everywhere => deterministic.  (Tried instead: per-lane running sums in private shared-memory slots behind the
// lane's row, warp merge once per item — +2-3 % at N <= 6 but -2 % at N = 7 and -16 % at N = 8: the larger
// shared-memory carve-out leaves too little L1 for the kernel's register spills.)
// ---------------------------------------------------------------------------------------------
constexpr int WACC_DOUBLES = 16;   // sy[3] syy[3] c95[3] c98[3] | mn shift n nan
enum WaccSlot { WA_SY = 0, WA_SYY = 3, WA_C95 = 6, WA_C98 = 9, WA_MN = 12, WA_SHIFT = 13, WA_N = 14, WA_NAN = 15 };

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// All 32 lanes call, converged; `valid` lanes hold a fidelity.  `first`: first pass of the item (lane 0 is valid).
__device__ __forceinline__ void warp_acc_pass(double* __restrict__ wacc, bool first, bool valid, double f, double eps) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (first) {
        const double f0 = __shfl_sync(0xffffffffu, f, 0);
        if (lane < WACC_DOUBLES) wacc[lane] = lane == WA_MN ? INFINITY : (lane == WA_SHIFT ? f0 : 0.0);
        __syncwarp();
    }
    const double shift = wacc[WA_SHIFT];
    const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
    const double sh[3] = {shift, clip01(shift - eps), clip01(shift + eps)};
    double tot[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double y = valid ? v[k] - sh[k] : 0.0;
        tot[WA_SY + k] = warp_sum_f64(y);
        tot[WA_SYY + k] = warp_sum_f64(y * y);
        tot[WA_C95 + k] = (double)__popc(__ballot_sync(0xffffffffu, valid && v[k] >= 0.95));
        tot[WA_C98 + k] = (double)__popc(__ballot_sync(0xffffffffu, valid && v[k] >= 0.98));
    }
    double mn = valid ? f : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned nanmask = __ballot_sync(0xffffffffu, valid && f != f);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 12; ++j) wacc[j] += tot[j];
        wacc[WA_MN] = fmin(wacc[WA_MN], mn);
        wacc[WA_N] += (double)__popc(vmask);
        wacc[WA_NAN] += (double)__popc(nanmask);
    }
}

// Lane 0: the warp accumulator of a finished item as a Moments partial.
__device__ __forceinline__ void warp_acc_finish(const double* __restrict__ wacc, double eps, Moments& m) {
    moments_init(m);
    const double n = wacc[WA_N];
    if (n == 0.0) return;
    const double shift = wacc[WA_SHIFT];
    const double sh[3] = {shift, clip01(shift - eps), clip01(shift + eps)};
    const bool nan = wacc[WA_NAN] != 0.0;
    m.n = n;
    m.mn = nan ? NAN : wacc[WA_MN];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sy = wacc[WA_SY + k], syy = wacc[WA_SYY + k];
        m.mean[k] = sh[k] + sy / n;
        m.m2[k] = fmax(syy - sy * sy / n, 0.0);
        if (nan || sy != sy || syy != syy) { m.mean[k] = NAN; m.m2[k] = NAN; }
        m.s1[k] = nan ? NAN : n * (1.0 - sh[k]) - sy;     // sum(1 - v) = n (1 - shift) - sum(v - shift)
        m.c95[k] = wacc[WA_C95 + k];
        m.c98[k] = wacc[WA_C98 + k];
    }
}

// ---------------------------------------------------------------------------------------------
// Per-lane accumulation (round 2): warp_acc_pass reduces the 32 fidelities of EVERY pass across the warp (six 5-step
// double sums, a 5-step minimum, eight ballots, lane 0's read-modify-write of the warp accumulator: ~270 warp
// instructions per pass, 5-6 % of an N=7 evaluation).  Where the shared memory of the CTA has room for eleven more
// doubles per lane, every lane keeps its own partial sums there for the whole item — ~70 instructions per pass, no
// cross-lane traffic — and the warp reduces once per item (lane_acc_reduce).  The grouping of the draws into the
// partial sums depends on B only (lane = draw mod 32 within an item), so the statistics stay independent of the
// controller sharding.
// lane slots: sy[3] syy[3] | c95[3] c98[3] as six u32 | mn | n, nan as two u32.  While an item accumulates, the warp's
// slots WA_C95.. hold the three shifts.
// ---------------------------------------------------------------------------------------------
#ifndef RC_FUSED_LANE_ACC
#define RC_FUSED_LANE_ACC 1
#endif
constexpr int LACC_DOUBLES = 11;
__host__ __device__ constexpr bool fused_lane_acc(int model, int n) {
    return RC_FUSED_LANE_ACC != 0 &&
           (long long)reg_cta_threads(n, false) * (reg_row_doubles(model, n) + LACC_DOUBLES) * 8 + 16384 + 6144 <= 232448;
}

// clip to [0, 1] on the bit pattern (6 integer instructions instead of the 14 of fmin(fmax())): negative -> +0, high
// word >= that of 1.0 -> exactly 1.0 (NaN: flagged by the caller, the item's statistics become NaN)
__device__ __forceinline__ double clip01_bits(double t) {
    int hi = __double2hiint(t), lo = __double2loint(t);
    const int keep = ~(hi >> 31);
    hi &= keep; lo &= keep;
    const bool ge1 = hi >= 0x3FF00000;
    return __hiloint2double(ge1 ? 0x3FF00000 : hi, ge1 ? 0 : lo);
}

// All 32 lanes call, converged; `valid` lanes hold a fidelity.  `first`: first pass of the item (lane 0 is valid).
__device__ __forceinline__ void lane_acc_pass(double* __restrict__ lacc, double* __restrict__ wacc, bool first, bool valid,
                                              double f, double eps) {
    if (first) {
        const int lane = threadIdx.x & 31;
        __syncwarp();
        const double f0 = __shfl_sync(0xffffffffu, f, 0);
        if (lane == 0) {
            wacc[WA_SHIFT] = f0;
            wacc[WA_C95] = f0; wacc[WA_C95 + 1] = clip01_bits(f0 - eps); wacc[WA_C95 + 2] = clip01_bits(f0 + eps);
        }
#pragma unroll
        for (int j = 0; j < LACC_DOUBLES; ++j) lacc[j] = 0.0;
        lacc[9] = INFINITY;
        __syncwarp();
    }
    if (valid) {
        const double xm = f - eps, xp = f + eps;
        const double v[3] = {f, clip01_bits(xm), clip01_bits(xp)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double y = v[k] - wacc[WA_C95 + k];
            lacc[k] += y;
            lacc[3 + k] = fma(y, y, lacc[3 + k]);
        }
        uint2* cnt = reinterpret_cast<uint2*>(lacc + 6);
        uint2 c0 = cnt[0], c1 = cnt[1], c2 = cnt[2];
        // clip to [0, 1] does not change a comparison with a threshold inside (0, 1); a NaN compares false
        if (f >= 0.95) c0.x += 1u;
        if (xm >= 0.95) c0.y += 1u;
        if (xp >= 0.95) c1.x += 1u;
        if (f >= 0.98) c1.y += 1u;
        if (xm >= 0.98) c2.x += 1u;
        if (xp >= 0.98) c2.y += 1u;
        cnt[0] = c0; cnt[1] = c1; cnt[2] = c2;
        lacc[9] = fmin(lacc[9], f);
        uint2 nn = *reinterpret_cast<uint2*>(lacc + 10);
        nn.x += 1u;
        if (f != f) nn.y += 1u;
        *reinterpret_cast<uint2*>(lacc + 10) = nn;
    }
}

// End of an item, all 32 lanes converged: the lanes' partial sums -> the warp accumulator (fixed shuffle trees).
__device__ __forceinline__ void lane_acc_reduce(const double* __restrict__ lacc, double* __restrict__ wacc) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    double tot[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) tot[k] = warp_sum_f64(lacc[k]);
    const uint2* cnt = reinterpret_cast<const uint2*>(lacc + 6);
    const uint2 c0 = cnt[0], c1 = cnt[1], c2 = cnt[2], nn = cnt[4];
    const unsigned c[6] = {__reduce_add_sync(0xffffffffu, c0.x), __reduce_add_sync(0xffffffffu, c0.y),
                           __reduce_add_sync(0xffffffffu, c1.x), __reduce_add_sync(0xffffffffu, c1.y),
                           __reduce_add_sync(0xffffffffu, c2.x), __reduce_add_sync(0xffffffffu, c2.y)};
    const unsigned n = __reduce_add_sync(0xffffffffu, nn.x), nnan = __reduce_add_sync(0xffffffffu, nn.y);
    double mn = lacc[9];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            wacc[WA_SY + k] = tot[k];
            wacc[WA_SYY + k] = tot[3 + k];
            wacc[WA_C95 + k] = (double)c[k];
            wacc[WA_C98 + k] = (double)c[3 + k];
        }
        wacc[WA_MN] = mn;
        wacc[WA_N] = (double)n;
        wacc[WA_NAN] = (double)nnan;
    }
    __syncwarp();
}

template <int N, int MODEL>
__global__ void __launch_bounds__(reg_cta_threads(N, false), reg_cta_min_blocks(N, false)) fidelity_stats_reg_warp_kernel(FusedArgs g) {
    constexpr int K = draws_per_site(MODEL) * N;
    constexpr int KP = reg_row_doubles(MODEL, N);
    extern __shared__ __align__(16) double smem_raw[];
    const FidArgs& a = g.f;
    const ZigEntry* kw = reinterpret_cast<const ZigEntry*>(smem_raw);
    double* row = smem_raw + 2 * ZIG_LAYERS + threadIdx.x * KP;
    load_zig_table(a.zig.kw, reinterpret_cast<ZigEntry*>(smem_raw));
    __shared__ double wacc_all[MAX_CTA_WARPS * WACC_DOUBLES];
    double* wacc = wacc_all + (threadIdx.x >> 5) * WACC_DOUBLES;
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    constexpr bool LACC = fused_lane_acc(MODEL, N);
    double* lacc = smem_raw + 2 * ZIG_LAYERS + (size_t)blockDim.x * KP + (size_t)threadIdx.x * LACC_DOUBLES;   // LACC only
    const long long nitems = (long long)a.S * a.C * g.nchunks;
    const long long stride = (long long)gridDim.x * wpc;
    for (long long item = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); item < nitems; item += stride) {
        const long long seg = item / g.nchunks, ch = item - seg * g.nchunks;
        const long long s = seg / a.C, c = seg - s * a.C;
        const long long b0 = ch * g.chunk;
        const long long b1 = b0 + g.chunk < a.B ? b0 + g.chunk : a.B;
        for (long long bt = b0; bt < b1; bt += 32) {
            const long long b = bt + lane;
            const bool valid = b < b1;
            double f = 0.0;
            if (valid) f = eval_reg<N, MODEL, false>(a, s, c, b, row, kw);
            if (LACC) lane_acc_pass(lacc, wacc, bt == b0, valid, f, g.eps);
            else warp_acc_pass(wacc, bt == b0, valid, f, g.eps);
        }
        if (LACC) lane_acc_reduce(lacc, wacc);
        __syncwarp();
        if (lane == 0) {
            Moments m;
            warp_acc_finish(wacc, g.eps, m);
            moments_store(m, g.partials + item * PART_DOUBLES);
        }
        __syncwarp();
    }
}

// CTA-per-item variant, kept for replay mode (the replay rows of a tile are staged behind CTA barriers).
template <int N, int MODEL, bool REPLAY>
__global__ void __launch_bounds__(reg_cta_threads(N, REPLAY), reg_cta_min_blocks(N, REPLAY)) fidelity_stats_reg_kernel(FusedArgs g) {
    constexpr int K = draws_per_site(MODEL) * N;
    constexpr int KP = reg_row_doubles(MODEL, N);
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ double scratch[MAX_CTA_WARPS * PART_DOUBLES];
    const FidArgs& a = g.f;
    const ZigEntry* kw = reinterpret_cast<const ZigEntry*>(smem_raw);
    double* stage = smem_raw + (REPLAY ? 0 : 2 * ZIG_LAYERS);
    if (!REPLAY) load_zig_table(a.zig.kw, reinterpret_cast<ZigEntry*>(smem_raw));
    const long long nseg = (long long)a.S * a.C;
    const long long nitems = nseg * g.nchunks;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const long long seg = item / g.nchunks, ch = item - seg * g.nchunks;
        const long long s = seg / a.C, c = seg - s * a.C;
        const long long b0 = ch * g.chunk;
        const long long b1 = b0 + g.chunk < a.B ? b0 + g.chunk : a.B;
        Moments m;
        moments_init(m);
        double shift[3] = {0.0, 0.0, 0.0};
        for (long long bt = b0; bt < b1; bt += blockDim.x) {
            const long long b = bt + threadIdx.x;
            if (REPLAY) {
                long long nvalid = b1 - bt < (long long)blockDim.x ? b1 - bt : (long long)blockDim.x;
                stage_replay_rows<K, KP>(a.replay + (seg * a.B + bt) * K, (int)nvalid, stage);
            }
            if (b < b1) {
                double f = eval_reg<N, MODEL, REPLAY>(a, s, c, b, stage + threadIdx.x * KP, kw);
                moments_add(m, f, g.eps, shift, m.n == 0.0);
            }
        }
        moments_finish_thread(m, shift);
        moments_block_merge(m, scratch);
        if (threadIdx.x == 0) moments_store(m, g.partials + item * PART_DOUBLES);
    }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory kernel for REG_MAX_N < N <= MAX_N: each lane owns a column of four [N] arrays.
// ---------------------------------------------------------------------------------------------
// One evaluation with the four [N] arrays in shared memory (column = this lane, stride blockDim).
// Raw standard normals are first parked in the arrays themselves (d <- z_ii, e <- nn, zi <- nn2) by a
// non-unrolled Philox pair loop (or read from the replay row), then transformed in place.
template <int MODEL, bool REPLAY, bool AMPS = false>
__device__ __forceinline__ double eval_smem(const FidArgs& a, long long s, long long c, long long b,
                                            const double* row /* global replay row */, double* sm) {
    const int n = a.N, ld = blockDim.x;
    double* d = sm + threadIdx.x;
    double* e = d + (size_t)n * ld;
    double* zi = e + (size_t)n * ld;
    double* zo = zi + (size_t)n * ld;
    constexpr int P = draws_per_site(MODEL);
    const double* x = a.ctrl + c * (n + 1);
    const double sigma = __ldg(a.sigma + s);
    if (REPLAY) {
        for (int i = 0; i < n; ++i) {
            d[(size_t)i * ld] = __ldg(row + P * i);
            if (i >= 1) {
                e[(size_t)(i - 1) * ld] = __ldg(row + P * i + 1);
                if (MODEL == MODEL_COMPLEX3) zi[(size_t)i * ld] = __ldg(row + P * i + 2);
            }
        }
    } else {
        // compact index: 0 -> z_00, then (z_ii, nn_i[, nn2_i]) for i >= 1; fast-path table read through L1
        normals_fill(noise_key(a, s, c, b), P * n - (P - 1), a.zig, [&](int jc) -> double& {
            const int site = jc == 0 ? 0 : 1 + (jc - 1) / P, kind = jc == 0 ? 0 : (jc - 1) % P;
            return *(kind == 0 ? d + (size_t)site * ld : (kind == 1 ? e + (size_t)(site - 1) * ld : zi + (size_t)site * ld));
        });
    }
    for (int i = 0; i < n; ++i) {
        const double base = a.zz ? zz_diag(i, n) : 0.0;
        d[(size_t)i * ld] = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, d[(size_t)i * ld])), __ldg(x + i));
        if (i >= 1) {
            const double aa = __dadd_rn(1.0, __dmul_rn(sigma, e[(size_t)(i - 1) * ld]));
            if (MODEL == MODEL_COMPLEX3) {
                const double bb = __dmul_rn(sigma, zi[(size_t)i * ld]);
                e[(size_t)(i - 1) * ld] = rc_sqrt(fma(aa, aa, bb * bb));
            } else {
                e[(size_t)(i - 1) * ld] = aa;
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        zi[(size_t)i * ld] = (i == a.in) ? 1.0 : 0.0;
        zo[(size_t)i * ld] = (i == a.out) ? 1.0 : 0.0;
    }
    int fail = 0;
    double re, im;
    amplitude_strided(d, e, zi, zo, ld, n, fabs(__ldg(x + n)), &fail, re, im);
    if (fail && a.nonconv) atomicAdd(a.nonconv, 1ull);
    if (AMPS) {   // two 8-byte stores: the caller's buffer is only guaranteed to be 8-byte aligned
        double* dst = a.amps + 2 * ((s * a.C + c) * a.B + b);
        dst[0] = re;
        dst[1] = im;
    }
    return fma(re, re, im * im);
}

// ---------------------------------------------------------------------------------------------
// Spectral-weights evaluation (rc_spectral.cuh) for the shared-memory family: each lane owns TWO [N] columns
// (d, e) plus, when in / out are not the chain ends, the original entries of the two outer blocks
// (2 * nx doubles, nx = min(in,out) + N-1-max(in,out)).  Half the footprint of eval_smem => twice the lanes.
// ---------------------------------------------------------------------------------------------
constexpr int ALGO_VECTORS = 0;    // QL accumulating the in / out eigenvector rows (rc_ql.cuh)
constexpr int ALGO_SPECTRAL = 1;   // eigenvalues only + characteristic-polynomial weights (rc_spectral.cuh)

__host__ __device__ inline int spec_outer_sites(int n, int in, int out) {
    const int lo = in < out ? in : out, hi = in < out ? out : in;
    return lo + (n - 1 - hi);
}
// + 1: the pad row in front of d that the unconditional operand prefetch of the chase may read (rc_spectral.cuh)
__host__ __device__ inline int spec_lane_doubles(int n, int in, int out) { return 2 * n + 2 * spec_outer_sites(n, in, out) + 1; }

// d / e of the gauge-transformed Hamiltonian of one evaluation into this lane's columns (same arithmetic, same
// rounding order as build_tridiagonal / eval_smem).  Philox mode, complex model: three draws per site but two
// columns — the imaginary coupling draw nn2_i is parked in the spare slot e[n-1] and folded into
// e[i-1] = |1 + sigma nn_i + i sigma nn2_i| as soon as its Philox block is complete (a block holds at most one).
// ZIGS: the CTA keeps a copy of the 16 KB ziggurat fast-path table behind the lanes' columns in dynamic shared
// memory (spec_zig_table; staged by the kernel) — at N = 16 a third of the noise generation's stall samples were
// waits on the L1 / L2 path of the table look-ups (profiles/r02_ncu_fidelity_smem_spectral_final_summary.txt).
template <int LD>
__device__ __forceinline__ ZigEntry* spec_zig_table(const FidArgs& a) {
    extern __shared__ double rc_spec_dyn_smem[];
    return reinterpret_cast<ZigEntry*>(rc_spec_dyn_smem + (size_t)spec_lane_doubles(a.N, a.in, a.out) * LD);
}

template <int MODEL, bool REPLAY, int LD = 0, bool ZIGS = false>
__device__ __noinline__ void build_spec(const FidArgs& a, long long s, long long c, long long b, const double* row,
                                        double* d, double* e, int ld_rt) {
    static_assert(!ZIGS || (LD > 0 && !REPLAY), "the shared table copy exists in the LD-templated Philox kernels only");
    const int n = a.N, ld = LD ? LD : ld_rt;
    constexpr int P = draws_per_site(MODEL);
    ZigTables zt = a.zig;
    if (ZIGS) zt.kw = spec_zig_table<LD ? LD : 1>(a);
    const double* x = a.ctrl + c * (n + 1);
    const double sigma = __ldg(a.sigma + s);
    if (REPLAY) {
        for (int i = 0; i < n; ++i) {
            d[(size_t)i * ld] = __ldg(row + P * i);
            if (i >= 1) {
                const double aa = __dadd_rn(1.0, __dmul_rn(sigma, __ldg(row + P * i + 1)));
                if (MODEL == MODEL_COMPLEX3) {
                    const double bb = __dmul_rn(sigma, __ldg(row + P * i + 2));
                    e[(size_t)(i - 1) * ld] = rc_sqrt(fma(aa, aa, bb * bb));
                } else {
                    e[(size_t)(i - 1) * ld] = aa;
                }
            }
        }
    } else if (MODEL == MODEL_REAL2) {
        normals_fill(noise_key(a, s, c, b), P * n - (P - 1), zt, [&](int jc) -> double& {
            const int site = jc == 0 ? 0 : 1 + (jc - 1) / P, kind = jc == 0 ? 0 : (jc - 1) % P;
            return *(kind == 0 ? d + (size_t)site * ld : e + (size_t)(site - 1) * ld);
        });
        for (int i = 1; i < n; ++i) e[(size_t)(i - 1) * ld] = __dadd_rn(1.0, __dmul_rn(sigma, e[(size_t)(i - 1) * ld]));
    } else {
        const NoiseKey key = noise_key(a, s, c, b);
        const int nc = 3 * n - 2, np = (nc + 1) / 2;
        double* spare = e + (size_t)(n - 1) * ld;
        auto slot = [&](int jc) -> double* {
            const int site = jc == 0 ? 0 : 1 + (jc - 1) / 3, kind = jc == 0 ? 0 : (jc - 1) % 3;
            return kind == 0 ? d + (size_t)site * ld : (kind == 1 ? e + (size_t)(site - 1) * ld : spare);
        };
        uint32_t pend = 0;   // parked draw indices (jc + 1), 8 bits each — see normals_fill
#pragma unroll 1
        for (int p = 0; p < np; ++p) {
            const Philox4 r = philox_block(key, (uint32_t)p);
            const int j0 = 2 * p;
            bool miss;
            *slot(j0) = zig_try(r.x, r.y, zt.kw, &miss);
            if (miss) pend = (pend << 8) | (uint32_t)(j0 + 1);
            if (j0 + 1 < nc) {
                *slot(j0 + 1) = zig_try(r.z, r.w, zt.kw, &miss);
                if (miss) pend = (pend << 8) | (uint32_t)(j0 + 2);
            }
            // the block's imaginary coupling draw, if any (jc = 3 i, i >= 1): uniform over the warp
            const int jk = (j0 % 3 == 0 && j0 > 0) ? j0 : ((j0 + 1) % 3 == 0 ? j0 + 1 : -1);
            const bool fold = jk > 0 && jk < nc;
            if (fold || (pend >> 16) || p == np - 1) {
                while (pend) {
                    const uint32_t jc = (pend & 0xFFu) - 1u;
                    pend >>= 8;
                    double* q = slot((int)jc);
                    *q = zig_complete(key, jc, *q, zt);
                }
            }
            if (fold) {
                double* ec = e + (size_t)(jk / 3 - 1) * ld;
                const double aa = __dadd_rn(1.0, __dmul_rn(sigma, *ec));
                const double bb = __dmul_rn(sigma, *spare);
                *ec = rc_sqrt(fma(aa, aa, bb * bb));
            }
        }
    }
    for (int i = 0; i < n; ++i) {
        const double base = a.zz ? zz_diag(i, n) : 0.0;
        d[(size_t)i * ld] = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, d[(size_t)i * ld])), __ldg(x + i));
    }
}

// Cold path of eval_spec (kept out of line: the hot code of this kernel family has to fit the instruction cache).
// Called by the whole warp; lanes with `mine` recompute their evaluation with the eigenvector-accumulating QL,
// two lanes' columns per matrix (own columns for d / e, the idle neighbour's for the in / out rows), even lanes
// first, then odd lanes.
template <int MODEL, bool REPLAY>
__device__ __noinline__ double2 spec_recompute(const FidArgs& a, bool mine, long long s, long long c, long long b,
                                               const double* row, double* sm, double T) {
    const int n = a.N, ld = blockDim.x, lane = threadIdx.x & 31;
    double* d = sm + ld + threadIdx.x;            // row 0 of the lane's column is the pad row
    double* e = d + (size_t)n * ld;
    double re = NAN, im = NAN;
#pragma unroll 1
    for (int par = 0; par < 2; ++par) {
        __syncwarp();
        if (mine && (lane & 1) == par) {
            build_spec<MODEL, REPLAY>(a, s, c, b, row, d, e, ld);
            double* zi = sm + ld + (threadIdx.x ^ 1);
            double* zo = zi + (size_t)n * ld;
            for (int i = 0; i < n; ++i) {
                zi[(size_t)i * ld] = (i == a.in) ? 1.0 : 0.0;
                zo[(size_t)i * ld] = (i == a.out) ? 1.0 : 0.0;
            }
            int fail = 0;
            amplitude_strided(d, e, zi, zo, ld, n, T, &fail, re, im);
            if (fail && a.nonconv) atomicAdd(a.nonconv, 1ull);
            if (a.respec) atomicAdd(a.respec, 1ull);
        }
        __syncwarp();
    }
    return make_double2(re, im);
}

// One evaluation per lane; ALL 32 lanes of the warp must call it converged (`valid` = this lane has work): the
// rare evaluations whose spectral error estimate is rejected are recomputed inside the call (spec_recompute).
template <int MODEL, bool REPLAY, bool AMPS = false, int LD = 0, bool ZIGS = false>
__device__ __forceinline__ double eval_spec(const FidArgs& a, bool valid, long long s, long long c, long long b,
                                            const double* row /* global replay row */, double* sm) {
    const int n = a.N, ld = LD ? LD : (int)blockDim.x;   // LD: compile-time CTA size (immediate shared-memory offsets)
    const int lo = a.in < a.out ? a.in : a.out, hi = a.in < a.out ? a.out : a.in;
    const int nx = lo + (n - 1 - hi);
    double* d = sm + ld + threadIdx.x;                    // row 0 of the lane's column is the pad row
    double* e = d + (size_t)n * ld;
    double* xd = e + (size_t)n * ld;
    double* xe = xd + (size_t)nx * ld;
    double re = NAN, im = NAN, T = 0.0;
    bool ok = true;
    if (valid) {
        T = fabs(__ldg(a.ctrl + c * (n + 1) + n));
        build_spec<MODEL, REPLAY, LD, ZIGS>(a, s, c, b, row, d, e, ld);
        SpecBlocks xb;
        xb.xd = xd; xb.xe = xe; xb.na = lo; xb.nb = n - 1 - hi;
        for (int j = 0; j < lo; ++j) { xd[(size_t)j * ld] = d[(size_t)j * ld]; xe[(size_t)j * ld] = e[(size_t)j * ld]; }
        for (int j = 0; j < xb.nb; ++j) {
            xd[(size_t)(lo + j) * ld] = d[(size_t)(hi + 1 + j) * ld];
            xe[(size_t)(lo + j) * ld] = j + 1 < xb.nb ? e[(size_t)(hi + 1 + j) * ld] : 0.0;
        }
        double pb = 1.0;
        for (int i = lo; i < hi; ++i) pb *= e[(size_t)i * ld];
        ok = amplitude_spectral_strided<LD>(d, e, ld, n, T, pb, xb, re, im);
    }
    const unsigned redo = __ballot_sync(0xffffffffu, valid && !ok);
    if (redo) {
        const double2 r = spec_recompute<MODEL, REPLAY>(a, valid && !ok, s, c, b, row, sm, T);
        if (valid && !ok) { re = r.x; im = r.y; }
    }
    if (AMPS && valid) {
        double* dst = a.amps + 2 * ((s * a.C + c) * a.B + b);
        dst[0] = re;
        dst[1] = im;
    }
    return fma(re, re, im * im);
}

template <int MODEL, bool REPLAY, bool AMPS = false, int ALGO = ALGO_VECTORS, int MAXT = SMEM_MAX_THREADS, int LD = 0,
          bool ZIGS = false>
__global__ void __launch_bounds__(MAXT, 1) fidelity_smem_kernel(FidArgs a) {
    extern __shared__ double sm[];
    if (ZIGS) load_zig_table(a.zig.kw, spec_zig_table<LD ? LD : 1>(a));
    const long long K = (long long)draws_per_site(MODEL) * a.N;
    const long long total = (long long)a.S * a.C * a.B;
    // CTA-uniform trip count: the spectral evaluator is a warp-collective call
    for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long ev = base + threadIdx.x;
        const bool valid = ev < total;
        EvalIndex ix = decode_eval(valid ? ev : 0, a.C, a.B);
        const double* row = REPLAY ? a.replay + (valid ? ev : 0) * K : nullptr;
        if (ALGO == ALGO_SPECTRAL) {
            const double f = eval_spec<MODEL, REPLAY, AMPS, LD, ZIGS>(a, valid, ix.s, ix.c, ix.b, row, sm);
            if (valid) a.fids[ev] = f;
        } else if (valid) {
            a.fids[ev] = eval_smem<MODEL, REPLAY, AMPS>(a, ix.s, ix.c, ix.b, row, sm);
        }
    }
}

// Warp-autonomous fused variant of the shared-memory kernel (Philox mode): see fidelity_stats_reg_warp_kernel.
template <int MODEL, int ALGO = ALGO_VECTORS, int MAXT = SMEM_MAX_THREADS, int LD = 0, bool ZIGS = false>
__global__ void __launch_bounds__(MAXT, 1) fidelity_stats_smem_warp_kernel(FusedArgs g) {
    extern __shared__ double sm[];
    if (ZIGS) load_zig_table(g.f.zig.kw, spec_zig_table<LD ? LD : 1>(g.f));
    __shared__ double wacc_all[(SMEM_MAX_THREADS / 32) * WACC_DOUBLES];
    double* wacc = wacc_all + (threadIdx.x >> 5) * WACC_DOUBLES;
    const FidArgs& a = g.f;
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long nitems = (long long)a.S * a.C * g.nchunks;
    const long long stride = (long long)gridDim.x * wpc;
    for (long long item = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); item < nitems; item += stride) {
        const long long seg = item / g.nchunks, ch = item - seg * g.nchunks;
        const long long s = seg / a.C, c = seg - s * a.C;
        const long long b0 = ch * g.chunk;
        const long long b1 = b0 + g.chunk < a.B ? b0 + g.chunk : a.B;
        for (long long bt = b0; bt < b1; bt += 32) {
            const long long b = bt + lane;
            const bool valid = b < b1;
            double f = 0.0;
            if (ALGO == ALGO_SPECTRAL) f = eval_spec<MODEL, false, false, LD, ZIGS>(a, valid, s, c, b, nullptr, sm);
            else if (valid) f = eval_smem<MODEL, false>(a, s, c, b, nullptr, sm);
            warp_acc_pass(wacc, bt == b0, valid, f, g.eps);
        }
        __syncwarp();
        if (lane == 0) {
            Moments m;
            warp_acc_finish(wacc, g.eps, m);
            moments_store(m, g.partials + item * PART_DOUBLES);
        }
        __syncwarp();
    }
}

template <int MODEL, bool REPLAY, int ALGO = ALGO_VECTORS>
__global__ void __launch_bounds__(SMEM_FUSED_MAX_THREADS) fidelity_stats_smem_kernel(FusedArgs g) {
    extern __shared__ double sm[];
    __shared__ double scratch[(SMEM_FUSED_MAX_THREADS / 32) * PART_DOUBLES];
    const FidArgs& a = g.f;
    const long long K = (long long)draws_per_site(MODEL) * a.N;
    const long long nseg = (long long)a.S * a.C;
    const long long nitems = nseg * g.nchunks;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const long long seg = item / g.nchunks, ch = item - seg * g.nchunks;
        const long long s = seg / a.C, c = seg - s * a.C;
        const long long b0 = ch * g.chunk;
        const long long b1 = b0 + g.chunk < a.B ? b0 + g.chunk : a.B;
        Moments m;
        moments_init(m);
        double shift[3] = {0.0, 0.0, 0.0};
        for (long long bt = b0; bt < b1; bt += blockDim.x) {   // CTA-uniform trips (warp-collective evaluator)
            const long long b = bt + threadIdx.x;
            const bool valid = b < b1;
            const double* row = REPLAY ? a.replay + (seg * a.B + (valid ? b : b0)) * K : nullptr;
            double f = 0.0;
            if (ALGO == ALGO_SPECTRAL) f = eval_spec<MODEL, REPLAY>(a, valid, s, c, b, row, sm);
            else if (valid) f = eval_smem<MODEL, REPLAY>(a, s, c, b, row, sm);
            if (valid) moments_add(m, f, g.eps, shift, m.n == 0.0);
        }
        moments_finish_thread(m, shift);
        moments_block_merge(m, scratch);
        if (threadIdx.x == 0) moments_store(m, g.partials + item * PART_DOUBLES);
    }
}
#endif  // __CUDACC__

void note_launch();   // rc_fidelity.cu: one kernel of this library was launched (rc_launch_count)

// launchers implemented in rc_fidelity_n.cu (one translation unit per N) / rc_fidelity.cu
typedef cudaError_t (*fid_launch_fn)(const FidArgs&, int sm_count, cudaStream_t);
typedef cudaError_t (*fused_launch_fn)(const FusedArgs&, int threads, int sm_count, cudaStream_t);

// CTA size of the register-resident kernels at run time: the compiled shape unless RC_FID_THREADS asks for less.
int reg_threads_runtime(int n, bool replay);
// CTA size of the fused (streaming statistics) register kernels for segments of B draws: the largest
// of {compiled, 384, 256, 128} that keeps the last pass over a chunk reasonably full.
int fused_reg_threads(int n, bool replay, long long B);

}  // namespace rc
