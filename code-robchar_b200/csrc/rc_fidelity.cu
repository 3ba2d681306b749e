// Dispatch of the evolution/fidelity kernels over chain length + the C-ABI entry points
// rc_fidelity_mc / rc_philox_normals.
#include <stdlib.h>
#include <string.h>
#include "rc_common.cuh"
#include "rc_fidelity.cuh"

namespace rc {

static thread_local char g_err[512];
char* last_error_buffer() { return g_err; }
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

static unsigned long long g_launches = 0;
void note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches_so_far() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < 64) cached[dev] = sm;
    return sm;
}

}  // namespace rc
#define RC_ZIG_QUAL static __device__
#include "rc_zig_table.inc"
namespace rc {

// Device addresses of the ziggurat tables (one copy per device, defined in this translation unit).
cudaError_t zig_tables_device(ZigTables* t) {
    static ZigTables cached[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (dev >= 0 && dev < 64 && cached[dev].kw) { *t = cached[dev]; return cudaSuccess; }
    void *kw = nullptr, *y = nullptr;
    err = cudaGetSymbolAddress(&kw, rc_zig_kw);
    if (err != cudaSuccess) return err;
    err = cudaGetSymbolAddress(&y, rc_zig_y);
    if (err != cudaSuccess) return err;
    t->kw = (const ZigEntry*)kw; t->y = (const double*)y;
    if (dev >= 0 && dev < 64) cached[dev] = *t;
    return cudaSuccess;
}

#define RC_DECL(n) cudaError_t launch_fid_reg_##n(const FidArgs&, int, cudaStream_t);
RC_DECL(2) RC_DECL(3) RC_DECL(4) RC_DECL(5) RC_DECL(6) RC_DECL(7) RC_DECL(8) RC_DECL(9) RC_DECL(10)
RC_DECL(11) RC_DECL(12) RC_DECL(13) RC_DECL(14) RC_DECL(15) RC_DECL(16)
#undef RC_DECL
#define RC_DECL(n) cudaError_t launch_fused_reg_##n(const FusedArgs&, int, int, cudaStream_t);
RC_DECL(2) RC_DECL(3) RC_DECL(4) RC_DECL(5) RC_DECL(6) RC_DECL(7) RC_DECL(8) RC_DECL(9) RC_DECL(10)
RC_DECL(11) RC_DECL(12) RC_DECL(13) RC_DECL(14) RC_DECL(15) RC_DECL(16)
#undef RC_DECL

static fid_launch_fn reg_table[REG_MAX_N + 1] = {
    nullptr, nullptr, launch_fid_reg_2, launch_fid_reg_3, launch_fid_reg_4, launch_fid_reg_5, launch_fid_reg_6,
    launch_fid_reg_7, launch_fid_reg_8, launch_fid_reg_9, launch_fid_reg_10, launch_fid_reg_11, launch_fid_reg_12,
    launch_fid_reg_13, launch_fid_reg_14, launch_fid_reg_15, launch_fid_reg_16};

static fused_launch_fn fused_table[REG_MAX_N + 1] = {
    nullptr, nullptr, launch_fused_reg_2, launch_fused_reg_3, launch_fused_reg_4, launch_fused_reg_5,
    launch_fused_reg_6, launch_fused_reg_7, launch_fused_reg_8, launch_fused_reg_9, launch_fused_reg_10,
    launch_fused_reg_11, launch_fused_reg_12, launch_fused_reg_13, launch_fused_reg_14, launch_fused_reg_15,
    launch_fused_reg_16};

int reg_threads_runtime(int n, bool replay) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("RC_FID_THREADS");
        env = e ? atoi(e) : 0;
        if (env < 32 || (env % 32)) env = 0;
    }
    const int compiled = reg_cta_threads(n, replay);
    return env && env < compiled ? env : compiled;
}

int fused_reg_threads(int n, bool replay, long long B) {
    const int top = reg_threads_runtime(n, replay);
    const int cand[4] = {top, 384, 256, 128};
    for (int k = 0; k < 4; ++k) {
        const long long t = cand[k];
        if (t > top) continue;
        const long long passes = (B + t - 1) / t;
        if (B * 10 >= passes * t * 9) return (int)t;   // >= 90 % of the lanes busy
    }
    return top < 128 ? top : 128;
}

// CTA size of the shared-memory kernels (each lane owns four [N] double arrays = 32 N bytes): ONE CTA per
// SM holding as many lanes as fit into the 227 KB of shared memory, rounded down to a multiple of 128
// (equal warp counts on the four schedulers; measured 5-14 % over several small CTAs, r01h), or to a
// multiple of 32 when fewer than 256 lanes fit (N >= 29: warp count matters more than balance).
// RC_SMEM_THREADS (environment) overrides for tuning.
constexpr size_t SMEM_BYTES_MAX = 227 * 1024;
constexpr size_t ZIG_TABLE_BYTES = sizeof(ZigEntry) * ZIG_LAYERS;
constexpr size_t WARP_STATIC_SMEM = (SMEM_MAX_THREADS / 32) * WACC_DOUBLES * sizeof(double);   // fidelity_stats_smem_warp_kernel
constexpr size_t FUSED_STATIC_SMEM = (SMEM_FUSED_MAX_THREADS / 32) * PART_DOUBLES * sizeof(double);   // merge scratch
static int smem_threads_bytes(size_t lane_bytes, int cap = SMEM_MAX_THREADS, size_t static_bytes = 0) {
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("RC_SMEM_THREADS");
        env = e ? atoi(e) : 0;
        if (env < 32 || env > SMEM_MAX_THREADS || (env % 32)) env = 0;
    }
    int t = env;
    if (!t) {
        const int lanes = (int)((SMEM_BYTES_MAX - static_bytes) / lane_bytes);
        t = lanes >= 256 ? lanes / 128 * 128 : lanes / 32 * 32;
    }
    if ((size_t)t * lane_bytes + static_bytes > SMEM_BYTES_MAX) t = (int)((SMEM_BYTES_MAX - static_bytes) / lane_bytes) / 32 * 32;
    return t > cap ? cap : t;
}

// Algorithm of the shared-memory family: spectral weights (rc_spectral.cuh) from N = 11 on — measured on B200
// (profiles/r02a): N=9 2.05e9 vs 2.14e9 evals/s in favour of the eigenvector rows, N=10 equal, N=12 +10 %,
// N=16 +26 %, N=24 +50 %, N=32 +58 % in favour of the spectral weights (both fit 768 lanes below N = 11, so
// the footprint does not matter there and the N^2 weight products cost what the shorter rotations save).
// RC_SMEM_ALGO=vectors|spectral (environment) forces one for A/B measurements.
constexpr int SPEC_MIN_N = 11;
static int smem_algo(int n) {
    static int v = -2;
    if (v == -2) {
        const char* s = getenv("RC_SMEM_ALGO");
        v = !s ? -1 : ((s[0] == 'v' || s[0] == '0') ? ALGO_VECTORS : ALGO_SPECTRAL);
    }
    return v >= 0 ? v : (n >= SPEC_MIN_N ? ALGO_SPECTRAL : ALGO_VECTORS);
}
// RC_SPEC_ZIGS=0 (environment) keeps the ziggurat table in global memory (A/B measurements)
static bool spec_zigs_enabled() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("RC_SPEC_ZIGS"); v = (s && s[0] == '0') ? 0 : 1; }
    return v != 0;
}
// bytes of shared memory per lane for a launch
static size_t lane_bytes_for(int algo, int n, int in, int out) {
    return algo == ALGO_SPECTRAL ? (size_t)spec_lane_doubles(n, in, out) * sizeof(double) : (size_t)32 * n;
}

// device counter of spectral evaluations that fell back to the eigenvector rows (one per device, never freed)
unsigned long long* respec_counter_device() {
    static unsigned long long* cached[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!cached[dev]) {
        unsigned long long* p = nullptr;
        if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, sizeof(unsigned long long));
        cached[dev] = p;
    }
    return cached[dev];
}

template <class Kern>
static cudaError_t launch_persistent(Kern kern, int threads, size_t smem, long long need_blocks, int sm_count, cudaStream_t st,
                                     const void* arg) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int occ = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long grid = (long long)sm_count * occ;
    if (grid > need_blocks) grid = need_blocks;
    if (grid < 1) return cudaSuccess;
    void* args[1] = {const_cast<void*>(arg)};
    note_launch();
    return cudaLaunchKernel((const void*)kern, dim3((unsigned)grid), dim3(threads), args, smem, st);
}

template <int MODEL, bool REPLAY, bool AMPS = false>
static cudaError_t launch_smem(const FidArgs& a0, int sm_count, cudaStream_t st) {
    FidArgs a = a0;
    const int algo = smem_algo(a.N);
    const size_t lane = lane_bytes_for(algo, a.N, a.in, a.out);
    const int threads = smem_threads_bytes(lane);
    const size_t smem = lane * threads;
    const long long total = (long long)a.S * a.C * a.B;
    const long long nblk = (total + threads - 1) / threads;
    if (algo == ALGO_SPECTRAL) {
        a.respec = respec_counter_device();
        // the CTA sizes the launcher picks (768 lanes for N <= 18, 640 to N = 22, 512 to N = 28, 384 above) have the stride
        // compiled in (Philox mode): shared-memory offsets become immediates
        // ... and keep the ziggurat fast-path table behind the columns when the CTA has 16 KB to spare
        const bool zigs = spec_zigs_enabled() && smem + ZIG_TABLE_BYTES <= SMEM_BYTES_MAX;
        const size_t smz = smem + (zigs ? ZIG_TABLE_BYTES : 0);
#define RC_SPEC_LAUNCH(T, MAXT)                                                                                          \
        if (!REPLAY && !AMPS && threads == T)                                                                            \
            return zigs ? launch_persistent(fidelity_smem_kernel<MODEL, false, false, ALGO_SPECTRAL, MAXT, T, true>,     \
                                            threads, smz, nblk, sm_count, st, &a)                                        \
                        : launch_persistent(fidelity_smem_kernel<MODEL, false, false, ALGO_SPECTRAL, MAXT, T, false>,    \
                                            threads, smz, nblk, sm_count, st, &a);
        RC_SPEC_LAUNCH(768, SMEM_MAX_THREADS)
        RC_SPEC_LAUNCH(640, SMEM_MAX_THREADS)
        RC_SPEC_LAUNCH(512, SMEM_WIDE_THREADS)
        RC_SPEC_LAUNCH(384, SMEM_WIDE_THREADS)
#undef RC_SPEC_LAUNCH
        if (threads <= SMEM_WIDE_THREADS)
            return launch_persistent(fidelity_smem_kernel<MODEL, REPLAY, AMPS, ALGO_SPECTRAL, SMEM_WIDE_THREADS>, threads, smem, nblk, sm_count, st, &a);
        return launch_persistent(fidelity_smem_kernel<MODEL, REPLAY, AMPS, ALGO_SPECTRAL>, threads, smem, nblk, sm_count, st, &a);
    }
    return launch_persistent(fidelity_smem_kernel<MODEL, REPLAY, AMPS, ALGO_VECTORS>, threads, smem, nblk, sm_count, st, &a);
}

// RC_REG_MAX_N=<n> (environment) moves the register/shared-memory crossover for tuning.
static int reg_crossover() {
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("RC_REG_MAX_N");
        v = s ? atoi(s) : REG_DEFAULT_N;
        if (v > REG_MAX_N) v = REG_MAX_N;
        if (v < 1) v = 1;
    }
    return v;
}

cudaError_t launch_fidelity(const FidArgs& a, cudaStream_t st) {
    int sm = device_sm_count();
    if (a.amps)   // complex amplitudes (real symmetric model only): the general-N kernel, any chain length
        return a.replay ? launch_smem<MODEL_REAL2, true, true>(a, sm, st) : launch_smem<MODEL_REAL2, false, true>(a, sm, st);
    if (a.N <= reg_crossover()) {
        FidArgs ar = a;
        ar.respec = respec_counter_device();
        return reg_table[a.N](ar, sm, st);
    }
    const bool replay = a.replay != nullptr;
    if (a.model == MODEL_COMPLEX3)
        return replay ? launch_smem<MODEL_COMPLEX3, true>(a, sm, st) : launch_smem<MODEL_COMPLEX3, false>(a, sm, st);
    return replay ? launch_smem<MODEL_REAL2, true>(a, sm, st) : launch_smem<MODEL_REAL2, false>(a, sm, st);
}

template <int MODEL>
static cudaError_t launch_fused_smem_warp(const FusedArgs& g0, int sm_count, cudaStream_t st) {
    FusedArgs g = g0;
    const int algo = smem_algo(g.f.N);
    const size_t lane = lane_bytes_for(algo, g.f.N, g.f.in, g.f.out);
    const int threads = smem_threads_bytes(lane, SMEM_MAX_THREADS, WARP_STATIC_SMEM);
    const size_t smem = lane * threads;
    const long long wpc = threads / 32;
    const long long nitems = (long long)g.f.S * g.f.C * g.nchunks;
    const long long need = (nitems + wpc - 1) / wpc;
    if (algo == ALGO_SPECTRAL) {
        g.f.respec = respec_counter_device();
        const bool zigs = spec_zigs_enabled() && smem + WARP_STATIC_SMEM + ZIG_TABLE_BYTES <= SMEM_BYTES_MAX;
        const size_t smz = smem + (zigs ? ZIG_TABLE_BYTES : 0);
#define RC_SPEC_LAUNCH(T, MAXT)                                                                                          \
        if (threads == T)                                                                                                \
            return zigs ? launch_persistent(fidelity_stats_smem_warp_kernel<MODEL, ALGO_SPECTRAL, MAXT, T, true>,        \
                                            threads, smz, need, sm_count, st, &g)                                        \
                        : launch_persistent(fidelity_stats_smem_warp_kernel<MODEL, ALGO_SPECTRAL, MAXT, T, false>,       \
                                            threads, smz, need, sm_count, st, &g);
        RC_SPEC_LAUNCH(768, SMEM_MAX_THREADS)
        RC_SPEC_LAUNCH(640, SMEM_MAX_THREADS)
        RC_SPEC_LAUNCH(512, SMEM_WIDE_THREADS)
        RC_SPEC_LAUNCH(384, SMEM_WIDE_THREADS)
#undef RC_SPEC_LAUNCH
        if (threads <= SMEM_WIDE_THREADS)
            return launch_persistent(fidelity_stats_smem_warp_kernel<MODEL, ALGO_SPECTRAL, SMEM_WIDE_THREADS>, threads, smem, need, sm_count, st, &g);
        return launch_persistent(fidelity_stats_smem_warp_kernel<MODEL, ALGO_SPECTRAL>, threads, smem, need, sm_count, st, &g);
    }
    return launch_persistent(fidelity_stats_smem_warp_kernel<MODEL, ALGO_VECTORS>, threads, smem, need, sm_count, st, &g);
}

static int fused_smem_cta_threads(int n, int in, int out) {
    return smem_threads_bytes(lane_bytes_for(smem_algo(n), n, in, out), SMEM_FUSED_MAX_THREADS, FUSED_STATIC_SMEM);
}

template <int MODEL, bool REPLAY>
static cudaError_t launch_fused_smem(const FusedArgs& g0, int sm_count, cudaStream_t st) {
    FusedArgs g = g0;
    const int algo = smem_algo(g.f.N);
    const size_t lane = lane_bytes_for(algo, g.f.N, g.f.in, g.f.out);
    const int threads = fused_smem_cta_threads(g.f.N, g.f.in, g.f.out);
    const size_t smem = lane * threads;
    const long long nitems = (long long)g.f.S * g.f.C * g.nchunks;
    if (algo == ALGO_SPECTRAL) {
        g.f.respec = respec_counter_device();
        return launch_persistent(fidelity_stats_smem_kernel<MODEL, REPLAY, ALGO_SPECTRAL>, threads, smem, nitems, sm_count, st, &g);
    }
    return launch_persistent(fidelity_stats_smem_kernel<MODEL, REPLAY, ALGO_VECTORS>, threads, smem, nitems, sm_count, st, &g);
}

cudaError_t launch_fused(const FusedArgs& g, cudaStream_t st) {
    int sm = device_sm_count();
    if (g.f.N <= reg_crossover()) {
        FusedArgs gr = g;
        gr.f.respec = respec_counter_device();
        return fused_table[g.f.N](gr, fused_reg_threads(g.f.N, g.f.replay != nullptr, g.f.B), sm, st);
    }
    const bool replay = g.f.replay != nullptr;
    if (!replay)
        return g.f.model == MODEL_COMPLEX3 ? launch_fused_smem_warp<MODEL_COMPLEX3>(g, sm, st)
                                           : launch_fused_smem_warp<MODEL_REAL2>(g, sm, st);
    return g.f.model == MODEL_COMPLEX3 ? launch_fused_smem<MODEL_COMPLEX3, true>(g, sm, st)
                                       : launch_fused_smem<MODEL_REAL2, true>(g, sm, st);
}

// chunking of the draw axis for the fused path: a multiple of the CTA size close to 4096 draws per item
static void fused_chunking_threads(long long threads, long long B, long long* chunk, long long* nchunks) {
    long long k = (4096 + threads / 2) / threads;
    if (k < 1) k = 1;
    long long ch = k * threads;
    if (B < ch) ch = (B + threads - 1) / threads * threads;
    if (ch < threads) ch = threads;
    *chunk = ch;
    *nchunks = (B + ch - 1) / ch;
    if (*nchunks < 1) *nchunks = 1;
}
// Draws per warp item of the warp-autonomous (Philox mode) fused kernels: 256 (8 passes), or the whole segment
// rounded up to 32 when it is shorter.  A function of B ONLY: the grouping of draws into partials fixes the
// order of the floating-point merges, so it must not depend on how many controllers this launch (this GPU's
// shard) holds — the statistics are bit-identical for any controller sharding.  256 keeps the tail of the static
// round-robin distribution small already for modest sweeps (>= 30 items per warp from 3e7 evaluations on) at
// 0.53 B of partials per draw.
constexpr long long WARP_CHUNK = 256;
constexpr long long WARP_MAX_CHUNKS = 512;   // per segment: bounds the partials workspace (136 B each) for B >= 1.3e5
static long long warp_chunk_for(long long B) {
    static long long base = 0;   // RC_WARP_CHUNK (environment, multiple of 32) overrides for tuning
    if (!base) {
        const char* e = getenv("RC_WARP_CHUNK");
        const long long v = e ? atoll(e) : 0;
        base = (v >= 32 && v % 32 == 0) ? v : WARP_CHUNK;
    }
    if (B < base) return (B + 31) / 32 * 32;
    long long chunk = base;      // long segments: double the item until a segment has at most WARP_MAX_CHUNKS of them
    while ((B + chunk - 1) / chunk > WARP_MAX_CHUNKS) chunk *= 2;
    return chunk;
}

static void fused_chunking(int nspin, int in, int out, bool replay, long long nseg, long long B, long long* chunk, long long* nchunks) {
    if (!replay) {
        *chunk = warp_chunk_for(B);
        *nchunks = (B + *chunk - 1) / *chunk;
        return;
    }
    const long long threads = nspin <= reg_crossover() ? fused_reg_threads(nspin, replay, B)
                                                       : fused_smem_cta_threads(nspin, in, out);
    fused_chunking_threads(threads, B, chunk, nchunks);
}

// Merging the chunk partials of a segment.  The order of the floating-point merges is FIXED and independent of
// how the draw axis is sharded over GPUs: the chunks of a segment are cut into MERGE_BLOCKS contiguous blocks
// (block v = chunks [v*nchunks/8, (v+1)*nchunks/8)); inside a block one warp merges lane-strided, then a fixed
// shuffle tree; the eight block results are merged sequentially.  A rank of a draw-sharded sweep owns whole
// blocks, so merging its blocks locally, all-gathering the [8][nseg] block results and finishing with
// blocks_finalize_kernel gives the single-GPU statistics bit for bit.
constexpr int MERGE_BLOCKS = 8;

// block v of a segment whose local partials start at global chunk index chunk0; result valid in lane 0
__device__ __forceinline__ Moments merge_block_warp(const double* __restrict__ seg_partials, long long nchunks_total,
                                                    long long chunk0, long long nchunks_local, int v) {
    const int lane = threadIdx.x & 31;
    long long lo = (long long)v * nchunks_total / MERGE_BLOCKS - chunk0;
    long long hi = (long long)(v + 1) * nchunks_total / MERGE_BLOCKS - chunk0;
    if (lo < 0) lo = 0;
    if (hi > nchunks_local) hi = nchunks_local;
    Moments m;
    moments_init(m);
    for (long long ch = lo + lane; ch < hi; ch += 32) {
        Moments o;
        moments_load(o, seg_partials + ch * PART_DOUBLES);
        moments_merge(m, o);
    }
    moments_warp_merge(m);
    return m;
}

__device__ __forceinline__ void emit_statistics(const Moments& m, long long B, double eps, double* __restrict__ stats,
                                                long long nseg, long long seg) {
    const double mn[3] = {m.mn, clip01(m.mn - eps), clip01(m.mn + eps)};
    for (int k = 0; k < 3; ++k) {
        stats[(ST_W + k) * nseg + seg] = m.s1[k] / (double)B;       // mean(1 - f) == sorted W1 formula
        stats[(ST_Q95 + k) * nseg + seg] = -1.0 * (m.c95[k] / (double)B);
        stats[(ST_Q98 + k) * nseg + seg] = -1.0 * (m.c98[k] / (double)B);
        stats[(ST_STD + k) * nseg + seg] = sqrt(m.m2[k] / (double)B);
        stats[(ST_WC + k) * nseg + seg] = -mn[k];
    }
}

// Whole segment on one GPU: one WARP per segment (the warp-autonomous fused kernels write up to B/256 partials per
// segment, so a single thread per segment would serialise hundreds of dependent L2 reads).
__global__ void __launch_bounds__(256) fused_finalize_kernel(const double* __restrict__ partials, long long nseg,
                                                             long long nchunks, long long B, double eps,
                                                             double* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long seg = warp0; seg < nseg; seg += nwarps) {
        Moments total;
        moments_init(total);
        for (int v = 0; v < MERGE_BLOCKS; ++v) {
            const Moments mb = merge_block_warp(partials + seg * nchunks * PART_DOUBLES, nchunks, 0, nchunks, v);
            if (lane == 0) moments_merge(total, mb);
        }
        if (lane == 0) emit_statistics(total, B, eps, stats, nseg, seg);
    }
}

// Draw-sharded sweep, local part: the blocks [v_lo, v_hi) of every segment -> blocks_out [v_hi - v_lo][nseg][17].
__global__ void __launch_bounds__(256) block_merge_kernel(const double* __restrict__ partials, long long nseg,
                                                          long long nchunks_local, long long chunk0, long long nchunks_total,
                                                          int v_lo, int v_hi, double* __restrict__ blocks_out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long seg = warp0; seg < nseg; seg += nwarps) {
        for (int v = v_lo; v < v_hi; ++v) {
            const Moments mb = merge_block_warp(partials + seg * nchunks_local * PART_DOUBLES, nchunks_total, chunk0,
                                                nchunks_local, v);
            if (lane == 0) moments_store(mb, blocks_out + ((long long)(v - v_lo) * nseg + seg) * PART_DOUBLES);
        }
    }
}

// ... and the finish on the gathered block results [8][nseg][17]: one thread per segment, same sequence of merges as
// lane 0 of fused_finalize_kernel.
__global__ void __launch_bounds__(256) blocks_finalize_kernel(const double* __restrict__ blocks, long long nseg, long long B,
                                                              double eps, double* __restrict__ stats) {
    for (long long seg = (long long)blockIdx.x * blockDim.x + threadIdx.x; seg < nseg; seg += (long long)gridDim.x * blockDim.x) {
        Moments total;
        moments_init(total);
        for (int v = 0; v < MERGE_BLOCKS; ++v) {
            Moments mb;
            moments_load(mb, blocks + ((long long)v * nseg + seg) * PART_DOUBLES);
            moments_merge(total, mb);
        }
        emit_statistics(total, B, eps, stats, nseg, seg);
    }
}

int check_model_args(int64_t C, int nspin, int inspin, int outspin, int S, int64_t B, int model) {
    if (nspin < 2 || nspin > MAX_N) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, MAX_N);
    if (inspin < 0 || inspin >= nspin || outspin < 0 || outspin >= nspin)
        return set_error(RC_ERR_BAD_ARG, "inspin=%d / outspin=%d outside [0,%d)", inspin, outspin, nspin);
    if (C < 0 || S < 0 || B < 0) return set_error(RC_ERR_BAD_ARG, "negative size C=%lld S=%d B=%lld", (long long)C, S, (long long)B);
    if (S > 65535) return set_error(RC_ERR_BAD_ARG, "S=%d exceeds 65535 sigma levels", S);
    if (model != MODEL_COMPLEX3 && model != MODEL_REAL2) return set_error(RC_ERR_BAD_ARG, "unknown model %d", model);
    return RC_OK;
}

__global__ void philox_normals_kernel(long long C, long long B, int S, int n, int model, uint32_t k0, uint32_t k1,
                                      long long c_off, long long b_off, ZigTables zig, double* out) {
    const int P = model == MODEL_COMPLEX3 ? 3 : 2;
    const int K = P * n;
    const long long total = (long long)S * C * B;
    for (long long ev = (long long)blockIdx.x * blockDim.x + threadIdx.x; ev < total;
         ev += (long long)gridDim.x * blockDim.x) {
        EvalIndex ix = decode_eval(ev, C, B);
        double* row = out + ev * K;
        for (int j = 0; j < K; ++j) row[j] = 0.0;
        NoiseKey key;
        key.seed_lo = k0; key.seed_hi = k1; key.sidx = (uint32_t)ix.s;
        key.cidx = (uint64_t)(ix.c + c_off); key.bidx = (uint64_t)(ix.b + b_off);
        normals_fill(key, K - (P - 1), zig, [&](int jc) -> double& { return row[jc == 0 ? 0 : jc + (P - 1)]; });
    }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_version(void) { return 200; }
namespace rc { unsigned long long launches_so_far(); }
extern "C" unsigned long long rc_launch_count(void) { return rc::launches_so_far(); }

// Name of the evolution kernel the launcher picks for a sweep of this shape (bench.py labels its roofline with it).
extern "C" int rc_evolution_kernel_name(int nspin, int replay, int fused, char* buf, size_t buf_bytes) {
    if (!buf || buf_bytes < 8) return set_error(RC_ERR_NULL, "rc_evolution_kernel_name: no buffer");
    if (nspin < 2 || nspin > MAX_N) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, MAX_N);
    const bool reg = nspin <= reg_crossover();
    const char* algo = smem_algo(nspin) == ALGO_SPECTRAL ? "spectral" : "vectors";
    if (reg)
        snprintf(buf, buf_bytes, "%s<%d>", fused ? (replay ? "fidelity_stats_reg_kernel" : "fidelity_stats_reg_warp_kernel") : "fidelity_reg_kernel", nspin);
    else
        snprintf(buf, buf_bytes, "%s<%s>", fused ? (replay ? "fidelity_stats_smem_kernel" : "fidelity_stats_smem_warp_kernel") : "fidelity_smem_kernel", algo);
    return RC_OK;
}

extern "C" int rc_spectral_fallbacks(unsigned long long* count_host, int reset, void* stream) {
    unsigned long long* p = respec_counter_device();
    if (!p) return set_error(RC_ERR_CUDA, "rc_spectral_fallbacks: no counter on this device");
    cudaStream_t st = (cudaStream_t)stream;
    if (count_host) {
        RC_CUDA_TRY(cudaMemcpyAsync(count_host, p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        RC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (reset) RC_CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(unsigned long long), st));
    return RC_OK;
}
extern "C" const char* rc_last_error(void) { return last_error_buffer(); }

extern "C" int rc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    RC_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    RC_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return RC_OK;
}

namespace rc {
// Shared implementation of rc_fidelity_mc.  The host sweep (rc_api.cu) also calls it once per sigma chunk:
// sigma_dev / fids_dev / replay_dev then point at the chunk, S is the chunk's level count and s_offset its
// first level (the Philox counters use the global level index, so chunking does not change the draws).
int fidelity_mc_impl(const char* who, const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                     const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed, int64_t c_offset,
                     int64_t b_offset, const double* replay_dev, double* fids_dev, unsigned long long* nonconv_dev,
                     int s_offset, cudaStream_t st, double* amps_dev) {
    int rcode = check_model_args(C, nspin, inspin, outspin, S, B, model);
    if (rcode) return rcode;
    if ((long long)S * C * B == 0) return RC_OK;
    if (!ctrl_dev || !sigma_dev || !fids_dev) return set_error(RC_ERR_NULL, "%s: null ctrl/sigma/fids pointer", who);
    if (s_offset < 0 || s_offset + S > 65535) return set_error(RC_ERR_BAD_ARG, "%s: sigma level range exceeds 65535", who);
    FidArgs a = {};
    a.ctrl = ctrl_dev; a.sigma = sigma_dev; a.replay = replay_dev; a.fids = fids_dev; a.nonconv = nonconv_dev;
    a.C = C; a.B = B; a.S = S; a.N = nspin; a.in = inspin; a.out = outspin; a.model = model; a.zz = zz;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
    a.c_offset = c_offset; a.b_offset = b_offset; a.s_offset = s_offset;
    a.amps = amps_dev;
    if (amps_dev && model != MODEL_REAL2)
        return set_error(RC_ERR_BAD_ARG, "%s: complex amplitudes need the real symmetric model (RC_MODEL_REAL2)", who);
    RC_CUDA_TRY(zig_tables_device(&a.zig));
    RC_CUDA_TRY(launch_fidelity(a, st));
    return RC_OK;
}
}  // namespace rc

extern "C" int rc_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                              const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                              int64_t c_offset, int64_t b_offset, const double* replay_dev, double* fids_dev,
                              unsigned long long* nonconv_dev, void* stream) {
    return fidelity_mc_impl("rc_fidelity_mc", ctrl_dev, C, nspin, inspin, outspin, sigma_dev, S, B, model, zz, seed,
                            c_offset, b_offset, replay_dev, fids_dev, nonconv_dev, 0, (cudaStream_t)stream);
}

extern "C" int rc_fidelity_mc_stats(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                    const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                                    int64_t c_offset, int64_t b_offset, const double* replay_dev, double dkw_eps,
                                    double* fids_dev, double* stats_dev, unsigned long long* nonconv_dev,
                                    unsigned long long* illegal_dev, void* stream) {
    if ((long long)S * C * B > 0 && !stats_dev) return set_error(RC_ERR_NULL, "rc_fidelity_mc_stats: null stats pointer");
    int rcode = fidelity_mc_impl("rc_fidelity_mc_stats", ctrl_dev, C, nspin, inspin, outspin, sigma_dev, S, B, model, zz,
                                 seed, c_offset, b_offset, replay_dev, fids_dev, nonconv_dev, 0, (cudaStream_t)stream);
    if (rcode || (long long)S * C * B == 0) return rcode;
    return rc_stats_unsorted(fids_dev, (int64_t)S * C, B, dkw_eps, stats_dev, illegal_dev, stream);
}

extern "C" int rc_philox_normals(int64_t C, int nspin, int S, int64_t B, int model, uint64_t seed, int64_t c_offset,
                                 int64_t b_offset, double* normals_dev, void* stream) {
    int rcode = check_model_args(C, nspin, 0, 0, S, B, model);
    if (rcode) return rcode;
    long long total = (long long)S * C * B;
    if (total == 0) return RC_OK;
    if (!normals_dev) return set_error(RC_ERR_NULL, "rc_philox_normals: null output");
    long long blocks = (total + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ZigTables zig;
    RC_CUDA_TRY(zig_tables_device(&zig));
    philox_normals_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(C, B, S, nspin, model, (uint32_t)seed,
                                                                              (uint32_t)(seed >> 32), c_offset, b_offset,
                                                                              zig, normals_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" size_t rc_fidelity_stats_workspace_bytes(int64_t nseg, int64_t B) {
    if (nseg <= 0 || B <= 0) return 256;
    // upper bound over every CTA size any path may pick
    long long nmax = 1;
    for (long long t = 32; t <= 1024; t += 32) {
        long long chunk, nchunks;
        fused_chunking_threads(t, B, &chunk, &nchunks);
        if (nchunks > nmax) nmax = nchunks;
    }
    const long long wc = warp_chunk_for(B);                            // Philox mode
    const long long warp_items = (B + wc - 1) / wc;
    if (warp_items > nmax) nmax = warp_items;
    return (size_t)nseg * nmax * PART_DOUBLES * sizeof(double) + 256;
}

extern "C" int rc_fidelity_stats(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                 const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                                 int64_t c_offset, int64_t b_offset, const double* replay_dev, double dkw_eps,
                                 double* stats_dev, unsigned long long* nonconv_dev, void* workspace_dev,
                                 size_t workspace_bytes, void* stream) {
    int rcode = check_model_args(C, nspin, inspin, outspin, S, B, model);
    if (rcode) return rcode;
    const long long nseg = (long long)S * C;
    if (nseg == 0) return RC_OK;
    if (B < 1) return set_error(RC_ERR_BAD_ARG, "rc_fidelity_stats: B must be >= 1");
    if (!ctrl_dev || !sigma_dev || !stats_dev) return set_error(RC_ERR_NULL, "rc_fidelity_stats: null ctrl/sigma/stats pointer");
    FusedArgs g = {};
    FidArgs& a = g.f;
    a.ctrl = ctrl_dev; a.sigma = sigma_dev; a.replay = replay_dev; a.fids = nullptr; a.nonconv = nonconv_dev;
    a.C = C; a.B = B; a.S = S; a.N = nspin; a.in = inspin; a.out = outspin; a.model = model; a.zz = zz;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
    a.c_offset = c_offset; a.b_offset = b_offset;
    RC_CUDA_TRY(zig_tables_device(&a.zig));
    g.eps = dkw_eps;
    fused_chunking(nspin, inspin, outspin, replay_dev != nullptr, nseg, B, &g.chunk, &g.nchunks);
    const size_t need = (size_t)nseg * g.nchunks * PART_DOUBLES * sizeof(double);
    if (!workspace_dev || workspace_bytes < need)
        return set_error(RC_ERR_WORKSPACE, "rc_fidelity_stats: workspace %zu < required %zu bytes", workspace_bytes, need);
    g.partials = (double*)workspace_dev;
    cudaStream_t st = (cudaStream_t)stream;
    RC_CUDA_TRY(launch_fused(g, st));
    long long blocks = (nseg + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fused_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(g.partials, nseg, g.nchunks, B, dkw_eps, stats_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}


// ------------------------------------------------------------------------------------------------
// Draw-sharded sweep (fewer controllers than GPUs: NStochOpt.get_rims, gen_fig_8_arim_fcall_scaling.py:121-132,
// one controller x B draws; LBFGS.wass_cost, qnewton.py:447-455).
// ------------------------------------------------------------------------------------------------
extern "C" int rc_draw_shard_range(int64_t B, int world, int rank, int64_t* b_lo, int64_t* b_hi, int* v_lo, int* v_hi) {
    if (B < 1 || world < 1 || world > MERGE_BLOCKS || MERGE_BLOCKS % world || rank < 0 || rank >= world)
        return set_error(RC_ERR_BAD_ARG, "rc_draw_shard_range: B=%lld world=%d rank=%d (world must divide %d)", (long long)B,
                         world, rank, MERGE_BLOCKS);
    const long long chunk = warp_chunk_for(B), nchunks = (B + chunk - 1) / chunk;
    const int vl = rank * (MERGE_BLOCKS / world), vh = (rank + 1) * (MERGE_BLOCKS / world);
    const long long cl = (long long)vl * nchunks / MERGE_BLOCKS, ch = (long long)vh * nchunks / MERGE_BLOCKS;
    if (b_lo) *b_lo = cl * chunk;
    if (b_hi) *b_hi = ch * chunk < B ? ch * chunk : B;
    if (v_lo) *v_lo = vl;
    if (v_hi) *v_hi = vh;
    return RC_OK;
}

extern "C" size_t rc_fidelity_stats_blocks_workspace_bytes(int64_t nseg, int64_t B, int world) {
    if (nseg <= 0 || B <= 0 || world < 1) return 256;
    const long long chunk = warp_chunk_for(B), nchunks = (B + chunk - 1) / chunk;
    const long long local = (nchunks + world - 1) / world + MERGE_BLOCKS;
    return (size_t)nseg * local * PART_DOUBLES * sizeof(double) + 256;
}

extern "C" int rc_fidelity_stats_blocks(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                        const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                                        int64_t c_offset, int64_t b_offset, int world, int rank, double dkw_eps,
                                        double* blocks_dev, unsigned long long* nonconv_dev, void* workspace_dev,
                                        size_t workspace_bytes, void* stream) {
    int rcode = check_model_args(C, nspin, inspin, outspin, S, B, model);
    if (rcode) return rcode;
    const long long nseg = (long long)S * C;
    if (nseg == 0) return RC_OK;
    int64_t b_lo, b_hi;
    int v_lo, v_hi;
    rcode = rc_draw_shard_range(B, world, rank, &b_lo, &b_hi, &v_lo, &v_hi);
    if (rcode) return rcode;
    if (!ctrl_dev || !sigma_dev || !blocks_dev) return set_error(RC_ERR_NULL, "rc_fidelity_stats_blocks: null ctrl/sigma/blocks pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long chunk = warp_chunk_for(B), nchunks_total = (B + chunk - 1) / chunk;
    const long long B_local = b_hi - b_lo;
    long long blocks = (nseg + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    FusedArgs g = {};
    g.chunk = chunk;
    g.nchunks = B_local > 0 ? (B_local + chunk - 1) / chunk : 0;
    const size_t need = (size_t)nseg * (g.nchunks > 0 ? g.nchunks : 1) * PART_DOUBLES * sizeof(double);
    if (!workspace_dev || workspace_bytes < need)
        return set_error(RC_ERR_WORKSPACE, "rc_fidelity_stats_blocks: workspace %zu < required %zu bytes", workspace_bytes, need);
    g.partials = (double*)workspace_dev;
    if (B_local > 0) {
        FidArgs& a = g.f;
        a.ctrl = ctrl_dev; a.sigma = sigma_dev; a.replay = nullptr; a.fids = nullptr; a.nonconv = nonconv_dev;
        a.C = C; a.B = B_local; a.S = S; a.N = nspin; a.in = inspin; a.out = outspin; a.model = model; a.zz = zz;
        a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
        a.c_offset = c_offset; a.b_offset = b_offset + b_lo;
        RC_CUDA_TRY(zig_tables_device(&a.zig));
        g.eps = dkw_eps;
        RC_CUDA_TRY(launch_fused(g, st));
    }
    block_merge_kernel<<<(unsigned)blocks, 256, 0, st>>>(g.partials, nseg, g.nchunks, b_lo / chunk, nchunks_total, v_lo, v_hi,
                                                         blocks_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" int rc_stats_from_blocks(const double* blocks_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                                    void* stream) {
    if (nseg < 0 || B < 1) return set_error(RC_ERR_BAD_ARG, "rc_stats_from_blocks: nseg=%lld B=%lld", (long long)nseg, (long long)B);
    if (nseg == 0) return RC_OK;
    if (!blocks_dev || !stats_dev) return set_error(RC_ERR_NULL, "rc_stats_from_blocks: null pointer");
    long long blocks = (nseg + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    blocks_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(blocks_dev, nseg, B, dkw_eps, stats_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
