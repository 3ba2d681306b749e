// Ranking stage on device: ordinal ranks, clustered ranks and the Kendall tau-b matrix.
//
// Reference path replaced: MCDataSim.get_ranks (mcsim.py:513-518),
// get_ranks_clustered_little (generate_fig4_kendallrankanalysis.py:146-164) and
// scipy.stats.kendalltau as called from jkt_or_ordinaltau_pairwise (…fig4…py:94-120).
// Everything that decides a rank is integer / comparison work on order-preserving 64-bit keys,
// so results are bit-exact given the same RIM values (ties broken by index = stable argsort).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "rc_common.cuh"
#include "rc_philox.cuh"

namespace rc {

constexpr int RANK_SMEM_MAX = 4096;

__device__ __forceinline__ unsigned long long rank_key(double f) {
    if (f != f) return ~0ull;        // NaN last (np.argsort)
    if (f == 0.0) f = 0.0;           // -0.0 == +0.0
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ---- small rows: one CTA per row, bitonic sort of (key, index) in shared memory -----------------
__global__ void __launch_bounds__(512) argsort_small_kernel(const double* __restrict__ values, long long R, int n,
                                                            int P, int* __restrict__ perm) {
    extern __shared__ unsigned long long sm[];
    unsigned long long* keys = sm;
    int* idx = (int*)(sm + P);
    for (long long row = blockIdx.x; row < R; row += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            keys[i] = i < n ? rank_key(values[row * n + i]) : ~0ull;
            idx[i] = i < n ? i : 0x7FFFFFFF;
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int p = i | j;
                    bool up = (i & k) == 0;
                    unsigned long long a = keys[i], b = keys[p];
                    int ia = idx[i], ib = idx[p];
                    bool gt = (a > b) || (a == b && ia > ib);  // index tie-break => stable order
                    if (gt == up) { keys[i] = b; keys[p] = a; idx[i] = ib; idx[p] = ia; }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) perm[row * n + i] = idx[i];
    }
}

// ---- large rows: key/index build + cub segmented radix sort (LSD radix sort is stable) ----------
__global__ void build_keys_kernel(const double* __restrict__ values, long long total, int n, unsigned long long* keys,
                                  int* idx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        keys[i] = rank_key(values[i]);
        idx[i] = (int)(i % n);
    }
}

struct RowOffset {
    int n;
    __host__ __device__ int operator()(int i) const { return i * n; }
};

__global__ void scatter_ranks_kernel(const int* __restrict__ perm, long long R, long long n, long long* __restrict__ ranks,
                                     long long offset = 0) {
    const long long total = R * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long row = i / n, pos = i - row * n;
        ranks[row * n + perm[i]] = pos + offset;  // ranks[argranks] = arange(n)  (mcsim.py:516-517)
    }
}

// One thread per row walks the sorted order (the cluster anchor x0 is a sequential dependency).
__global__ void clustered_walk_kernel(const double* __restrict__ values, const int* __restrict__ perm, long long R,
                                      long long n, double alpha, double r_fixed, double* __restrict__ out) {
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < R; row += (long long)gridDim.x * blockDim.x) {
        const double* v = values + row * n;
        const int* p = perm + row * n;
        long long last = n - 1;
        while (last > 0 && v[p[last]] != v[p[last]]) --last;  // NaNs are sorted last
        const double mn = v[p[0]], mx = v[p[last]];
        const double r = alpha < 0.0 ? r_fixed : alpha * (mx - mn);  // …fig4…py:97
        double x0 = mn, rank = 0.0;
        for (long long i = 0; i < n; ++i) {
            const double x = v[p[i]];
            if (x - x0 > r) { rank += 1.0; x0 = x; }
            out[row * n + p[i]] = rank;
        }
    }
}

// Same walk for rows of up to CW_MAX_N values with one WARP per row: the sorted values are gathered into shared
// memory by all lanes first (the one-thread version chases two dependent global loads per element — 40 us for
// the paper's 209 rows of 100), lane 0 walks them there and the ranks are scattered back by all lanes.
constexpr int CW_MAX_N = 2048;
constexpr int CW_WARPS = 4;
__global__ void __launch_bounds__(CW_WARPS * 32) clustered_walk_warp_kernel(const double* __restrict__ values,
                                                                            const int* __restrict__ perm, long long R, int n,
                                                                            double alpha, double r_fixed,
                                                                            double* __restrict__ out) {
    extern __shared__ double cw_smem[];                 // per warp: n sorted values, then n ranks (reusing the slots)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sv = cw_smem + (size_t)warp * n;
    for (long long row = (long long)blockIdx.x * CW_WARPS + warp; row < R; row += (long long)gridDim.x * CW_WARPS) {
        const double* v = values + row * n;
        const int* p = perm + row * n;
        for (int i = lane; i < n; i += 32) sv[i] = v[p[i]];
        __syncwarp();
        if (lane == 0) {
            int last = n - 1;
            while (last > 0 && sv[last] != sv[last]) --last;  // NaNs are sorted last
            const double mn = sv[0], mx = sv[last];
            const double r = alpha < 0.0 ? r_fixed : alpha * (mx - mn);  // …fig4…py:97
            double x0 = mn, rank = 0.0;
            for (int i = 0; i < n; ++i) {
                const double x = sv[i];
                if (x - x0 > r) { rank += 1.0; x0 = x; }
                sv[i] = rank;
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) out[row * n + p[i]] = sv[i];
        __syncwarp();
    }
}

static cudaError_t launch_clustered_walk(const double* values, const int* perm, long long R, long long n, double alpha,
                                         double r_fixed, double* out, cudaStream_t st) {
    if (n <= CW_MAX_N) {
        const size_t smem = (size_t)CW_WARPS * n * sizeof(double);
        cudaError_t err = cudaSuccess;
        if (smem > 40 * 1024)
            err = cudaFuncSetAttribute(clustered_walk_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        long long blocks = (R + CW_WARPS - 1) / CW_WARPS;
        if (blocks > 148 * 8) blocks = 148 * 8;
        clustered_walk_warp_kernel<<<(unsigned)blocks, CW_WARPS * 32, smem, st>>>(values, perm, R, (int)n, alpha, r_fixed, out); rc::note_launch();
    } else {
        const long long blocks = (R + 31) / 32;
        clustered_walk_kernel<<<(unsigned)blocks, 32, 0, st>>>(values, perm, R, n, alpha, r_fixed, out); rc::note_launch();
    }
    return cudaGetLastError();
}

// ---- Kendall tau-b ---------------------------------------------------------------------------
// grid.y = pair (j, i); grid.x tiles the first index a; each thread owns one a and scans b > a
// through shared-memory tiles.  Integer counts: 0 dis, 1 xtie, 2 ytie, 3 joint ties.
constexpr int KT_THREADS = 128;
__global__ void __launch_bounds__(KT_THREADS) kendall_count_kernel(const double* __restrict__ x, long long Rx,
                                                                   const long long* __restrict__ y, long long Ry, long long n,
                                                                   unsigned long long* __restrict__ counts) {
    __shared__ double sx[KT_THREADS];
    __shared__ long long sy[KT_THREADS];
    const long long grp = blockIdx.z;
    const long long pair = blockIdx.y;
    const long long j = pair / Ry, i = pair - j * Ry;
    const double* xr = x + (grp * Rx + j) * n;
    const long long* yr = y + (grp * Ry + i) * n;
    const long long a = (long long)blockIdx.x * KT_THREADS + threadIdx.x;
    const double xa = a < n ? xr[a] : 0.0;
    const long long ya = a < n ? yr[a] : 0;
    unsigned long long dis = 0, xt = 0, yt = 0, nt = 0;
    for (long long b0 = (long long)blockIdx.x * KT_THREADS; b0 < n; b0 += KT_THREADS) {
        __syncthreads();
        const long long bl = b0 + threadIdx.x;
        sx[threadIdx.x] = bl < n ? xr[bl] : 0.0;
        sy[threadIdx.x] = bl < n ? yr[bl] : 0;
        __syncthreads();
        const int lim = n - b0 < KT_THREADS ? (int)(n - b0) : KT_THREADS;
        if (a < n) {
            for (int t = 0; t < lim; ++t) {
                const long long b = b0 + t;
                if (b <= a) continue;
                const double dx = xa - sx[t];
                const long long dy = ya - sy[t];
                const bool xe = dx == 0.0, ye = dy == 0;
                xt += xe; yt += ye; nt += xe && ye;
                dis += (!xe && !ye && ((dx < 0.0) != (dy < 0)));
            }
        }
    }
    // warp reduce then one atomic per warp (integer adds commute: deterministic result)
    for (int o = 16; o > 0; o >>= 1) {
        dis += __shfl_down_sync(0xffffffffu, dis, o);
        xt += __shfl_down_sync(0xffffffffu, xt, o);
        yt += __shfl_down_sync(0xffffffffu, yt, o);
        nt += __shfl_down_sync(0xffffffffu, nt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* c = counts + (grp * Rx * Ry + pair) * 4;
        if (dis) atomicAdd(c + 0, dis);
        if (xt) atomicAdd(c + 1, xt);
        if (yt) atomicAdd(c + 2, yt);
        if (nt) atomicAdd(c + 3, nt);
    }
}

__global__ void kendall_finalize_kernel(const unsigned long long* __restrict__ counts, long long npairs, long long n,
                                        double* __restrict__ tau) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long long)gridDim.x * blockDim.x) {
        const long long tot = n * (n - 1) / 2;
        const long long dis = (long long)counts[p * 4 + 0], xtie = (long long)counts[p * 4 + 1];
        const long long ytie = (long long)counts[p * 4 + 2], ntie = (long long)counts[p * 4 + 3];
        double t;
        if (n < 2 || xtie == tot || ytie == tot) {
            t = NAN;
        } else {
            const long long cmd = tot - xtie - ytie + ntie - 2 * dis;
            t = (double)cmd / sqrt((double)(tot - xtie)) / sqrt((double)(tot - ytie));  // scipy _kendalltau, variant 'b'
            t = fmin(1.0, fmax(-1.0, t));
        }
        tau[p] = t;
    }
}

// ---- top-k selection per controller group (mcsim.py:651-660: mask = rank at sigma index 0 <= k-1,
// columns kept in their original order) -------------------------------------------------------------
__global__ void mark_topk_kernel(const int* __restrict__ perm, long long G, long long Cg, long long k,
                                 unsigned char* __restrict__ mark) {
    const long long total = G * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long g = i / k, j = i - g * k;
        mark[g * Cg + perm[g * Cg + j]] = 1;
    }
}

// one CTA per group: order-preserving compaction of the marked columns (tiled block scan)
__global__ void __launch_bounds__(256) compact_topk_kernel(const unsigned char* __restrict__ mark, long long Cg,
                                                           long long k, long long* __restrict__ sel) {
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int base;
    const long long g = blockIdx.x;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (long long c0 = 0; c0 < Cg; c0 += 256) {
        const long long c = c0 + threadIdx.x;
        const int flag = (c < Cg) ? (int)mark[g * Cg + c] : 0;
        int pos, tot;
        Scan(tmp).ExclusiveSum(flag, pos, tot);
        const int b = base;
        if (flag) sel[g * k + b + pos] = c;
        __syncthreads();
        if (threadIdx.x == 0) base = b + tot;
        __syncthreads();
    }
}

// Wsel[g][s][j] = W[s][g*Cg + sel[g][j]]
__global__ void gather_topk_kernel(const double* __restrict__ W, long long S, long long G, long long Cg, long long k,
                                   const long long* __restrict__ sel, double* __restrict__ Wsel) {
    const long long total = G * S * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long j = i % k, gs = i / k, s = gs % S, g = gs / S;
        Wsel[i] = W[s * (G * Cg) + g * Cg + sel[g * k + j]];
    }
}

// ---- ARIM + bootstrap error bar (generate_arim_all_fig5.py:119-126, mcsim.py:267-275) --------------
// One CTA per row (one (group, sigma) RIM vector of the top-k controllers).  ARIM = wd_from_ideal_zero of
// the row = its mean; error bar = population std of the same statistic over `nboot` resamples with
// replacement.  Resampling indices come from Philox4x32-10 keyed by (seed; row, resample, draw) —
// device-side counterpart of upstream's np.random.randint stream (the numpy-stream variant lives in
// the Python layer, arim.arim_bootstrap(rng_mode="numpy")).
__global__ void __launch_bounds__(128) arim_bootstrap_kernel(const double* __restrict__ rims, long long R, int k, int nboot,
                                                             uint32_t seed_lo, uint32_t seed_hi, double* __restrict__ arim,
                                                             double* __restrict__ stdv) {
    extern __shared__ double sh[];
    double* row = sh;            // [k]
    double* boot = sh + k;       // [nboot]
    __shared__ double red[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (long long r = blockIdx.x; r < R; r += gridDim.x) {
        __syncthreads();
        for (int j = threadIdx.x; j < k; j += blockDim.x) row[j] = rims[r * k + j];
        __syncthreads();
        for (int b = warp; b < nboot; b += nwarp) {   // one warp per resample
            double acc = 0.0;
            for (int j0 = 0; j0 < k; j0 += 128) {      // 4 indices per Philox call, 32 lanes
                Philox4 c;
                c.x = (uint32_t)(j0 / 128 * 32 + lane); c.y = (uint32_t)b; c.z = (uint32_t)r; c.w = (uint32_t)(r >> 32);
                Philox4 q = philox4x32_10(c, seed_lo, seed_hi);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = j0 + lane * 4 + t;
                    if (j < k) acc += row[(int)(((unsigned long long)w[t] * (unsigned)k) >> 32)];   // uniform index in [0,k)
                }
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) boot[b] = acc / (double)k;
        }
        __syncthreads();
        // centre (mean of the row) and population std over the resamples: two passes by warp 0
        if (warp == 0) {
            double s0 = 0.0, s1 = 0.0;
            for (int j = lane; j < k; j += 32) s0 += row[j];
            for (int b = lane; b < nboot; b += 32) s1 += boot[b];
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
            const double mb = s1 / (double)nboot;
            double m2 = 0.0;
            for (int b = lane; b < nboot; b += 32) { const double dlt = boot[b] - mb; m2 += dlt * dlt; }
            for (int o = 16; o > 0; o >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            if (lane == 0) { arim[r] = s0 / (double)k; stdv[r] = sqrt(m2 / (double)nboot); red[0] = 0.0; }
        }
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t rank_cub_temp(long long R, long long n) {
    size_t bytes = 0;
    auto off = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), RowOffset{(int)n});
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                             (const int*)nullptr, (int*)nullptr, (int)(R * n), (int)R, off, off + 1);
    return align256(bytes);
}

// perm (int32 [R][n]) at the start of the workspace
static int argsort_rows(const double* values, long long R, long long n, void* ws, size_t ws_bytes, cudaStream_t st) {
    int* perm = (int*)ws;
    const int sm = device_sm_count();
    if (n <= RANK_SMEM_MAX) {
        int P = 2;
        while (P < n) P <<= 1;
        int threads = P / 2 < 32 ? 32 : (P / 2 > 512 ? 512 : P / 2);
        size_t smem = (size_t)P * (sizeof(unsigned long long) + sizeof(int));
        if (smem > 40 * 1024)
            RC_CUDA_TRY(cudaFuncSetAttribute(argsort_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long grid = R < (long long)sm * 8 ? R : (long long)sm * 8;
        argsort_small_kernel<<<(unsigned)grid, threads, smem, st>>>(values, R, (int)n, P, perm); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        return RC_OK;
    }
    if (R * n >= (1ll << 31)) return set_error(RC_ERR_BAD_ARG, "rank: R*n=%lld exceeds 2^31", R * n);
    char* p = (char*)ws + align256((size_t)R * n * sizeof(int));
    unsigned long long* k_in = (unsigned long long*)p; p += align256((size_t)R * n * 8);
    unsigned long long* k_out = (unsigned long long*)p; p += align256((size_t)R * n * 8);
    int* i_in = (int*)p; p += align256((size_t)R * n * 4);
    size_t tb = rank_cub_temp(R, n);
    if ((size_t)(p - (char*)ws) + tb > ws_bytes) return set_error(RC_ERR_WORKSPACE, "rank: workspace too small");
    build_keys_kernel<<<sm * 8, 256, 0, st>>>(values, R * n, (int)n, k_in, i_in); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    auto off = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), RowOffset{(int)n});
    RC_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortPairs(p, tb, k_in, k_out, i_in, perm, (int)(R * n), (int)R, off, off + 1,
                                                         0, 64, st));
    return RC_OK;
}

}  // namespace rc

using namespace rc;

extern "C" size_t rc_ranks_workspace_bytes(int64_t R, int64_t n) {
    if (R <= 0 || n <= 0) return 256;
    size_t b = align256((size_t)R * n * sizeof(int));
    if (n > RANK_SMEM_MAX) {
        if (R * n >= (1ll << 31)) return 0;
        b += 2 * align256((size_t)R * n * 8) + align256((size_t)R * n * 4) + rank_cub_temp(R, n);
    }
    return b + 256;
}

extern "C" int rc_ranks(const double* values_dev, int64_t R, int64_t n, int64_t* ranks_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream) {
    if (R < 0 || n < 0) return set_error(RC_ERR_BAD_ARG, "rc_ranks: negative size");
    if (R == 0 || n == 0) return RC_OK;
    if (!values_dev || !ranks_dev || !workspace_dev) return set_error(RC_ERR_NULL, "rc_ranks: null pointer");
    if (workspace_bytes < rc_ranks_workspace_bytes(R, n))
        return set_error(RC_ERR_WORKSPACE, "rc_ranks: workspace %zu < %zu", workspace_bytes, rc_ranks_workspace_bytes(R, n));
    cudaStream_t st = (cudaStream_t)stream;
    int rcode = argsort_rows(values_dev, R, n, workspace_dev, workspace_bytes, st);
    if (rcode) return rcode;
    long long total = R * n;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scatter_ranks_kernel<<<(unsigned)blocks, 256, 0, st>>>((const int*)workspace_dev, R, n, (long long*)ranks_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" int rc_clustered_ranks(const double* values_dev, int64_t R, int64_t n, double alpha, double r_fixed,
                                  double* cranks_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (R < 0 || n < 0) return set_error(RC_ERR_BAD_ARG, "rc_clustered_ranks: negative size");
    if (R == 0 || n == 0) return RC_OK;
    if (!values_dev || !cranks_dev || !workspace_dev) return set_error(RC_ERR_NULL, "rc_clustered_ranks: null pointer");
    if (workspace_bytes < rc_ranks_workspace_bytes(R, n))
        return set_error(RC_ERR_WORKSPACE, "rc_clustered_ranks: workspace %zu < %zu", workspace_bytes, rc_ranks_workspace_bytes(R, n));
    cudaStream_t st = (cudaStream_t)stream;
    int rcode = argsort_rows(values_dev, R, n, workspace_dev, workspace_bytes, st);
    if (rcode) return rcode;
    RC_CUDA_TRY(launch_clustered_walk(values_dev, (const int*)workspace_dev, R, n, alpha, r_fixed, cranks_dev, st));
    return RC_OK;
}

extern "C" int rc_kendall_tau_b_batched(const double* x_dev, const int64_t* y_dev, int64_t G, int64_t Rx, int64_t Ry,
                                        int64_t n, double* tau_dev, long long* counts_dev, void* stream) {
    if (G < 0 || Rx < 0 || Ry < 0 || n < 0) return set_error(RC_ERR_BAD_ARG, "rc_kendall_tau_b: negative size");
    if (G == 0 || Rx == 0 || Ry == 0) return RC_OK;
    if (!x_dev || !y_dev || !tau_dev || !counts_dev) return set_error(RC_ERR_NULL, "rc_kendall_tau_b: null pointer");
    if (Rx * Ry > 65535 || G > 65535) return set_error(RC_ERR_BAD_ARG, "rc_kendall_tau_b: more than 65535 row pairs or groups");
    cudaStream_t st = (cudaStream_t)stream;
    const long long np = G * Rx * Ry;
    RC_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, (size_t)np * 4 * sizeof(long long), st));
    if (n >= 2) {
        dim3 grid((unsigned)((n + KT_THREADS - 1) / KT_THREADS), (unsigned)(Rx * Ry), (unsigned)G);
        kendall_count_kernel<<<grid, KT_THREADS, 0, st>>>(x_dev, Rx, (const long long*)y_dev, Ry, n,
                                                         (unsigned long long*)counts_dev); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
    }
    kendall_finalize_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>((const unsigned long long*)counts_dev, np, n, tau_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}

extern "C" int rc_kendall_tau_b(const double* x_dev, int64_t Rx, const int64_t* y_dev, int64_t Ry, int64_t n,
                                double* tau_dev, long long* counts_dev, void* stream) {
    return rc_kendall_tau_b_batched(x_dev, y_dev, 1, Rx, Ry, n, tau_dev, counts_dev, stream);
}

// Workspace layout of rc_rank_consistency: [argsort workspace (max of the two problems)] [mark bytes]
// [clustered ranks double G*S*k] [ordinal ranks int64 G*S*k] [kendall counts int64 G*S*S*4]
extern "C" size_t rc_rank_consistency_workspace_bytes(int64_t S, int64_t G, int64_t Cg, int64_t topk) {
    if (S <= 0 || G <= 0 || Cg <= 0) return 256;
    const long long k = topk < Cg ? topk : Cg;
    size_t a = rc_ranks_workspace_bytes(G, Cg), b = rc_ranks_workspace_bytes(G * S, k);
    if (a == 0 || b == 0) return 0;
    size_t ws = a > b ? a : b;
    size_t kl = 0;
    if (k > RANK_SMEM_MAX) {          // long rank vectors: the sort + merge-pass Kendall (rc_kendall_tau_b_large)
        kl = rc_kendall_large_workspace_bytes(G, S, S, k);
        if (kl == 0) return 0;
    }
    return align256(ws) + align256((size_t)G * Cg) + 2 * align256((size_t)G * S * k * 8) + align256((size_t)G * S * S * 4 * 8) +
           align256(kl) + 256;
}

extern "C" int rc_rank_consistency(const double* W_dev, int64_t S, int64_t G, int64_t Cg, int64_t topk, double alpha,
                                   double* tau_dev, int64_t* sel_dev, double* Wsel_dev, void* workspace_dev,
                                   size_t workspace_bytes, void* stream) {
    if (S < 0 || G < 0 || Cg < 0 || topk < 1) return set_error(RC_ERR_BAD_ARG, "rc_rank_consistency: bad sizes");
    if (S == 0 || G == 0 || Cg == 0) return RC_OK;
    if (!W_dev || !tau_dev || !sel_dev || !Wsel_dev || !workspace_dev) return set_error(RC_ERR_NULL, "rc_rank_consistency: null pointer");
    const size_t need = rc_rank_consistency_workspace_bytes(S, G, Cg, topk);
    if (need == 0) return set_error(RC_ERR_BAD_ARG, "rc_rank_consistency: problem too large");
    if (workspace_bytes < need) return set_error(RC_ERR_WORKSPACE, "rc_rank_consistency: workspace %zu < %zu", workspace_bytes, need);
    if (S * S > 65535 || G > 65535) return set_error(RC_ERR_BAD_ARG, "rc_rank_consistency: too many sigma pairs or groups");
    cudaStream_t st = (cudaStream_t)stream;
    const long long k = topk < Cg ? topk : Cg;
    size_t a = rc_ranks_workspace_bytes(G, Cg), b = rc_ranks_workspace_bytes(G * S, k);
    const size_t ws_sort = align256(a > b ? a : b);
    char* p = (char*)workspace_dev;
    void* sortws = p; p += ws_sort;
    unsigned char* mark = (unsigned char*)p; p += align256((size_t)G * Cg);
    double* cr = (double*)p; p += align256((size_t)G * S * k * 8);
    long long* rk = (long long*)p; p += align256((size_t)G * S * k * 8);
    long long* counts = (long long*)p; p += align256((size_t)G * S * S * 4 * 8);
    void* klws = p;
    const int sm = device_sm_count();
    // 1. ranks of the sigma-index-0 row inside every group (W row 0 is [G][Cg] contiguous)
    int rcode = argsort_rows(W_dev, G, Cg, sortws, ws_sort, st);
    if (rcode) return rcode;
    RC_CUDA_TRY(cudaMemsetAsync(mark, 0, (size_t)G * Cg, st));
    long long blocks = (G * k + 255) / 256;
    if (blocks > sm * 8) blocks = sm * 8;
    mark_topk_kernel<<<(unsigned)blocks, 256, 0, st>>>((const int*)sortws, G, Cg, k, mark); rc::note_launch();
    compact_topk_kernel<<<(unsigned)G, 256, 0, st>>>(mark, Cg, k, (long long*)sel_dev); rc::note_launch();
    blocks = (G * S * k + 255) / 256;
    if (blocks > sm * 8) blocks = sm * 8;
    gather_topk_kernel<<<(unsigned)blocks, 256, 0, st>>>(W_dev, S, G, Cg, k, (const long long*)sel_dev, Wsel_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    // 2. clustered ranks (radius alpha*(max-min) per row) and ordinal ranks + 1 of every selected row
    rcode = argsort_rows(Wsel_dev, G * S, k, sortws, ws_sort, st);
    if (rcode) return rcode;
    RC_CUDA_TRY(launch_clustered_walk(Wsel_dev, (const int*)sortws, G * S, k, alpha, 0.0, cr, st));
    blocks = (G * S * k + 255) / 256;
    if (blocks > sm * 8) blocks = sm * 8;
    scatter_ranks_kernel<<<(unsigned)blocks, 256, 0, st>>>((const int*)sortws, G * S, k, rk, 1); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    // 3. S x S Kendall tau-b per group (O(k^2) pair count for the paper's k = 100; sort + merge passes for long vectors)
    if (k > RANK_SMEM_MAX)
        return rc_kendall_tau_b_large(cr, (const int64_t*)rk, G, S, S, k, tau_dev, counts, klws,
                                      rc_kendall_large_workspace_bytes(G, S, S, k), stream);
    return rc_kendall_tau_b_batched(cr, (const int64_t*)rk, G, S, S, k, tau_dev, counts, stream);
}

extern "C" int rc_arim_bootstrap(const double* rims_dev, int64_t R, int64_t k, int nboot, uint64_t seed, double* arim_dev,
                                 double* std_dev, void* stream) {
    if (R < 0 || k < 1 || nboot < 1) return set_error(RC_ERR_BAD_ARG, "rc_arim_bootstrap: bad sizes");
    if (R == 0) return RC_OK;
    if (!rims_dev || !arim_dev || !std_dev) return set_error(RC_ERR_NULL, "rc_arim_bootstrap: null pointer");
    size_t smem = (size_t)(k + nboot) * sizeof(double);
    if (smem > 200 * 1024) return set_error(RC_ERR_BAD_ARG, "rc_arim_bootstrap: k + nboot too large for shared memory");
    if (smem > 40 * 1024)
        RC_CUDA_TRY(cudaFuncSetAttribute(arim_bootstrap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = R < (long long)device_sm_count() * 8 ? R : (long long)device_sm_count() * 8;
    arim_bootstrap_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(rims_dev, R, (int)k, nboot, (uint32_t)seed,
                                                                               (uint32_t)(seed >> 32), arim_dev, std_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
