// Kendall tau-b for long rank vectors (n > 4096): O(n log^2 n) instead of the O(n^2) pair count.
//
// Reference: scipy.stats.kendalltau as called by jkt_or_ordinaltau(_pairwise)
// (generate_fig4_kendallrankanalysis.py:90,117) — SciPy itself uses Knight's merge-sort algorithm; SURVEY 8 a17
// names top-k up to 1e5 for the scaled rank sizes.  Same integer quantities as kendall_count_kernel, so the
// result is bit-identical to the pair count (and to SciPy):
//   1. every x row and every y row gets order-preserving dense integer labels (segmented radix sort of the
//      values, label = position of the first equal element); the number of tied pairs of a row is
//      sum over sorted positions of (position - first equal position);
//   2. per (x row j, y row i): the composite keys (label_x << 32 | label_y) are sorted (segmented radix sort);
//      joint ties = the same sum on the composite keys; discordant pairs = inversions of the label_y sequence
//      in that order (equal x are ordered by y, so they contribute none; equal y never count) — counted by
//      log2(n) parallel merge passes, every element finding its rank in the partner run by binary search.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "rc_common.cuh"

namespace rc {

struct SegOffset {
    long long n;
    __host__ __device__ int operator()(int r) const { return (int)(r * n); }
};

__device__ __forceinline__ unsigned long long kl_key_f64(double f) {
    if (f != f) return ~0ull;
    if (f == 0.0) f = 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void kl_build_keys_kernel(const double* __restrict__ x, const long long* __restrict__ y, long long nx, long long ny,
                                     int n, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
    const long long total = nx + ny;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        keys[i] = i < nx ? kl_key_f64(x[i]) : ((unsigned long long)y[i - nx] ^ 0x8000000000000000ull);
        idx[i] = (int)(i % n);
    }
}

__device__ __forceinline__ int kl_lower_bound(const unsigned long long* a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// sorted keys of every row -> dense labels scattered back to the original positions + tied pairs of the row
__global__ void __launch_bounds__(256) kl_labels_kernel(const unsigned long long* __restrict__ ks, const int* __restrict__ perm,
                                                        int n, unsigned int* __restrict__ label,
                                                        unsigned long long* __restrict__ row_ties) {
    const long long row = blockIdx.y;
    const unsigned long long* k = ks + row * n;
    unsigned long long ties = 0;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x) {
        const int first = kl_lower_bound(k, pos + 1, k[pos]);
        label[row * n + perm[row * n + pos]] = (unsigned int)first;
        ties += (unsigned long long)(pos - first);
    }
    typedef cub::BlockReduce<unsigned long long, 256> Red;
    __shared__ typename Red::TempStorage tmp;
    ties = Red(tmp).Sum(ties);
    if (threadIdx.x == 0 && ties) atomicAdd(row_ties + row, ties);
}

// comp[g][j][i][a] = label_x[g][j][a] << 32 | label_y[g][i][a]
__global__ void kl_composite_kernel(const unsigned int* __restrict__ lx, const unsigned int* __restrict__ ly, long long G,
                                    long long Rx, long long Ry, int n, unsigned long long* __restrict__ comp) {
    const long long total = G * Rx * Ry * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long a = t % n, pair = t / n;
        const long long i = pair % Ry, gj = pair / Ry, j = gj % Rx, g = gj / Rx;
        comp[t] = ((unsigned long long)lx[(g * Rx + j) * n + a] << 32) | ly[(g * Ry + i) * n + a];
    }
}

// sorted composite keys of every pair: joint ties + the label_y sequence for the inversion count; also copies the
// row tie counts into the pair's count record [dis, xtie, ytie, ntie]
__global__ void __launch_bounds__(256) kl_joint_kernel(const unsigned long long* __restrict__ cs, int n, long long Rx, long long Ry,
                                                       const unsigned long long* __restrict__ row_ties, long long nxrows,
                                                       unsigned int* __restrict__ seq, unsigned long long* __restrict__ counts) {
    const long long pair = blockIdx.y;
    const unsigned long long* k = cs + pair * n;
    unsigned long long ties = 0;
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n; pos += gridDim.x * blockDim.x) {
        const unsigned long long v = k[pos];
        ties += (unsigned long long)(pos - kl_lower_bound(k, pos + 1, v));
        seq[pair * n + pos] = (unsigned int)v;
    }
    typedef cub::BlockReduce<unsigned long long, 256> Red;
    __shared__ typename Red::TempStorage tmp;
    ties = Red(tmp).Sum(ties);
    if (threadIdx.x == 0) {
        if (ties) atomicAdd(counts + pair * 4 + 3, ties);
        if (blockIdx.x == 0) {
            const long long i = pair % Ry, gj = pair / Ry;   // gj = g * Rx + j
            const long long g = gj / Rx;
            counts[pair * 4 + 1] = row_ties[gj];
            counts[pair * 4 + 2] = row_ties[nxrows + g * Ry + i];
        }
    }
}

// one merge pass over runs of length w: stable merge by ranking + inversion count
__global__ void __launch_bounds__(256) kl_merge_pass_kernel(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst,
                                                            int n, int w, unsigned long long* __restrict__ counts) {
    const long long pair = blockIdx.y;
    const unsigned int* s = src + pair * n;
    unsigned int* d = dst + pair * n;
    unsigned long long inv = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int base = q / (2 * w) * (2 * w);
        const int mid = base + w < n ? base + w : n, end = base + 2 * w < n ? base + 2 * w : n;
        const unsigned int v = s[q];
        int rank;
        if (q < mid) {                         // left run: elements of the right run strictly smaller go first
            int lo = mid, hi = end;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (s[m] < v) lo = m + 1; else hi = m; }
            rank = (q - base) + (lo - mid);
        } else {                               // right run: elements of the left run <= v go first
            int lo = base, hi = mid;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (s[m] <= v) lo = m + 1; else hi = m; }
            rank = (q - mid) + (lo - base);
            inv += (unsigned long long)(mid - lo);   // left elements strictly greater: inversions
        }
        d[base + rank] = v;
    }
    typedef cub::BlockReduce<unsigned long long, 256> Red;
    __shared__ typename Red::TempStorage tmp;
    inv = Red(tmp).Sum(inv);
    if (threadIdx.x == 0 && inv) atomicAdd(counts + pair * 4, inv);
}

__global__ void kl_finalize_kernel(const unsigned long long* __restrict__ counts, long long npairs, long long n,
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

static size_t kl_align(size_t v) { return (v + 255) & ~(size_t)255; }

struct KlLayout {
    size_t keys_in, keys_out, idx_in, perm, label, row_ties, comp_in, comp_out, seq_a, seq_b, cub, total;
};
static KlLayout kl_layout(long long G, long long Rx, long long Ry, long long n) {
    KlLayout L;
    const long long rows = G * (Rx + Ry), np = G * Rx * Ry;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += kl_align(bytes); return o; };
    L.keys_in = take((size_t)rows * n * 8); L.keys_out = take((size_t)rows * n * 8);
    L.idx_in = take((size_t)rows * n * 4); L.perm = take((size_t)rows * n * 4);
    L.label = take((size_t)rows * n * 4); L.row_ties = take((size_t)rows * 8);
    L.comp_in = take((size_t)np * n * 8); L.comp_out = take((size_t)np * n * 8);
    L.seq_a = take((size_t)np * n * 4); L.seq_b = take((size_t)np * n * 4);
    size_t t1 = 0, t2 = 0;
    SegOffset so{n};
    auto offs = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), so);
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                             (const int*)nullptr, (int*)nullptr, (int)(rows * n), (int)rows, offs, offs + 1);
    cub::DeviceSegmentedRadixSort::SortKeys(nullptr, t2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                            (int)(np * n), (int)np, offs, offs + 1);
    L.cub = take(t1 > t2 ? t1 : t2);
    L.total = off + 256;
    return L;
}

}  // namespace rc

using namespace rc;

extern "C" size_t rc_kendall_large_workspace_bytes(int64_t G, int64_t Rx, int64_t Ry, int64_t n) {
    if (G <= 0 || Rx <= 0 || Ry <= 0 || n <= 0) return 256;
    if (G * (Rx + Ry) * n >= (1ll << 31) || G * Rx * Ry * n >= (1ll << 31)) return 0;
    return kl_layout(G, Rx, Ry, n).total;
}

extern "C" int rc_kendall_tau_b_large(const double* x_dev, const int64_t* y_dev, int64_t G, int64_t Rx, int64_t Ry,
                                      int64_t n, double* tau_dev, long long* counts_dev, void* workspace_dev,
                                      size_t workspace_bytes, void* stream) {
    if (G < 0 || Rx < 0 || Ry < 0 || n < 0) return set_error(RC_ERR_BAD_ARG, "rc_kendall_tau_b_large: negative size");
    if (G == 0 || Rx == 0 || Ry == 0) return RC_OK;
    if (!x_dev || !y_dev || !tau_dev || !counts_dev) return set_error(RC_ERR_NULL, "rc_kendall_tau_b_large: null pointer");
    const size_t need = rc_kendall_large_workspace_bytes(G, Rx, Ry, n);
    if (need == 0) return set_error(RC_ERR_BAD_ARG, "rc_kendall_tau_b_large: more than 2^31 elements");
    if (Rx * Ry * G > 65535) return set_error(RC_ERR_BAD_ARG, "rc_kendall_tau_b_large: more than 65535 row pairs");
    if (!workspace_dev || workspace_bytes < need)
        return set_error(RC_ERR_WORKSPACE, "rc_kendall_tau_b_large: workspace %zu < required %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = G * (Rx + Ry), nxrows = G * Rx, np = G * Rx * Ry;
    RC_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, (size_t)np * 4 * sizeof(long long), st));
    if (n >= 2) {
        const KlLayout L = kl_layout(G, Rx, Ry, n);
        char* ws = (char*)workspace_dev;
        unsigned long long* keys_in = (unsigned long long*)(ws + L.keys_in);
        unsigned long long* keys_out = (unsigned long long*)(ws + L.keys_out);
        int* idx_in = (int*)(ws + L.idx_in);
        int* perm = (int*)(ws + L.perm);
        unsigned int* label = (unsigned int*)(ws + L.label);
        unsigned long long* row_ties = (unsigned long long*)(ws + L.row_ties);
        unsigned long long* comp_in = (unsigned long long*)(ws + L.comp_in);
        unsigned long long* comp_out = (unsigned long long*)(ws + L.comp_out);
        unsigned int* seq_a = (unsigned int*)(ws + L.seq_a);
        unsigned int* seq_b = (unsigned int*)(ws + L.seq_b);
        size_t cub_bytes = L.total - 256 - L.cub;
        const int sm = device_sm_count();
        SegOffset so{n};
        auto offs = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), so);
        RC_CUDA_TRY(cudaMemsetAsync(row_ties, 0, (size_t)rows * 8, st));
        kl_build_keys_kernel<<<sm * 8, 256, 0, st>>>(x_dev, (const long long*)y_dev, nxrows * n, G * Ry * n, (int)n, keys_in, idx_in); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        RC_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortPairs(ws + L.cub, cub_bytes, keys_in, keys_out, idx_in, perm, (int)(rows * n),
                                                             (int)rows, offs, offs + 1, 0, 64, st));
        const unsigned gx = (unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
        kl_labels_kernel<<<dim3(gx, (unsigned)rows), 256, 0, st>>>(keys_out, perm, (int)n, label, row_ties); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        kl_composite_kernel<<<sm * 8, 256, 0, st>>>(label, label + nxrows * n, G, Rx, Ry, (int)n, comp_in); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        RC_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(ws + L.cub, cub_bytes, comp_in, comp_out, (int)(np * n), (int)np, offs,
                                                            offs + 1, 0, 64, st));
        kl_joint_kernel<<<dim3(gx, (unsigned)np), 256, 0, st>>>(comp_out, (int)n, Rx, Ry, row_ties, nxrows, seq_a,
                                                               (unsigned long long*)counts_dev); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        unsigned int* src = seq_a;
        unsigned int* dst = seq_b;
        for (long long w = 1; w < n; w *= 2) {
            kl_merge_pass_kernel<<<dim3(gx, (unsigned)np), 256, 0, st>>>(src, dst, (int)n, (int)w, (unsigned long long*)counts_dev); rc::note_launch();
            RC_CUDA_TRY(cudaGetLastError());
            unsigned int* t = src; src = dst; dst = t;
        }
    }
    kl_finalize_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>((const unsigned long long*)counts_dev, np, n, tau_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
