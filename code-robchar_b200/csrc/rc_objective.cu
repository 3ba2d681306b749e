// Low-latency objective evaluation for optimiser loops (one controller, m explicit perturbations).
//
// Reference call sites: LBFGS.fidelity_ss / fidelity_ss_av / wass_cost (qnewton.py:383-455) inside
// scipy.optimize.fmin_l_bfgs_b (qnewton.py:497,513) and PPO's value loss (ppo.py:282), Environment.step
// (RLreinforceXXchain_actionedtime.py:260-276): thousands of small calls, each ~66-90 us on the CPU.
//
// Fast path (the m evaluations fit one CTA): the inputs are written into a pinned, device-mapped mailbox; ONE
// kernel reads them through the mapping (no H2D copy command), evaluates one matrix per lane, reduces the 15
// statistics in the same launch, writes fidelities / statistics / amplitudes back through the mapping and
// finally raises a sequence flag; the host spins on that flag (no stream synchronisation, whose wake-up
// alone costs more than the evaluation).  Larger m: H2D + sweep kernels + D2H through cached staging.
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "rc_common.cuh"
#include "rc_fidelity.cuh"

using namespace rc;

namespace rc {

struct ObjParams {
    const double* in;           // mapped mailbox: [x (N+1)] [sigma] [rows m*K]
    double* out;                // mapped mailbox: [fids m] [stats 15] [amps 2m]
    unsigned long long* flag;   // mapped: [0] sequence flag, [1] non-convergence count
    unsigned long long seq;
    int N, in_site, out_site, m, K, zz, want_stats, want_amps, has_rows;
    double eps;
};

__device__ __forceinline__ double ld_sys(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int MODEL>
__global__ void __launch_bounds__(256) objective_kernel(ObjParams q) {
    extern __shared__ double sm[];
    const int n = q.N, ld = blockDim.x, lane = threadIdx.x;
    double* inp = sm + (size_t)4 * n * ld;                               // [x (n+1)] [sigma] [rows m*K]
    double* fs = inp + (n + 2) + (q.has_rows ? (size_t)q.m * q.K : 0);   // fidelities of the m evaluations
    const int nin = n + 2 + (q.has_rows ? q.m * q.K : 0);
    for (int k = lane; k < nin; k += ld) inp[k] = ld_sys(q.in + k);
    __shared__ unsigned nonconv;
    if (lane == 0) nonconv = 0;
    __syncthreads();
    constexpr int P = draws_per_site(MODEL);
    if (lane < q.m) {
        double* d = sm + lane;
        double* e = d + (size_t)n * ld;
        double* zi = e + (size_t)n * ld;
        double* zo = zi + (size_t)n * ld;
        const double sigma = inp[n + 1];
        const double* row = inp + n + 2 + (size_t)lane * q.K;
        for (int i = 0; i < n; ++i) {
            const double base = q.zz ? zz_diag(i, n) : 0.0;
            const double z = q.has_rows ? row[P * i] : 0.0;
            d[(size_t)i * ld] = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, z)), inp[i]);
            if (i >= 1) {
                const double aa = __dadd_rn(1.0, __dmul_rn(sigma, q.has_rows ? row[P * i + 1] : 0.0));
                if (MODEL == MODEL_COMPLEX3) {
                    const double bb = __dmul_rn(sigma, q.has_rows ? row[P * i + 2] : 0.0);
                    e[(size_t)(i - 1) * ld] = rc_sqrt(fma(aa, aa, bb * bb));
                } else {
                    e[(size_t)(i - 1) * ld] = aa;
                }
            }
            zi[(size_t)i * ld] = (i == q.in_site) ? 1.0 : 0.0;
            zo[(size_t)i * ld] = (i == q.out_site) ? 1.0 : 0.0;
        }
        int fail = 0;
        double re, im;
        amplitude_strided(d, e, zi, zo, ld, n, fabs(inp[n]), &fail, re, im);
        if (fail) atomicAdd(&nonconv, 1u);
        const double f = fma(re, re, im * im);
        fs[lane] = f;
        q.out[lane] = f;
        if (q.want_amps) {
            q.out[q.m + RC_NUM_STATS + 2 * lane] = re;
            q.out[q.m + RC_NUM_STATS + 2 * lane + 1] = im;
        }
    }
    __syncthreads();
    if (q.want_stats && lane < 32) {
        // the 15 statistics of the m fidelities as ONE segment: same formulas as stats_unsorted_warp_kernel
        const double nB = (double)q.m, eps = q.eps;
        double sv[3] = {0, 0, 0}, mn = INFINITY;
        int c95[3] = {0, 0, 0}, c98[3] = {0, 0, 0};
        unsigned nan = 0;
        for (int j = lane; j < q.m; j += 32) {
            const double f = fs[j];
            const double v[3] = {f, fmin(fmax(f - eps, 0.0), 1.0), fmin(fmax(f + eps, 0.0), 1.0)};
#pragma unroll
            for (int k = 0; k < 3; ++k) { sv[k] += v[k]; c95[k] += v[k] >= 0.95; c98[k] += v[k] >= 0.98; }
            mn = fmin(mn, f);
            nan |= (f != f) ? 1u : 0u;
        }
        double mean[3], m2[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) { sv[k] = warp_sum_f64(sv[k]); mean[k] = sv[k] / nB; }
        for (int j = lane; j < q.m; j += 32) {
            const double f = fs[j];
            const double v[3] = {f, fmin(fmax(f - eps, 0.0), 1.0), fmin(fmax(f + eps, 0.0), 1.0)};
#pragma unroll
            for (int k = 0; k < 3; ++k) { const double dlt = v[k] - mean[k]; m2[k] = fma(dlt, dlt, m2[k]); }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            m2[k] = warp_sum_f64(m2[k]);
            c95[k] = __reduce_add_sync(0xffffffffu, c95[k]);
            c98[k] = __reduce_add_sync(0xffffffffu, c98[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        const bool anynan = __reduce_or_sync(0xffffffffu, nan) != 0;
        if (lane < 3) {
            const int k = lane;
            const double mk = k == 0 ? mn : fmin(fmax(mn + (k == 1 ? -eps : eps), 0.0), 1.0);
            const double svk = k == 0 ? sv[0] : (k == 1 ? sv[1] : sv[2]);
            const double a95 = (double)(k == 0 ? c95[0] : (k == 1 ? c95[1] : c95[2]));
            const double a98 = (double)(k == 0 ? c98[0] : (k == 1 ? c98[1] : c98[2]));
            const double mm = k == 0 ? m2[0] : (k == 1 ? m2[1] : m2[2]);
            double* out = q.out + q.m;
            out[ST_W + k] = anynan ? NAN : (nB - svk) / nB;
            out[ST_Q95 + k] = -1.0 * (a95 / nB);
            out[ST_Q98 + k] = -1.0 * (a98 / nB);
            out[ST_STD + k] = anynan ? NAN : sqrt(mm / nB);
            out[ST_WC + k] = anynan ? NAN : -mk;
        }
    }
    __syncthreads();
    if (lane == 0) {
        q.flag[1] = nonconv;
        __threadfence_system();
        *(volatile unsigned long long*)q.flag = q.seq;
    }
}

// Mailbox kept per (host thread, device): pinned + device-mapped, grown on demand.
struct Mailbox {
    char* host = nullptr;
    char* dev = nullptr;
    size_t bytes = 0;
    unsigned long long seq = 0;
    int smem_set[2] = {0, 0};
};
static cudaError_t mailbox(size_t need, Mailbox** out) {
    static thread_local Mailbox cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    Mailbox& c = cache[dev];
    if (c.bytes < need) {
        if (c.host) cudaFreeHost(c.host);
        c.host = c.dev = nullptr;
        c.bytes = 0;
        size_t cap = 1 << 16;
        while (cap < need) cap *= 2;
        e = cudaHostAlloc((void**)&c.host, cap, cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        e = cudaHostGetDevicePointer((void**)&c.dev, c.host, 0);
        if (e != cudaSuccess) { cudaFreeHost(c.host); c.host = nullptr; return e; }
        memset(c.host, 0, cap);
        c.bytes = cap;
    }
    *out = &c;
    return cudaSuccess;
}

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

constexpr size_t OBJ_SMEM_MAX = 200 * 1024;
constexpr int OBJ_MAX_LANES = 256;

// 1 when RC_OBJECTIVE_PATH=general forces the copy-based path (A/B measurements)
static bool objective_force_general() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("RC_OBJECTIVE_PATH"); v = (s && s[0] == 'g') ? 1 : 0; }
    return v == 1;
}

}  // namespace rc

namespace rc {
// Pinned staging + device scratch kept per (host thread, device): an optimiser calls the objective thousands
// of times with the same shapes, so nothing is allocated in steady state.
struct ObjectiveCtx {
    char* pin = nullptr;
    char* dev = nullptr;
    size_t bytes = 0;
};
static cudaError_t objective_ctx(size_t need, ObjectiveCtx** out) {
    static thread_local ObjectiveCtx cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    ObjectiveCtx& c = cache[dev];
    if (c.bytes < need) {
        if (c.pin) cudaFreeHost(c.pin);
        if (c.dev) cudaFree(c.dev);
        c.pin = c.dev = nullptr;
        c.bytes = 0;
        size_t cap = 4096;
        while (cap < need) cap *= 2;
        e = cudaHostAlloc((void**)&c.pin, cap, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void**)&c.dev, cap);
        if (e != cudaSuccess) { cudaFreeHost(c.pin); c.pin = nullptr; return e; }
        c.bytes = cap;
    }
    *out = &c;
    return cudaSuccess;
}
}  // namespace rc

static int objective_host_general(const double* x_host, int nspin, int inspin, int outspin, const double* rows_host,
                                 int64_t m, int model, int zz, double dkw_eps, double* fids_host, double* stats_host,
                                 double* amps_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int K = (model == RC_MODEL_COMPLEX3 ? 3 : 2) * nspin;
    // one buffer, same layout on both sides: [x (N+1)] [sigma] [rows m*K] | [nonconv counter] [fids m] [stats 15] [amps 2m]
    const size_t n_in = (size_t)(nspin + 2) + (rows_host ? (size_t)m * K : 0);
    const size_t in_bytes = (n_in + 1) * 8;                 // inputs + the zeroed counter
    const size_t out_bytes = (1 + (size_t)m + ((stats_host || amps_host) ? RC_NUM_STATS : 0) + (amps_host ? 2 * (size_t)m : 0)) * 8;
    ObjectiveCtx* c = nullptr;
    RC_CUDA_TRY(objective_ctx((n_in + 1 + 3 * (size_t)m + RC_NUM_STATS) * 8, &c));
    double* hp = (double*)c->pin;
    double* dp = (double*)c->dev;
    memcpy(hp, x_host, (size_t)(nspin + 1) * 8);
    hp[nspin + 1] = rows_host ? 1.0 : 0.0;                  // rows are explicit perturbations (sigma 1); none: sigma 0
    if (rows_host) memcpy(hp + nspin + 2, rows_host, (size_t)m * K * 8);
    hp[n_in] = 0.0;                                          // bit pattern of the uint64 counter 0
    RC_CUDA_TRY(cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, st));
    int rcode = fidelity_mc_impl("rc_objective_host", dp, 1, nspin, inspin, outspin, dp + nspin + 1, 1, m, model, zz, 0, 0, 0,
                                 rows_host ? dp + nspin + 2 : nullptr, dp + n_in + 1, (unsigned long long*)(dp + n_in), 0, st,
                                 amps_host ? dp + n_in + 1 + m + RC_NUM_STATS : nullptr);
    if (rcode) return rcode;
    if (stats_host) {   // the 15 statistics of the m fidelities as ONE segment (W = wass_cost's objective, 1 - W = their mean)
        rcode = stats_unsorted_impl(dp + n_in + 1, 1, m, dkw_eps, dp + n_in + 1 + m, 1, nullptr, st);
        if (rcode) return rcode;
    }
    RC_CUDA_TRY(cudaMemcpyAsync(hp + n_in, dp + n_in, out_bytes, cudaMemcpyDeviceToHost, st));
    RC_CUDA_TRY(cudaStreamSynchronize(st));
    unsigned long long nonconv;
    memcpy(&nonconv, hp + n_in, 8);
    if (fids_host) memcpy(fids_host, hp + n_in + 1, (size_t)m * 8);
    if (stats_host) memcpy(stats_host, hp + n_in + 1 + m, (size_t)RC_NUM_STATS * 8);
    if (amps_host) memcpy(amps_host, hp + n_in + 1 + m + RC_NUM_STATS, 2 * (size_t)m * 8);
    if (nonconv) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu evaluations (NaN written)", nonconv);
    return RC_OK;
}

static int objective_host_fast(const double* x_host, int nspin, int inspin, int outspin, const double* rows_host, int64_t m,
                               int model, int zz, double dkw_eps, double* fids_host, double* stats_host, double* amps_host,
                               cudaStream_t st, int K, int threads, size_t smem) {
    const size_t n_in = (size_t)(nspin + 2) + (rows_host ? (size_t)m * K : 0);
    const size_t n_out = (size_t)m + RC_NUM_STATS + 2 * (size_t)m;
    Mailbox* mb = nullptr;
    RC_CUDA_TRY(mailbox((2 + n_in + n_out) * 8, &mb));
    double* hp = (double*)mb->host;
    volatile unsigned long long* hflag = (volatile unsigned long long*)mb->host;
    memcpy(hp + 2, x_host, (size_t)(nspin + 1) * 8);
    hp[2 + nspin + 1] = rows_host ? 1.0 : 0.0;              // rows are explicit perturbations (sigma 1); none: sigma 0
    if (rows_host) memcpy(hp + 2 + nspin + 2, rows_host, (size_t)m * K * 8);
    ObjParams q;
    q.in = (const double*)mb->dev + 2;
    q.out = (double*)mb->dev + 2 + n_in;
    q.flag = (unsigned long long*)mb->dev;
    q.seq = ++mb->seq;
    q.N = nspin; q.in_site = inspin; q.out_site = outspin; q.m = (int)m; q.K = K; q.zz = zz;
    q.want_stats = stats_host != nullptr; q.want_amps = amps_host != nullptr; q.has_rows = rows_host != nullptr;
    q.eps = dkw_eps;
    const int mi = model == RC_MODEL_COMPLEX3 ? 0 : 1;
    if (smem > 40 * 1024 && mb->smem_set[mi] < (int)smem) {
        RC_CUDA_TRY(mi == 0 ? cudaFuncSetAttribute(objective_kernel<MODEL_COMPLEX3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OBJ_SMEM_MAX)
                            : cudaFuncSetAttribute(objective_kernel<MODEL_REAL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OBJ_SMEM_MAX));
        mb->smem_set[mi] = (int)OBJ_SMEM_MAX;
    }
    if (mi == 0) objective_kernel<MODEL_COMPLEX3><<<1, threads, smem, st>>>(q);
    else objective_kernel<MODEL_REAL2><<<1, threads, smem, st>>>(q);
    rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    // spin on the mapped flag; fall back to a stream synchronisation (which also surfaces launch failures) after 2 s
    const double t0 = now_s();
    unsigned spins = 0;
    while (*hflag != q.seq) {
        if ((++spins & 0xFFFu) == 0 && now_s() - t0 > 2.0) {
            RC_CUDA_TRY(cudaStreamSynchronize(st));
            if (*hflag != q.seq) return set_error(RC_ERR_CUDA, "rc_objective_host: the objective kernel did not publish its result");
        }
    }
    __sync_synchronize();
    const double* op = hp + 2 + n_in;
    const unsigned long long nonconv = hflag[1];
    if (fids_host) memcpy(fids_host, op, (size_t)m * 8);
    if (stats_host) memcpy(stats_host, op + m, (size_t)RC_NUM_STATS * 8);
    if (amps_host) memcpy(amps_host, op + m + RC_NUM_STATS, 2 * (size_t)m * 8);
    if (nonconv) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu evaluations (NaN written)", nonconv);
    return RC_OK;
}

extern "C" int rc_objective_host(const double* x_host, int nspin, int inspin, int outspin, const double* rows_host,
                                 int64_t m, int model, int zz, double dkw_eps, double* fids_host, double* stats_host,
                                 double* amps_host, void* stream) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (inspin < 0 || inspin >= nspin || outspin < 0 || outspin >= nspin)
        return set_error(RC_ERR_BAD_ARG, "inspin=%d / outspin=%d outside [0,%d)", inspin, outspin, nspin);
    if (model != RC_MODEL_COMPLEX3 && model != RC_MODEL_REAL2) return set_error(RC_ERR_BAD_ARG, "unknown model %d", model);
    if (m < 0) return set_error(RC_ERR_BAD_ARG, "rc_objective_host: m=%lld", (long long)m);
    if (m == 0) return RC_OK;
    if (!x_host || (!fids_host && !stats_host && !amps_host)) return set_error(RC_ERR_NULL, "rc_objective_host: null x / no output");
    if (!rows_host && m != 1) return set_error(RC_ERR_BAD_ARG, "rc_objective_host: the nominal evaluation (rows == NULL) takes m = 1");
    if (amps_host && model != RC_MODEL_REAL2)
        return set_error(RC_ERR_BAD_ARG, "rc_objective_host: complex amplitudes need the real symmetric model (RC_MODEL_REAL2)");
    const int K = (model == RC_MODEL_COMPLEX3 ? 3 : 2) * nspin;
    if (m <= OBJ_MAX_LANES && !objective_force_general()) {
        const int threads = (int)((m + 31) / 32 * 32);
        const size_t smem = ((size_t)4 * nspin * threads + (size_t)(nspin + 2) + (rows_host ? (size_t)m * K : 0) + (size_t)m) * 8;
        if (smem <= OBJ_SMEM_MAX)
            return objective_host_fast(x_host, nspin, inspin, outspin, rows_host, m, model, zz, dkw_eps, fids_host, stats_host,
                                       amps_host, (cudaStream_t)stream, K, threads, smem);
    }
    return objective_host_general(x_host, nspin, inspin, outspin, rows_host, m, model, zz, dkw_eps, fids_host, stats_host,
                                  amps_host, stream);
}
