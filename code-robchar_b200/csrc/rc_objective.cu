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
    int reg_ql;                 // chains of N <= 8: eigensolve in registers (fidelity_reg_compact) instead of shared memory
    double eps;
};

__device__ __forceinline__ double ld_sys(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// One evaluation with the register-resident QL of the short-chain sweep kernels (same build arithmetic as the strided
// path below; the eigensolver differs in rounding only).  scratch: 2N doubles at stride ld.
template <int MODEL, int N>
__device__ __noinline__ double objective_eval_reg(const ObjParams& q, const double* inp, const double* row, double* scratch,
                                                  int ld, int* fail, double* amp) {
    constexpr int P = draws_per_site(MODEL);
    double d[N], e[N];
    const double sigma = inp[N + 1];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double base = q.zz ? zz_diag(i, N) : 0.0;
        const double z = q.has_rows ? row[P * i] : 0.0;
        d[i] = __dadd_rn(__dadd_rn(base, __dmul_rn(sigma, z)), inp[i]);
        if (i >= 1) {
            const double aa = __dadd_rn(1.0, __dmul_rn(sigma, q.has_rows ? row[P * i + 1] : 0.0));
            if (MODEL == MODEL_COMPLEX3) {
                const double bb = __dmul_rn(sigma, q.has_rows ? row[P * i + 2] : 0.0);
                e[i - 1] = rc_sqrt(fma(aa, aa, bb * bb));
            } else {
                e[i - 1] = aa;
            }
        }
    }
    e[N - 1] = 0.0;
    return fidelity_reg_compact<N, true>(d, e, q.in_site, q.out_site, fabs(inp[N]), scratch, ld, fail, amp);
}

// One request: stage the inputs, one matrix per lane, the 15 statistics, results + flag through the mapping.
// Called by every thread of the CTA.
template <int MODEL>
__device__ __forceinline__ void objective_body(const ObjParams& q, double* sm) {
    const int n = q.N, ld = blockDim.x, lane = threadIdx.x;
    double* inp = sm + (size_t)4 * n * ld;                               // [x (n+1)] [sigma] [rows m*K]
    double* fs = inp + (n + 2) + (q.has_rows ? (size_t)q.m * q.K : 0);   // fidelities of the m evaluations
    const int nin = n + 2 + (q.has_rows ? q.m * q.K : 0);
    // reads through the mapping cost a PCIe round trip each: four in flight per lane before the first one is used
    for (int k0 = lane; k0 < nin; k0 += 4 * ld) {
        double r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = (k0 + u * ld < nin) ? ld_sys(q.in + k0 + u * ld) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (k0 + u * ld < nin) inp[k0 + u * ld] = r[u];
    }
    __shared__ unsigned nonconv;
    if (lane == 0) nonconv = 0;
    __syncthreads();
    // one fidelity and nothing else wanted: it travels WITH the sequence flag in one 16-byte store (no fence)
    const bool fast_publish = q.m == 1 && !q.want_stats && !q.want_amps;
    constexpr int P = draws_per_site(MODEL);
    if (lane < q.m && q.reg_ql && n <= 8) {
        const double* row = inp + n + 2 + (size_t)lane * q.K;
        int fail = 0;
        double amp[2], f;
        switch (n) {
            case 2: f = objective_eval_reg<MODEL, 2>(q, inp, row, sm + lane, ld, &fail, amp); break;
            case 3: f = objective_eval_reg<MODEL, 3>(q, inp, row, sm + lane, ld, &fail, amp); break;
            case 4: f = objective_eval_reg<MODEL, 4>(q, inp, row, sm + lane, ld, &fail, amp); break;
            case 5: f = objective_eval_reg<MODEL, 5>(q, inp, row, sm + lane, ld, &fail, amp); break;
            case 6: f = objective_eval_reg<MODEL, 6>(q, inp, row, sm + lane, ld, &fail, amp); break;
            case 7: f = objective_eval_reg<MODEL, 7>(q, inp, row, sm + lane, ld, &fail, amp); break;
            default: f = objective_eval_reg<MODEL, 8>(q, inp, row, sm + lane, ld, &fail, amp); break;
        }
        if (fail) atomicAdd(&nonconv, 1u);
        fs[lane] = f;
        if (!fast_publish) q.out[lane] = f;
        if (q.want_amps) {
            q.out[q.m + RC_NUM_STATS + 2 * lane] = amp[0];
            q.out[q.m + RC_NUM_STATS + 2 * lane + 1] = amp[1];
        }
    } else if (lane < q.m) {
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
        if (!fast_publish) q.out[lane] = f;
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
        // flag words: [0] sequence, [1] the fidelity of a fast publish, [3] non-convergence count (written only when
        // non-zero; the host clears it after reading)
        if (nonconv) { *(volatile unsigned long long*)(q.flag + 3) = nonconv; __threadfence_system(); }
        if (fast_publish) {
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(q.flag), "l"(q.seq),
                         "l"((unsigned long long)__double_as_longlong(fs[0]))
                         : "memory");
        } else {
            __threadfence_system();
            *(volatile unsigned long long*)q.flag = q.seq;
        }
    }
}

template <int MODEL>
__global__ void __launch_bounds__(256) objective_kernel(ObjParams q) {
    extern __shared__ double sm[];
    objective_body<MODEL>(q, sm);
}

// ---------------------------------------------------------------------------------------------------------------
// Resident server: an optimiser loop calls the objective every few tens of microseconds, and ~10 us of each
// one-shot call are the kernel launch.  The server is ONE CTA that stays on the device between calls: lane 0 polls
// a request sequence number in the pinned mailbox, the CTA then reads the request through the mapping, evaluates
// it with the same code as the one-shot kernel and publishes results + flag.  It leaves by itself after
// `idle_ns` without a request (so device-wide synchronisations wait at most that long), announcing it in the
// mailbox: state = EXITING, one last look at the request word (a request posted in that window is still served),
// state = EXITED.  The host relaunches it when it finds it gone.
// Mailbox words (8 bytes each): device -> host [0] done sequence, [1] the fidelity of a one-value answer, [2] state,
// [3] non-convergence count; host -> device [16] request sequence, [17] m | flags << 32 (flag bits: 1 statistics,
// 2 amplitudes, 4 rows present, 8 leave), [18] dkw eps — one 32-byte sector, read by one coalesced load;
// [SRV_IN..) inputs [x (N+1)] [sigma] [rows m*K], then (next 128-byte line) outputs [fids m] [stats 15] [amps 2m].
// Measured on B200 (tools/server_probe.py): every extra warp that read the three header words itself added 2.4-3.5 us
// to a call, ten dependent 256-byte input reads 11 us — reads through the mapping are issued as few, as wide and as
// early as possible.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SRV_REQ = 16, SRV_HDR = 17, SRV_EPS = 18, SRV_IN = 32;
// outputs start on their own 128-byte line behind the inputs
__host__ __device__ inline size_t srv_out_offset(int n, long long m_rows, int K) {
    return ((size_t)(n + 2) + (size_t)m_rows * K + 15) / 16 * 16;
}
constexpr unsigned long long SRV_RUNNING = 1, SRV_EXITING = 2, SRV_EXITED = 3;   // state = generation * 4 + one of these

struct SrvParams {
    unsigned long long* ctl;      // mapped mailbox
    unsigned long long start_seq; // last request already answered
    unsigned long long idle_ns, gen;
    int N, in_site, out_site, K, zz, reg_ql;
};

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int MODEL>
__global__ void __launch_bounds__(256) objective_server_kernel(SrvParams p) {
    extern __shared__ double sm[];
    __shared__ unsigned long long s_seq, s_hdr, s_eps;
    __shared__ int s_last;
    const int lane = threadIdx.x;
    unsigned long long done = p.start_seq;
    volatile unsigned long long* ctl = p.ctl;
    if (lane == 0) ctl[2] = p.gen * 4 + SRV_RUNNING;
    bool last = false;
    while (!last) {
        if (lane < 32) {
            // warp 0 polls: lanes 0..2 read {request sequence, header, eps} — one 32-byte sector, one PCIe read.
            // (Reads through the mapping are expensive and those of one line do not overlap: nobody else reads it.)
            unsigned long long w = 0, v, t0 = global_ns();
            int leaving = 0;
            unsigned polls = 0;
            for (;;) {
                if (lane < 3) w = ld_sys_u64(p.ctl + SRV_REQ + lane);
                v = __shfl_sync(0xffffffffu, w, 0);
                if (v != done) break;
                if (leaving) { v = 0; break; }
                if ((++polls & 15u) == 0 && global_ns() - t0 > p.idle_ns) {   // idle: announce, then look once more
                    if (lane == 0) { ctl[2] = p.gen * 4 + SRV_EXITING; __threadfence_system(); }
                    leaving = 1;
                }
                leaving = __shfl_sync(0xffffffffu, leaving, 0);
            }
            const unsigned long long hdr = __shfl_sync(0xffffffffu, w, 1), eps = __shfl_sync(0xffffffffu, w, 2);
            if (lane == 0) { s_seq = v; s_hdr = hdr; s_eps = eps; s_last = leaving; }
        }
        __syncthreads();
        const unsigned long long v = s_seq;
        last = s_last != 0;
        if (v != 0 && v != done) {
            const unsigned long long hdr = s_hdr;
            ObjParams q;
            q.m = (int)(hdr & 0xffffffffull);
            const unsigned flags = (unsigned)(hdr >> 32);
            q.want_stats = flags & 1; q.want_amps = (flags >> 1) & 1; q.has_rows = (flags >> 2) & 1;
            q.eps = __longlong_as_double((long long)s_eps);
            q.N = p.N; q.in_site = p.in_site; q.out_site = p.out_site; q.K = p.K; q.zz = p.zz; q.reg_ql = p.reg_ql;
            q.in = reinterpret_cast<const double*>(p.ctl + SRV_IN);
            q.out = reinterpret_cast<double*>(p.ctl + SRV_IN) + srv_out_offset(p.N, q.has_rows ? q.m : 0, p.K);
            q.flag = p.ctl;
            q.seq = v;
            if (flags & 8u) {                                 // leave
                last = true;
                if (lane == 0) { __threadfence_system(); ctl[0] = v; }
            } else if (q.m >= 1 && q.m <= (int)blockDim.x) {
                objective_body<MODEL>(q, sm);
            }
            done = v;
        }
        __syncthreads();
    }
    if (lane == 0) {
        __threadfence_system();
        ctl[2] = p.gen * 4 + SRV_EXITED;
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

// RC_OBJECTIVE_REGQL=0: chains of N <= 8 use the shared-memory eigensolver like the longer ones (A/B measurements)
static int objective_reg_ql() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("RC_OBJECTIVE_REGQL"); v = (s && s[0] == '0') ? 0 : 1; }
    return v;
}
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
    RC_CUDA_TRY(mailbox((16 + n_in + n_out) * 8, &mb));
    double* hp = (double*)mb->host;
    volatile unsigned long long* hflag = (volatile unsigned long long*)mb->host;
    memcpy(hp + 16, x_host, (size_t)(nspin + 1) * 8);
    hp[16 + nspin + 1] = rows_host ? 1.0 : 0.0;             // rows are explicit perturbations (sigma 1); none: sigma 0
    if (rows_host) memcpy(hp + 16 + nspin + 2, rows_host, (size_t)m * K * 8);
    ObjParams q;
    q.in = (const double*)mb->dev + 16;
    q.out = (double*)mb->dev + 16 + n_in;
    q.flag = (unsigned long long*)mb->dev;
    q.seq = ++mb->seq;
    q.N = nspin; q.in_site = inspin; q.out_site = outspin; q.m = (int)m; q.K = K; q.zz = zz;
    q.want_stats = stats_host != nullptr; q.want_amps = amps_host != nullptr; q.has_rows = rows_host != nullptr;
    q.eps = dkw_eps;
    q.reg_ql = objective_reg_ql();
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
    const double* op = hp + 16 + n_in;
    const unsigned long long nonconv = hflag[3];
    if (nonconv) hflag[3] = 0;
    if (m == 1 && !stats_host && !amps_host) op = hp + 1;   // the one-value answer came with the flag
    if (fids_host) memcpy(fids_host, op, (size_t)m * 8);
    if (stats_host) memcpy(stats_host, op + m, (size_t)RC_NUM_STATS * 8);
    if (amps_host) memcpy(amps_host, op + m + RC_NUM_STATS, 2 * (size_t)m * 8);
    if (nonconv) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu evaluations (NaN written)", nonconv);
    return RC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Host side of the resident server (one per host thread and device).
// ---------------------------------------------------------------------------------------------------------------
namespace rc {
struct Server {
    char* host = nullptr;
    char* dev = nullptr;
    size_t bytes = 0;
    cudaStream_t st = nullptr;
    unsigned long long seq = 0, gen = 0;
    bool launched = false, broken = false;
    int model = -1, N = 0, in = 0, out = 0, zz = 0, threads = 0;
    int smem_set[2] = {0, 0};
};
static Server* server_slot() {
    static thread_local Server cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return &cache[dev];
}
// RC_OBJECTIVE_SERVER=0 keeps the one-shot launches; RC_OBJECTIVE_IDLE_US = idle time after which the server leaves
static bool server_enabled() {
    static int v = -1;
    if (v < 0) { const char* s = getenv("RC_OBJECTIVE_SERVER"); v = (s && s[0] == '0') ? 0 : 1; }
    return v == 1;
}
static unsigned long long server_idle_ns() {
    static long long v = -1;
    if (v < 0) { const char* s = getenv("RC_OBJECTIVE_IDLE_US"); v = s ? atoll(s) : 1000; if (v < 1) v = 1; if (v > 1000000) v = 1000000; }
    return (unsigned long long)v * 1000ull;
}
static bool server_alive(const Server& sv) {
    return sv.launched && ((volatile unsigned long long*)sv.host)[2] != sv.gen * 4 + SRV_EXITED;
}
// waits until the server has answered request `seq` or is gone; 0 = answered, 1 = gone without answering, 2 = timeout
static int server_wait(const Server& sv, unsigned long long seq, double timeout_s) {
    volatile unsigned long long* c = (volatile unsigned long long*)sv.host;
    const double t0 = now_s();
    unsigned spins = 0;
    for (;;) {
        if (c[0] == seq) return 0;
        if (c[2] == sv.gen * 4 + SRV_EXITED) return c[0] == seq ? 0 : 1;
        if ((++spins & 0xFFFu) == 0 && now_s() - t0 > timeout_s) return 2;
    }
}
static cudaError_t server_launch(Server& sv, int model, int N, int in, int out, int zz, int threads, int K) {
    const size_t smem = ((size_t)4 * N * threads + (size_t)(N + 2) + (size_t)threads * K + (size_t)threads) * 8;
    const int mi = model == RC_MODEL_COMPLEX3 ? 0 : 1;
    cudaError_t e;
    if (smem > 40 * 1024 && sv.smem_set[mi] < (int)OBJ_SMEM_MAX) {
        e = mi == 0 ? cudaFuncSetAttribute(objective_server_kernel<MODEL_COMPLEX3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OBJ_SMEM_MAX)
                    : cudaFuncSetAttribute(objective_server_kernel<MODEL_REAL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OBJ_SMEM_MAX);
        if (e != cudaSuccess) return e;
        sv.smem_set[mi] = (int)OBJ_SMEM_MAX;
    }
    SrvParams p;
    p.ctl = (unsigned long long*)sv.dev;
    p.start_seq = sv.seq;
    p.idle_ns = server_idle_ns();
    p.gen = ++sv.gen;
    p.N = N; p.in_site = in; p.out_site = out; p.K = K; p.zz = zz; p.reg_ql = objective_reg_ql();
    volatile unsigned long long* c = (volatile unsigned long long*)sv.host;
    c[2] = sv.gen * 4 + SRV_RUNNING;
    c[SRV_REQ] = sv.seq;
    __sync_synchronize();
    if (mi == 0) objective_server_kernel<MODEL_COMPLEX3><<<1, threads, smem, sv.st>>>(p);
    else objective_server_kernel<MODEL_REAL2><<<1, threads, smem, sv.st>>>(p);
    rc::note_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sv.launched = true;
    sv.model = model; sv.N = N; sv.in = in; sv.out = out; sv.zz = zz; sv.threads = threads;
    return cudaSuccess;
}
// asks a live server to leave and waits for the stream to drain
static void server_quit(Server& sv) {
    if (!sv.launched) return;
    if (server_alive(sv)) {
        volatile unsigned long long* c = (volatile unsigned long long*)sv.host;
        c[SRV_HDR] = 8ull << 32;
        __sync_synchronize();
        c[SRV_REQ] = ++sv.seq;
    }
    cudaStreamSynchronize(sv.st);
    sv.launched = false;
}
}  // namespace rc

// 0 = served; -1 = not applicable / server unusable (the caller takes the one-shot path); > 0 = error code
static int objective_host_server(const double* x_host, int nspin, int inspin, int outspin, const double* rows_host, int64_t m,
                                 int model, int zz, double dkw_eps, double* fids_host, double* stats_host, double* amps_host,
                                 int K) {
    Server* svp = server_slot();
    if (!svp || svp->broken) return -1;
    Server& sv = *svp;
    // CTA size: at least 128 lanes (more lanes = more reads of the rows in flight through the mapping: 30 rows 20.4 ->
    // 17.9 us at N = 5; the nominal call does not care), never smaller than a live evaluator of the same chain
    const int need_threads = (int)((m + 31) / 32 * 32);
    const bool same = sv.launched && sv.model == model && sv.N == nspin && sv.in == inspin && sv.out == outspin && sv.zz == zz;
    auto smem_for = [&](int t) { return ((size_t)4 * nspin * t + (size_t)(nspin + 2) + (size_t)t * K + (size_t)t) * 8; };
    int threads = need_threads < 128 ? 128 : need_threads;
    if (same && sv.threads > threads) threads = sv.threads;
    if (smem_for(threads) > OBJ_SMEM_MAX) threads = need_threads;
    if (smem_for(threads) > OBJ_SMEM_MAX) return -1;
    const size_t words = SRV_IN + srv_out_offset(nspin, threads, K) + (size_t)threads + RC_NUM_STATS + 2 * (size_t)threads;
    if (!sv.st && cudaStreamCreateWithFlags(&sv.st, cudaStreamNonBlocking) != cudaSuccess) { sv.broken = true; return -1; }
    if (sv.bytes < words * 8) {
        server_quit(sv);
        if (sv.host) cudaFreeHost(sv.host);
        sv.host = sv.dev = nullptr; sv.bytes = 0;
        size_t cap = 1 << 16;
        while (cap < words * 8) cap *= 2;
        if (cudaHostAlloc((void**)&sv.host, cap, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&sv.dev, sv.host, 0) != cudaSuccess) { sv.broken = true; cudaGetLastError(); return -1; }
        memset(sv.host, 0, cap);
        sv.bytes = cap;
        sv.seq = 0;
    }
    if (sv.launched && !(same && sv.threads == threads)) server_quit(sv);
    volatile unsigned long long* c = (volatile unsigned long long*)sv.host;
    if (!server_alive(sv)) {
        if (sv.launched) { cudaStreamSynchronize(sv.st); sv.launched = false; }   // gone: idle timeout
        if (server_launch(sv, model, nspin, inspin, outspin, zz, threads, K) != cudaSuccess) { sv.broken = true; cudaGetLastError(); return -1; }
    }
    // the request: header, inputs, then the sequence number (x86 keeps the store order; the device reads the
    // inputs only after it has seen the new sequence number)
    double* hp = (double*)sv.host;
    memcpy(hp + SRV_IN, x_host, (size_t)(nspin + 1) * 8);
    hp[SRV_IN + nspin + 1] = rows_host ? 1.0 : 0.0;
    if (rows_host) memcpy(hp + SRV_IN + nspin + 2, rows_host, (size_t)m * K * 8);
    const unsigned flags = (stats_host ? 1u : 0u) | (amps_host ? 2u : 0u) | (rows_host ? 4u : 0u);
    c[SRV_HDR] = (unsigned long long)(unsigned)m | ((unsigned long long)flags << 32);
    hp[SRV_EPS] = dkw_eps;
    __sync_synchronize();
    const unsigned long long seq = ++sv.seq;
    c[SRV_REQ] = seq;
    int w = server_wait(sv, seq, 2.0);
    if (w == 1) {
        // the server left (idle timeout) just before the request arrived: start a new one, which finds it pending
        cudaStreamSynchronize(sv.st);
        sv.seq = seq - 1;
        if (server_launch(sv, model, nspin, inspin, outspin, zz, threads, K) != cudaSuccess) { sv.broken = true; cudaGetLastError(); return -1; }
        sv.seq = seq;
        c[SRV_REQ] = seq;
        w = server_wait(sv, seq, 2.0);
    }
    if (w != 0) {
        // never expected: give the one-shot path the call and stop using the server in this thread
        sv.broken = true;
        c[SRV_HDR] = 8ull << 32; __sync_synchronize(); c[SRV_REQ] = ++sv.seq;
        cudaError_t e = cudaStreamSynchronize(sv.st);
        if (e != cudaSuccess) return set_error(RC_ERR_CUDA, "rc_objective_host: resident evaluator: %s", cudaGetErrorString(e));
        return -1;
    }
    __sync_synchronize();
    const double* op = hp + SRV_IN + srv_out_offset(nspin, rows_host ? m : 0, K);
    const unsigned long long nonconv = c[3];
    if (nonconv) c[3] = 0;
    if (m == 1 && !stats_host && !amps_host) op = hp + 1;   // the one-value answer came with the flag
    if (fids_host) memcpy(fids_host, op, (size_t)m * 8);
    if (stats_host) memcpy(stats_host, op + m, (size_t)RC_NUM_STATS * 8);
    if (amps_host) memcpy(amps_host, op + m + RC_NUM_STATS, 2 * (size_t)m * 8);
    if (nonconv) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu evaluations (NaN written)", nonconv);
    return RC_OK;
}

// Asks the calling thread's resident evaluator (if any) on the current device to leave now instead of after its idle
// time, and waits for it: for callers about to time device work, and before a device reset.
extern "C" int rc_objective_release(void) {
    Server* sv = server_slot();
    if (sv && sv->launched) server_quit(*sv);
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
        if (smem <= OBJ_SMEM_MAX && server_enabled()) {
            const int r = objective_host_server(x_host, nspin, inspin, outspin, rows_host, m, model, zz, dkw_eps, fids_host,
                                                stats_host, amps_host, K);
            if (r >= 0) return r;
        }
        if (smem <= OBJ_SMEM_MAX)
            return objective_host_fast(x_host, nspin, inspin, outspin, rows_host, m, model, zz, dkw_eps, fids_host, stats_host,
                                       amps_host, (cudaStream_t)stream, K, threads, smem);
    }
    return objective_host_general(x_host, nspin, inspin, outspin, rows_host, m, model, zz, dkw_eps, fids_host, stats_host,
                                  amps_host, stream);
}

// The same call through a caller-built frame: one pointer argument instead of thirteen (what a foreign-function layer
// such as ctypes spends per call is proportional to the argument count: ~2 us of a 13 us call).
extern "C" int rc_objective_call(const rc_objective_frame* f) {
    if (!f) return set_error(RC_ERR_NULL, "rc_objective_call: null frame");
    return rc_objective_host(f->x_host, f->nspin, f->inspin, f->outspin, f->rows_host, f->m, f->model, f->zz, f->dkw_eps,
                             f->fids_host, f->stats_host, f->amps_host, f->stream);
}
