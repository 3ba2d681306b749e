// Host-buffer pipeline entry point (what MCDataSim.get_metrics_dict computes from scratch,
// mcsim.py:463-510) and the FP64 roofline micro-benchmark.
#include <stdlib.h>
#include <string.h>
#include "rc_common.cuh"

using namespace rc;

namespace rc {

struct DevBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    template <class T> T* as() { return (T*)p; }
};

// Stream-ordered scratch comes from the device's default memory pool; keep up to 2 GiB of freed blocks cached
// in the pool (default behaviour returns them to the OS at every synchronisation, which makes each
// sweep pay cudaMalloc/cudaFree of the fidelity tensor again).
static cudaError_t keep_pool_memory() {
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || done[dev]) return cudaSuccess;
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    // bounded: the pool keeps at most this much freed memory cached (the scratch of a paper-size sweep is a few
    // hundred MB); anything above goes back to the driver at the next synchronisation, so other users of the
    // default pool (PyTorch's allocator in the same process) are not starved after one large sweep
    unsigned long long thr = 2ull << 30;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

// Second stream + events of the host sweep's copy pipeline, created once per (host thread, device).
struct SweepAsync {
    static constexpr int MAX_CHUNKS = 16;
    cudaStream_t copy = nullptr;
    cudaEvent_t ev[MAX_CHUNKS] = {};
    cudaEvent_t done = nullptr;
};
static cudaError_t sweep_async(SweepAsync** out) {
    static thread_local SweepAsync cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    SweepAsync& a = cache[dev];
    if (!a.copy) {
        e = cudaStreamCreateWithFlags(&a.copy, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        for (int k = 0; k < SweepAsync::MAX_CHUNKS; ++k) {
            e = cudaEventCreateWithFlags(&a.ev[k], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        e = cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    *out = &a;
    return cudaSuccess;
}

// 8 independent dependent-FMA chains per thread; 2 flops per DFMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

}  // namespace rc

extern "C" int rc_fp64_peak_tflops(double* tflops, void* stream) {
    if (!tflops) return set_error(RC_ERR_NULL, "rc_fp64_peak_tflops: null output");
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf out(st);
    RC_CUDA_TRY(out.alloc(8));
    const int sm = device_sm_count();
    const int blocks = sm * 8, threads = 256, iters = 1 << 15;
    cudaEvent_t e0, e1;
    RC_CUDA_TRY(cudaEventCreate(&e0));
    RC_CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        RC_CUDA_TRY(cudaEventRecord(e0, st));
        dfma_peak_kernel<<<blocks, threads, 0, st>>>(out.as<double>(), iters, 0.999999, 1e-7); rc::note_launch();
        RC_CUDA_TRY(cudaEventRecord(e1, st));
        RC_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        RC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// Whole fig-4/5 sweep on DEVICE buffers: every launch of the step issued from one C call (no allocation, no
// synchronisation) — the device-resident twin of rc_robustness_sweep_host.
// ------------------------------------------------------------------------------------------------
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t rc_robustness_sweep_workspace_bytes(int64_t C, int S, int64_t B, int fused, int64_t G, int64_t topk) {
    if (C <= 0 || S <= 0 || B <= 0 || G <= 0 || C % G) return 256;
    const size_t ws_f = fused ? rc_fidelity_stats_workspace_bytes((int64_t)S * C, B) : 0;
    const size_t ws_r = rc_rank_consistency_workspace_bytes(S, G, C / G, topk);
    if (ws_r == 0) return 0;
    return align256(ws_f) + align256(ws_r) + 256;
}

extern "C" int rc_robustness_sweep(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                                   const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                                   int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G, int64_t topk,
                                   double alpha_cluster, double* fids_dev, double* stats_dev, double* tau_dev,
                                   int64_t* sel_dev, double* wsel_dev, int nboot, double* arim_dev, double* arim_std_dev,
                                   unsigned long long* counters_dev, void* workspace_dev, size_t workspace_bytes,
                                   void* ev_evolution_begin, void* ev_evolution_end, void* stream) {
    if (C < 0 || S < 0 || B < 1) return set_error(RC_ERR_BAD_ARG, "rc_robustness_sweep: bad sizes C=%lld S=%d B=%lld", (long long)C, S, (long long)B);
    if ((long long)S * C == 0) return RC_OK;
    if (G < 1 || C % G) return set_error(RC_ERR_BAD_ARG, "rc_robustness_sweep: C=%lld is not a multiple of G=%lld", (long long)C, (long long)G);
    if (!stats_dev || !tau_dev || !sel_dev || !wsel_dev) return set_error(RC_ERR_NULL, "rc_robustness_sweep: null stats/tau/sel/wsel output");
    if (!fused && !fids_dev) return set_error(RC_ERR_NULL, "rc_robustness_sweep: the materialising path needs fids_dev");
    const size_t need = rc_robustness_sweep_workspace_bytes(C, S, B, fused, G, topk);
    if (need == 0) return set_error(RC_ERR_BAD_ARG, "rc_robustness_sweep: ranking problem too large");
    if (!workspace_dev || workspace_bytes < need)
        return set_error(RC_ERR_WORKSPACE, "rc_robustness_sweep: workspace %zu < required %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nseg = (long long)S * C;
    unsigned long long* nonconv = counters_dev;
    unsigned long long* illegal = counters_dev ? counters_dev + 1 : nullptr;
    char* ws = (char*)workspace_dev;
    const size_t ws_f = fused ? rc_fidelity_stats_workspace_bytes(nseg, B) : 0;
    int rcode;
    // optional caller-owned events around the evolution launch(es): lets a benchmark time the dominant kernel
    // inside its timed region without splitting the step into several calls
    if (ev_evolution_begin) RC_CUDA_TRY(cudaEventRecord((cudaEvent_t)ev_evolution_begin, st));
    if (fused) {
        rcode = rc_fidelity_stats(ctrl_dev, C, nspin, inspin, outspin, sigma_dev, S, B, model, zz, seed, c_offset, b_offset,
                                  nullptr, dkw_eps, stats_dev, nonconv, ws, ws_f, st);
        if (rcode) return rcode;
        if (ev_evolution_end) RC_CUDA_TRY(cudaEventRecord((cudaEvent_t)ev_evolution_end, st));
    } else {
        rcode = fidelity_mc_impl("rc_robustness_sweep", ctrl_dev, C, nspin, inspin, outspin, sigma_dev, S, B, model, zz, seed,
                                 c_offset, b_offset, nullptr, fids_dev, nonconv, 0, st);
        if (rcode) return rcode;
        if (ev_evolution_end) RC_CUDA_TRY(cudaEventRecord((cudaEvent_t)ev_evolution_end, st));
        rcode = stats_unsorted_impl(fids_dev, nseg, B, dkw_eps, stats_dev, nseg, illegal, st);
        if (rcode) return rcode;
    }
    char* ws_rank = ws + align256(ws_f);
    rcode = rc_rank_consistency(stats_dev, S, G, C / G, topk, alpha_cluster, tau_dev, sel_dev, wsel_dev, ws_rank,
                                rc_rank_consistency_workspace_bytes(S, G, C / G, topk), st);
    if (rcode) return rcode;
    if (arim_dev && arim_std_dev) {
        const int64_t Cg = C / G, k = topk < Cg ? topk : Cg;
        rcode = rc_arim_bootstrap(wsel_dev, G * S, k, nboot > 0 ? nboot : 100, seed ^ 0x9E3779B97F4A7C15ull, arim_dev, arim_std_dev, st);
    }
    return rcode;
}

static int sweep_host_impl(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                           const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                           int64_t c_offset, int64_t b_offset, const double* replay_host, double dkw_eps,
                           int fused, double* fids_host, double* stats_host, int64_t G, int64_t topk,
                           double alpha_cluster, double* tau_host, int64_t* sel_host, int nboot, double* arim_host,
                           double* arim_std_host, void* stream, double* stats_dev_keep = nullptr) {
    if (nspin < 2 || nspin > RC_MAX_NSPIN) return set_error(RC_ERR_BAD_ARG, "nspin=%d outside [2,%d]", nspin, RC_MAX_NSPIN);
    if (C < 0 || S < 0 || B < 1) return set_error(RC_ERR_BAD_ARG, "sweep: bad sizes C=%lld S=%d B=%lld", (long long)C, S, (long long)B);
    const long long nseg = (long long)S * C, total = nseg * B;
    if (total == 0) return RC_OK;
    if (!ctrl_host || !sigma_host) return set_error(RC_ERR_NULL, "sweep: null ctrl/sigma");
    if (!stats_host && !fids_host) return set_error(RC_ERR_NULL, "sweep: no output requested");
    if (fused && fids_host) return set_error(RC_ERR_BAD_ARG, "sweep: fused mode does not materialise fidelities");
    const bool want_rank = tau_host != nullptr;
    if (want_rank && (G < 1 || C % G)) return set_error(RC_ERR_BAD_ARG, "sweep: C=%lld is not a multiple of G=%lld", (long long)C, (long long)G);
    cudaStream_t st = (cudaStream_t)stream;
    RC_CUDA_TRY(keep_pool_memory());
    const int K = (model == RC_MODEL_COMPLEX3 ? 3 : 2) * nspin;
    DevBuf ctrl(st), sigma(st), replay(st), fids(st), stats(st), ws(st), counters(st), tau(st), sel(st), wsel(st), rws(st), ar(st);
    RC_CUDA_TRY(ctrl.alloc((size_t)C * (nspin + 1) * 8));
    RC_CUDA_TRY(sigma.alloc((size_t)S * 8));
    RC_CUDA_TRY(counters.alloc(16));
    RC_CUDA_TRY(cudaMemsetAsync(counters.p, 0, 16, st));
    RC_CUDA_TRY(cudaMemcpyAsync(ctrl.p, ctrl_host, (size_t)C * (nspin + 1) * 8, cudaMemcpyHostToDevice, st));
    RC_CUDA_TRY(cudaMemcpyAsync(sigma.p, sigma_host, (size_t)S * 8, cudaMemcpyHostToDevice, st));
    if (replay_host) {
        RC_CUDA_TRY(replay.alloc((size_t)total * K * 8));
        RC_CUDA_TRY(cudaMemcpyAsync(replay.p, replay_host, (size_t)total * K * 8, cudaMemcpyHostToDevice, st));
    }
    unsigned long long* nonconv = counters.as<unsigned long long>();
    unsigned long long* illegal = nonconv + 1;
    // stats_dev_keep: caller-owned device tensor [15][S][C] that receives the statistics and stays valid after the
    // call (the multi-GPU sweep pushes it to its peers while the next call runs)
    if (stats_dev_keep) stats.p = stats_dev_keep;
    else if (stats_host || want_rank) RC_CUDA_TRY(stats.alloc((size_t)RC_NUM_STATS * nseg * 8));
    struct Unown { DevBuf& b; bool on; ~Unown() { if (on) b.p = nullptr; } } unown{stats, stats_dev_keep != nullptr};
    int rcode;
    SweepAsync* async = nullptr;
    cudaStream_t copy = nullptr;
    bool stats_copied = false;
    // every exit path (error returns included): the stream-ordered frees of the scratch buffers above are issued on
    // st by the DevBuf destructors, so st must first wait for the copy stream that may still be reading them
    struct CopyJoin {
        cudaStream_t st; cudaStream_t* copy; SweepAsync** async;
        ~CopyJoin() {
            if (*copy && *async && cudaEventRecord((*async)->done, *copy) == cudaSuccess) cudaStreamWaitEvent(st, (*async)->done, 0);
        }
    } copy_join{st, &copy, &async};
    if (fused) {
        size_t wb = rc_fidelity_stats_workspace_bytes(nseg, B);
        RC_CUDA_TRY(ws.alloc(wb));
        rcode = rc_fidelity_stats(ctrl.as<double>(), C, nspin, inspin, outspin, sigma.as<double>(), S, B, model, zz, seed,
                                  c_offset, b_offset, replay_host ? replay.as<double>() : nullptr, dkw_eps,
                                  stats.as<double>(), nonconv, ws.p, wb, st);
        if (rcode) return rcode;
    } else {
        // Materialising path: evolution, then the sort-free statistics pass over the fresh fidelities (mostly
        // still in L2).  Large sweeps are cut into a few sigma chunks so that the D2H copies of chunk k (its
        // columns of the statistics tensor, its slab of the fidelity tensor) run on a second stream underneath
        // the evolution of chunk k+1; results do not depend on the chunking (global sigma indices feed the
        // Philox counters).
        const bool want_stats = stats_host || want_rank;
        RC_CUDA_TRY(fids.alloc((size_t)total * 8));
        int nch = 1;
        if (S >= 2 && total >= 2000000) {
            nch = S < 4 ? S : 4;
            if (const char* e = getenv("RC_SWEEP_CHUNKS")) { int v = atoi(e); if (v >= 1) nch = v < S ? v : S; }
            if (nch > SweepAsync::MAX_CHUNKS) nch = SweepAsync::MAX_CHUNKS;
        }
        if (nch > 1) {
            RC_CUDA_TRY(sweep_async(&async));
            copy = async->copy;
        }
        for (int k = 0; k < nch; ++k) {
            // ceil-based boundaries: the remainder goes to the FIRST chunks, so the last chunk — whose D2H copy nothing
            // overlaps — is the smallest (S = 11, 4 chunks: 3, 3, 3, 2 levels; a single-level last chunk measured no better)
            const long long s0 = ((long long)S * k + nch - 1) / nch, s1 = ((long long)S * (k + 1) + nch - 1) / nch, Sk = s1 - s0;
            const long long e0 = s0 * C * B;
            rcode = fidelity_mc_impl("sweep", ctrl.as<double>(), C, nspin, inspin, outspin, sigma.as<double>() + s0, (int)Sk,
                                     B, model, zz, seed, c_offset, b_offset,
                                     replay_host ? replay.as<double>() + e0 * K : nullptr, fids.as<double>() + e0, nonconv,
                                     (int)s0, st);
            if (rcode) return rcode;
            if (want_stats) {
                rcode = stats_unsorted_impl(fids.as<double>() + e0, Sk * C, B, dkw_eps, stats.as<double>() + s0 * C, nseg,
                                            illegal, st);
                if (rcode) return rcode;
            }
            cudaStream_t cs = st;
            if (nch > 1) {
                RC_CUDA_TRY(cudaEventRecord(async->ev[k], st));
                RC_CUDA_TRY(cudaStreamWaitEvent(copy, async->ev[k], 0));
                cs = copy;
            }
            // the reference dumps the UNSORTED tensor to .mc before any metric sorts it (mcsim.py:457-459)
            if (fids_host)
                RC_CUDA_TRY(cudaMemcpyAsync(fids_host + e0, fids.as<double>() + e0, (size_t)Sk * C * B * 8, cudaMemcpyDeviceToHost, cs));
            if (stats_host)
                RC_CUDA_TRY(cudaMemcpy2DAsync(stats_host + s0 * C, (size_t)nseg * 8, stats.as<double>() + s0 * C, (size_t)nseg * 8,
                                              (size_t)Sk * C * 8, RC_NUM_STATS, cudaMemcpyDeviceToHost, cs));
        }
        stats_copied = true;
    }
    if (stats_host && !stats_copied)
        RC_CUDA_TRY(cudaMemcpyAsync(stats_host, stats.p, (size_t)RC_NUM_STATS * nseg * 8, cudaMemcpyDeviceToHost, st));
    if (want_rank) {
        const int64_t Cg = C / G, k = topk < Cg ? topk : Cg;
        size_t wb = rc_rank_consistency_workspace_bytes(S, G, Cg, topk);
        if (wb == 0) return set_error(RC_ERR_BAD_ARG, "sweep: ranking problem too large");
        RC_CUDA_TRY(rws.alloc(wb));
        RC_CUDA_TRY(tau.alloc((size_t)G * S * S * 8));
        RC_CUDA_TRY(sel.alloc((size_t)G * k * 8));
        RC_CUDA_TRY(wsel.alloc((size_t)G * S * k * 8));
        rcode = rc_rank_consistency(stats.as<double>(), S, G, Cg, topk, alpha_cluster, tau.as<double>(), sel.as<int64_t>(),
                                    wsel.as<double>(), rws.p, wb, st);
        if (rcode) return rcode;
        RC_CUDA_TRY(cudaMemcpyAsync(tau_host, tau.p, (size_t)G * S * S * 8, cudaMemcpyDeviceToHost, st));
        if (sel_host) RC_CUDA_TRY(cudaMemcpyAsync(sel_host, sel.p, (size_t)G * k * 8, cudaMemcpyDeviceToHost, st));
        if (arim_host && arim_std_host) {
            RC_CUDA_TRY(ar.alloc((size_t)2 * G * S * 8));
            rcode = rc_arim_bootstrap(wsel.as<double>(), G * S, k, nboot > 0 ? nboot : 100, seed ^ 0x9E3779B97F4A7C15ull,
                                      ar.as<double>(), ar.as<double>() + G * S, st);
            if (rcode) return rcode;
            RC_CUDA_TRY(cudaMemcpyAsync(arim_host, ar.p, (size_t)G * S * 8, cudaMemcpyDeviceToHost, st));
            RC_CUDA_TRY(cudaMemcpyAsync(arim_std_host, ar.as<double>() + G * S, (size_t)G * S * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    unsigned long long hc[2] = {0, 0};
    RC_CUDA_TRY(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, st));
    if (copy) {   // the stream-ordered frees of the scratch buffers (on st) must follow the copies that read them
        RC_CUDA_TRY(cudaEventRecord(async->done, copy));
        RC_CUDA_TRY(cudaStreamWaitEvent(st, async->done, 0));
    }
    RC_CUDA_TRY(cudaStreamSynchronize(st));
    if (hc[0]) return set_error(RC_ERR_NONCONV, "eigensolver did not converge for %llu evaluations (NaN written)", hc[0]);
    if (hc[1]) return set_error(RC_ERR_ILLEGAL_FIDS, "illegal fids values - must be in [0,1] (%llu samples)", hc[1]);
    return RC_OK;
}

extern "C" int rc_mc_sweep_host(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                                const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                                int64_t c_offset, int64_t b_offset, const double* replay_host, double dkw_eps,
                                int fused, double* fids_host, double* stats_host, void* stream) {
    return sweep_host_impl(ctrl_host, C, nspin, inspin, outspin, sigma_host, S, B, model, zz, seed, c_offset, b_offset,
                           replay_host, dkw_eps, fused, fids_host, stats_host, 0, 0, 0.0, nullptr, nullptr, 0, nullptr, nullptr,
                           stream);
}

extern "C" int rc_robustness_sweep_host(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                                        const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                                        int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G,
                                        int64_t topk, double alpha_cluster, double* stats_host, double* tau_host,
                                        int64_t* sel_host, int nboot, double* arim_host, double* arim_std_host,
                                        void* stream) {
    if (!tau_host) return set_error(RC_ERR_NULL, "rc_robustness_sweep_host: null tau output");
    return sweep_host_impl(ctrl_host, C, nspin, inspin, outspin, sigma_host, S, B, model, zz, seed, c_offset, b_offset,
                           nullptr, dkw_eps, fused, nullptr, stats_host, G, topk, alpha_cluster, tau_host, sel_host, nboot,
                           arim_host, arim_std_host, stream);
}

extern "C" int rc_robustness_sweep_host_keep(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                                             const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                                             int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G,
                                             int64_t topk, double alpha_cluster, double* stats_host, double* tau_host,
                                             int64_t* sel_host, int nboot, double* arim_host, double* arim_std_host,
                                             double* stats_dev_keep, void* stream) {
    if (!tau_host) return set_error(RC_ERR_NULL, "rc_robustness_sweep_host_keep: null tau output");
    if (!stats_dev_keep) return set_error(RC_ERR_NULL, "rc_robustness_sweep_host_keep: null device statistics tensor");
    return sweep_host_impl(ctrl_host, C, nspin, inspin, outspin, sigma_host, S, B, model, zz, seed, c_offset, b_offset,
                           nullptr, dkw_eps, fused, nullptr, stats_host, G, topk, alpha_cluster, tau_host, sel_host, nboot,
                           arim_host, arim_std_host, stream, stats_dev_keep);
}

