/* robchar_b200 — C-ABI of the B200-native RobChar Monte-Carlo robustness hot path.
 *
 * The reference (qyber-black/Code-RobChar) has no FFI layer: its boundary is Python call
 * signatures.  Each entry point below names the reference interface it stands behind
 * (file:line in the upstream repo).  All pointers are plain; "dev" pointers are CUDA device
 * addresses (any CUDA allocation), "host" pointers are ordinary host memory.  `stream`
 * is a cudaStream_t passed as void* (NULL = default stream).  Every function returns an int
 * status (RC_OK = 0) and records a message retrievable with rc_last_error() (thread local).
 * The library owns no persistent state; scratch memory is caller-provided (size-query calls) in
 * the device API and stream-ordered internal allocations in the *_host API.
 *
 * Tensor layouts (row-major, float64 unless noted):
 *   ctrl    [C][N+1]        N biases then the evolution time (sign ignored, noise_model.py:99)
 *   sigma   [S]             simulation noise levels (mcsim.py:204, 424-425)
 *   replay  [S][C][B][K]    STANDARD normals in the reference's draw order, K = 3N (complex model,
 *                           (z_ii, nn_i, nn2_i) per site, noise_model.py:135-147) or 2N (real
 *                           model, qnewton.py:366-379); the device scales them by sigma
 *   fids    [S][C][B]       fidelity samples (mcsim.py:423 `allfids`)
 *   stats   [15][S][C]      metric tensors in the reference's .mcm key order (mcsim.py:178-183,
 *                           487-498): for each of W, Q0.95, Q0.98, std, worst-case: centre,
 *                           upper, lower
 */
#ifndef ROBCHAR_B200_H
#define ROBCHAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_OK 0
#define RC_ERR_BAD_ARG 1      /* bad nspin / inspin / outspin / sizes / model            */
#define RC_ERR_NULL 2         /* required pointer is NULL                                */
#define RC_ERR_CUDA 3         /* CUDA runtime error (message in rc_last_error)           */
#define RC_ERR_NONCONV 4      /* eigensolver non-convergence count > 0 (host API only)   */
#define RC_ERR_ILLEGAL_FIDS 5 /* fidelity outside [0,1]: wd_sortof_fast_implementation.py:23-25 */
#define RC_ERR_WORKSPACE 6    /* workspace too small                                     */

#define RC_MODEL_COMPLEX3 0 /* structured_perturbation.perturbation, noise_model.py:122-147 */
#define RC_MODEL_REAL2 1    /* LBFGS.structured_perturabation, qnewton.py:366-379            */

#define RC_NUM_STATS 15
#define RC_MAX_NSPIN 32

int rc_version(void);
const char* rc_last_error(void);
/* Number of kernels of THIS library launched by the process so far (counted at the launch sites; library kernels
 * such as the CUB segmented sort of the B > 4096 statistics path are not included). */
unsigned long long rc_launch_count(void);
/* SM count and compute capability of the current device. */
int rc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Evolution + fidelity for every (sigma level, controller, draw).
 * Stands behind MCDataSim.get_algo_fid_dist's triple loop (mcsim.py:422-456) calling
 * noise_model_base.evaluate_noisy_fidelity(x, ham_noisy=True) (noise_model.py:98-109), and
 * LBFGS.fidelity_ss(x, ham_noisy=True) (qnewton.py:383-400) for RC_MODEL_REAL2 / zz.
 * replay_dev == NULL: noise from in-kernel Philox4x32-10 keyed by `seed`, counters from the
 * GLOBAL indices (sigma idx, c_offset + c, b_offset + b) — sharding invariant; the 64 bits of each
 * draw feed a 1024-layer ziggurat (exact N(0,1); csrc/rc_philox.cuh).
 * NaN controller rows give NaN fidelities (mcsim.py:369-374, 443).
 * nonconv_dev: optional device counter (uint64) incremented per non-converged evaluation. */
int rc_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                   const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                   int64_t c_offset, int64_t b_offset, const double* replay_dev, double* fids_dev,
                   unsigned long long* nonconv_dev, void* stream);

/* rc_fidelity_mc followed by rc_stats_unsorted on the same stream: the fidelity tensor
 * (get_algo_fid_dist, mcsim.py:422-460) and the 15 metric tensors get_metrics_dict derives from it
 * (mcsim.py:463-510).  fids_dev [S][C][B] (the unsorted `.mc` tensor), stats_dev [15][S][C]. */
int rc_fidelity_mc_stats(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                         const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                         int64_t c_offset, int64_t b_offset, const double* replay_dev, double dkw_eps,
                         double* fids_dev, double* stats_dev, unsigned long long* nonconv_dev,
                         unsigned long long* illegal_dev, void* stream);

/* Name of the evolution kernel the launcher picks for this chain length / mode (register family N <= 12,
 * shared-memory family above, eigenvector rows or spectral weights): what a benchmark labels its roofline with. */
int rc_evolution_kernel_name(int nspin, int replay, int fused, char* buf, size_t buf_bytes);

/* Chain lengths above the register-resident range (N >= 13) evaluate <out|exp(-iHT)|in> from the eigenvalues
 * alone (characteristic-polynomial weights, csrc/rc_spectral.cuh) and recompute an evaluation with accumulated
 * eigenvector rows when its a-posteriori error estimate exceeds 1e-11 (near-coincident eigenvalues with large
 * weights).  This diagnostic returns how many evaluations took that fallback on the current device since the
 * last reset (synchronises `stream` when count_host != NULL). */
int rc_spectral_fallbacks(unsigned long long* count_host, int reset, void* stream);

/* The standard normals rc_fidelity_mc's Philox mode uses, written in replay layout [S][C][B][K]
 * (discarded site-0 coupling slots are zero).  Lets callers replay a GPU sweep through the
 * reference CPU path. */
int rc_philox_normals(int64_t C, int nspin, int S, int64_t B, int model, uint64_t seed, int64_t c_offset,
                      int64_t b_offset, double* normals_dev, void* stream);

/* Statistics of each (sigma, controller) segment of B samples: segmented sort + fused
 * 1-Wasserstein/RIM, Q-threshold, std and worst-case reductions for the centre and the two
 * DKW-shifted variants.  Stands behind wd_from_ideal (wd_sortof_fast_implementation.py:82-116),
 * the metric registry (mcsim.py:144-183) and get_metrics_dict's DKW loop (mcsim.py:482-498).
 * dkw_eps = compute_dkw_error(alpha, B) (wd_sortof_fast_implementation.py:38-39); pass 0 for none.
 * sorted_dev: optional [nseg][B] output of the ascending-sorted samples (may alias fids_dev:
 * wd_from_ideal sorts its argument in place).  illegal_dev: optional uint64 device counter of
 * samples violating |f - 1e-8| <= 1.  stats_dev: [15][nseg]. */
size_t rc_stats_workspace_bytes(int64_t nseg, int64_t B);
int rc_stats(const double* fids_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
             double* sorted_dev, unsigned long long* illegal_dev, void* workspace_dev, size_t workspace_bytes,
             void* stream);

/* The same 15 statistics WITHOUT sorting: none of them depends on the order of the samples once W is
 * written as mean(1 - f), the value of wd_from_ideal's sorted telescoping sum
 * (wd_sortof_fast_implementation.py:105-114) up to rounding; two-pass population std, exact threshold
 * counts and minimum.  One streaming pass over fids_dev (8 B/sample, HBM bound) instead of a sort:
 * what the sweep entry points use.  fids_dev is not modified.  stats_dev: [15][nseg]. */
int rc_stats_unsorted(const double* fids_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                      unsigned long long* illegal_dev, void* stream);

/* p-RIM of each segment: (mean((1 - f)^p))^(1/p), p = 0 -> 1 (RIM_p, wd_sortof_fast_implementation.py:147-174).
 * fids_dev [nseg][B] -> out_dev [nseg]; illegal_dev as in rc_stats. */
int rc_rim_p(const double* fids_dev, int64_t nseg, int64_t B, double p, double* out_dev, unsigned long long* illegal_dev,
             void* stream);

/* Fused evolution + statistics that never materialises the fidelity tensor (streaming moments;
 * W = mean(1 - f), identical to the sorted formula up to rounding).  Same arguments as
 * rc_fidelity_mc; stats_dev [15][S][C].  workspace: rc_fidelity_stats_workspace_bytes(S*C). */
size_t rc_fidelity_stats_workspace_bytes(int64_t nseg, int64_t B);
int rc_fidelity_stats(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                      const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                      int64_t c_offset, int64_t b_offset, const double* replay_dev, double dkw_eps,
                      double* stats_dev, unsigned long long* nonconv_dev, void* workspace_dev,
                      size_t workspace_bytes, void* stream);

/* Draw-sharded sweep (fewer controllers than GPUs: NStochOpt.get_rims, gen_fig_8_arim_fcall_scaling.py:121-132 —
 * one controller x B draws per sigma level — and LBFGS.wass_cost, qnewton.py:447-455).  The chunk partials of a
 * segment are always merged in the same fixed order: eight contiguous blocks of chunks, each merged by one warp,
 * the eight results merged sequentially.  Rank r of `world` (1, 2, 4 or 8) owns blocks [8r/world, 8(r+1)/world),
 * i.e. the draws [b_lo, b_hi) rc_draw_shard_range reports (Philox counters use the global draw index).
 * rc_fidelity_stats_blocks runs the fused evolution + streaming statistics on that range and writes the rank's
 * block results blocks_dev [8/world][S*C][17]; all-gathered in rank order they form [8][S*C][17], which
 * rc_stats_from_blocks turns into stats_dev [15][S*C] — bit-identical to rc_fidelity_stats on one GPU. */
int rc_draw_shard_range(int64_t B, int world, int rank, int64_t* b_lo, int64_t* b_hi, int* v_lo, int* v_hi);
size_t rc_fidelity_stats_blocks_workspace_bytes(int64_t nseg, int64_t B, int world);
int rc_fidelity_stats_blocks(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                             const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                             int64_t c_offset, int64_t b_offset, int world, int rank, double dkw_eps,
                             double* blocks_dev, unsigned long long* nonconv_dev, void* workspace_dev,
                             size_t workspace_bytes, void* stream);
int rc_stats_from_blocks(const double* blocks_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                         void* stream);

/* Peer exchange of the per-rank statistics blocks (controller-sharded sweep, one process per GPU): what assembles
 * MCDataSim.get_metrics_dict's tensors (mcsim.py:463-510) for the whole controller set from the shards.  The
 * reference has no multi-device path.  rc_peer_alloc: cudaMalloc'ed, zeroed exchange buffer + its 64-byte CUDA IPC
 * handle; rc_peer_open maps a peer's buffer from its handle (NVLink peer access enabled lazily); rc_peer_close /
 * rc_peer_free undo them.  rc_peer_push_columns: local_dev [rows][c_local] -> columns [col_offset, col_offset +
 * c_local) of the [rows][c_total] tensor at peer_tensors[r] of EVERY rank r (host array of `world` device addresses,
 * this rank's own buffer included): one 2-D copy-engine transfer per destination on `stream`, no kernel.
 * rc_peer_signal: afterwards, writes `seq` into slot `rank` of every rank's flag array (peer_flags[r], uint64[world]).
 * rc_peer_wait: on the consumer's stream, waits until all `world` slots of local_flags have reached `seq`; after
 * timeout_s seconds it gives up and increments *timed_out_dev (uint64, optional) instead of hanging the GPU. */
int rc_peer_alloc(size_t bytes, void** dev_ptr_out, unsigned char* handle64_out);
int rc_peer_open(const unsigned char* handle64, void** dev_ptr_out);
int rc_peer_close(void* dev_ptr);
int rc_peer_free(void* dev_ptr);
int rc_peer_push_columns(void* const* peer_tensors, int world, const double* local_dev, int64_t rows, int64_t c_local,
                         int64_t c_total, int64_t col_offset, void* stream);
int rc_peer_signal(void* const* peer_flags, int world, int rank, uint64_t seq, void* stream);
int rc_peer_wait(const void* local_flags, int world, uint64_t seq, double timeout_s, void* timed_out_dev, void* stream);

/* Ordinal ranks 0..n-1 of each row, ascending, NaN last, ties by index (stable).
 * MCDataSim.get_ranks (mcsim.py:513-518).  values [R][n] -> ranks int64 [R][n]. */
size_t rc_ranks_workspace_bytes(int64_t R, int64_t n);
int rc_ranks(const double* values_dev, int64_t R, int64_t n, int64_t* ranks_dev, void* workspace_dev,
             size_t workspace_bytes, void* stream);

/* Clustered ranks with discrepancy radius r_row = alpha * (max - min) of the row, or the fixed
 * radius `r_fixed` when alpha < 0.  get_ranks_clustered_little
 * (generate_fig4_kendallrankanalysis.py:146-164).  values [R][n] -> double [R][n]. */
int rc_clustered_ranks(const double* values_dev, int64_t R, int64_t n, double alpha, double r_fixed,
                       double* cranks_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Kendall tau-b matrix tau[j][i] between x rows j (double ranks, ties allowed) and y rows i
 * (int64 ranks): scipy.stats.kendalltau as called by jkt_or_ordinaltau_pairwise
 * (generate_fig4_kendallrankanalysis.py:94-120).  Integer pair counts, then
 * (tot - xtie - ytie + ntie - 2 dis) / sqrt(tot - xtie) / sqrt(tot - ytie).
 * counts_dev: scratch int64 [Rx][Ry][4]. */
int rc_kendall_tau_b(const double* x_dev, int64_t Rx, const int64_t* y_dev, int64_t Ry, int64_t n,
                     double* tau_dev, long long* counts_dev, void* stream);

/* Same for G independent controller groups at once: x [G][Rx][n], y [G][Ry][n] -> tau [G][Rx][Ry]
 * (one group = one (algorithm, training noise) controller set of the paper's fig. 4 sweep).
 * counts_dev: scratch int64 [G][Rx][Ry][4]. */
int rc_kendall_tau_b_batched(const double* x_dev, const int64_t* y_dev, int64_t G, int64_t Rx, int64_t Ry,
                             int64_t n, double* tau_dev, long long* counts_dev, void* stream);

/* The same matrices for LONG rank vectors (top-k up to 1e5) in O(n log^2 n): dense labels by segmented
 * radix sort, tie counts from the sorted runs, discordant pairs = inversions counted by parallel merge passes
 * (Knight's algorithm, what SciPy itself uses).  Integer counts => bit-identical to rc_kendall_tau_b_batched.
 * workspace: rc_kendall_large_workspace_bytes(G, Rx, Ry, n) (0 = more than 2^31 elements). */
size_t rc_kendall_large_workspace_bytes(int64_t G, int64_t Rx, int64_t Ry, int64_t n);
int rc_kendall_tau_b_large(const double* x_dev, const int64_t* y_dev, int64_t G, int64_t Rx, int64_t Ry, int64_t n,
                           double* tau_dev, long long* counts_dev, void* workspace_dev, size_t workspace_bytes,
                           void* stream);

/* The paper's fig-4 rank-consistency analysis for G controller groups in one call.
 * W_dev: RIM matrix [S][G*Cg] (row 0 of the statistics tensor).  Per group: keep the topk
 * controllers with the smallest RIM at sigma index 0, in their original column order
 * (get_top_k_by_fid, mcsim.py:651-660 with fid_thres=None), then the S x S Kendall matrix of
 * clustered ranks (radius alpha*(max-min)) against ordinal ranks + 1
 * (jkt_or_ordinaltau_pairwise, generate_fig4_kendallrankanalysis.py:94-120).
 * Outputs: tau_dev [G][S][S], sel_dev int64 [G][k] (column index inside the group),
 * Wsel_dev [G][S][k], with k = min(topk, Cg). */
size_t rc_rank_consistency_workspace_bytes(int64_t S, int64_t G, int64_t Cg, int64_t topk);
int rc_rank_consistency(const double* W_dev, int64_t S, int64_t G, int64_t Cg, int64_t topk, double alpha,
                        double* tau_dev, int64_t* sel_dev, double* Wsel_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream);

/* ARIM (algorithm-level RIM) of every row of rims_dev [R][k] and its bootstrap error bar:
 * arim[r] = wd_from_ideal_zero(row) = mean(row) (generate_arim_all_fig5.py:119,166), std[r] = population
 * std of that statistic over `nboot` resamples with replacement (MCDataSim.bootstrap_resampling_std,
 * mcsim.py:267-275), resampling indices from Philox4x32-10 keyed by `seed`. */
int rc_arim_bootstrap(const double* rims_dev, int64_t R, int64_t k, int nboot, uint64_t seed, double* arim_dev,
                      double* std_dev, void* stream);

/* Whole sweep with HOST buffers: H2D of controllers/sigmas(/replay), evolution, statistics, D2H of
 * the 15 metric tensors (and of the fidelity tensor when fids_host != NULL).  This is what
 * MCDataSim.get_metrics_dict (mcsim.py:463-510) computes from scratch.
 * Returns RC_ERR_NONCONV / RC_ERR_ILLEGAL_FIDS after completing if either counter is non-zero. */
int rc_mc_sweep_host(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                     const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                     int64_t c_offset, int64_t b_offset, const double* replay_host, double dkw_eps,
                     int fused, double* fids_host, double* stats_host, void* stream);

/* rc_mc_sweep_host followed by rc_rank_consistency on the device, everything returned to HOST
 * buffers: stats_host [15][S][C], tau_host [G][S][S], sel_host int64 [G][k] (C = G*Cg), and, when
 * arim_host != NULL, the fig-5 ARIM [G][S] with its bootstrap std [G][S] over `nboot` resamples of the
 * top-k RIMs.  This is the whole fig-4/5 sweep of one problem as a single call (pinned host buffers
 * make the copies asynchronous up to the final synchronisation). */
int rc_robustness_sweep_host(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                             const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                             int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G,
                             int64_t topk, double alpha_cluster, double* stats_host, double* tau_host,
                             int64_t* sel_host, int nboot, double* arim_host, double* arim_std_host, void* stream);

/* rc_robustness_sweep_host that ALSO leaves the statistics in the caller's device tensor stats_dev_keep [15][S][C]
 * (valid after the call): the multi-GPU sweep pushes that block to its peers (rc_peer_push_columns) while the next
 * call computes. */
int rc_robustness_sweep_host_keep(const double* ctrl_host, int64_t C, int nspin, int inspin, int outspin,
                                  const double* sigma_host, int S, int64_t B, int model, int zz, uint64_t seed,
                                  int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G,
                                  int64_t topk, double alpha_cluster, double* stats_host, double* tau_host,
                                  int64_t* sel_host, int nboot, double* arim_host, double* arim_std_host,
                                  double* stats_dev_keep, void* stream);

/* The same fig-4/5 sweep on DEVICE buffers: evolution (+ statistics), per-group top-k / Kendall matrices and the
 * ARIM bootstrap issued from one C call, no allocation and no synchronisation inside (what bench.py times as the
 * device-resident step).  fids_dev [S][C][B] is required unless `fused`; stats_dev [15][S][C], tau_dev [G][S][S],
 * sel_dev int64 [G][k], wsel_dev [G][S][k] (k = min(topk, C/G)); arim_dev / arim_std_dev [G][S] optional;
 * counters_dev optional uint64[2] = {non-converged evaluations, illegal samples}, accumulated (zero them first).
 * ev_evolution_begin / ev_evolution_end: optional caller-owned cudaEvent_t recorded on `stream` around the
 * evolution launch (with its finalize kernel when fused), so a benchmark can time the dominant kernel inside its
 * timed region.  workspace: rc_robustness_sweep_workspace_bytes(C, S, B, fused, G, topk). */
size_t rc_robustness_sweep_workspace_bytes(int64_t C, int S, int64_t B, int fused, int64_t G, int64_t topk);
int rc_robustness_sweep(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                        const double* sigma_dev, int S, int64_t B, int model, int zz, uint64_t seed,
                        int64_t c_offset, int64_t b_offset, double dkw_eps, int fused, int64_t G, int64_t topk,
                        double alpha_cluster, double* fids_dev, double* stats_dev, double* tau_dev, int64_t* sel_dev,
                        double* wsel_dev, int nboot, double* arim_dev, double* arim_std_dev,
                        unsigned long long* counters_dev, void* workspace_dev, size_t workspace_bytes,
                        void* ev_evolution_begin, void* ev_evolution_end, void* stream);

/* Low-latency objective evaluation for optimiser loops: ONE controller x_host [N+1] against m explicit
 * perturbations, host buffers in and out, one H2D + one launch + one D2H through cached pinned staging (no
 * allocation in steady state).  rows_host [m][K]: the perturbation of each evaluation in replay layout with
 * sigma = 1 (K = 3N or 2N, reference draw order) — what LBFGS.fidelity_ss(x, use_fixed_ham=True, rH=...) /
 * fidelity_ss_av / wass_cost evaluate (qnewton.py:383-455) and Environment.step's fidelity (RL...py:260);
 * rows_host == NULL with m = 1: the nominal fidelity of x (fidelity_ss(x), qnewton.py:383-400).
 * fids_host [m] and/or stats_host [15]: the statistics of the m fidelities taken as one segment (rc_stats_unsorted
 * order; row 0 = W1 to the ideal distribution = wass_cost's value, 1 - row 0 = their mean = fidelity_ss_av's value).
 * amps_host [m][2] (optional, RC_MODEL_REAL2 only): the complex transfer amplitudes U_k[out,in] = (re, im) — what
 * Environment.state's fixed-Hamiltonian mode averages before squaring (RL...py:147-160: mean propagator, then
 * |<out|mean_k U_k|in>|^2).  Returns RC_ERR_NONCONV after completing if an evaluation did not converge. */
int rc_objective_host(const double* x_host, int nspin, int inspin, int outspin, const double* rows_host, int64_t m,
                      int model, int zz, double dkw_eps, double* fids_host, double* stats_host, double* amps_host,
                      void* stream);

/* m <= 256 evaluations are served by a RESIDENT evaluator: one CTA that stays on the device between calls, polls the
 * calling thread's pinned mailbox for the next request and leaves by itself after RC_OBJECTIVE_IDLE_US (environment,
 * default 1000) microseconds without one — an optimiser loop (scipy.optimize.fmin_l_bfgs_b around LBFGS.fidelity_ss,
 * qnewton.py:497,513) then pays no kernel launch per call.  It runs on a private non-blocking stream (`stream` orders
 * only the one-shot and large-m paths); a device-wide synchronisation issued while it is idle waits for it at most
 * the idle time.  rc_objective_release() asks the calling thread's evaluator on the current device to leave now and
 * waits for it (before timing other device work, or before tearing the context down).  RC_OBJECTIVE_SERVER=0
 * (environment) disables the resident evaluator: every call launches one kernel. */
int rc_objective_release(void);

/* rc_objective_host through a frame the caller fills once per shape and reuses: an optimiser calls its objective
 * thousands of times with the same buffers (the preallocated arrays of engine.ObjectiveEvaluator), and a foreign-function
 * layer pays per argument.  Field meaning = the parameters of rc_objective_host. */
typedef struct rc_objective_frame {
    const double* x_host;
    const double* rows_host;
    double* fids_host;
    double* stats_host;
    double* amps_host;
    void* stream;
    int64_t m;
    double dkw_eps;
    int32_t nspin, inspin, outspin, model, zz, reserved;
} rc_objective_frame;
int rc_objective_call(const rc_objective_frame* frame);

/* Infidelity 1 - |U[out,in]|^2 and its analytic gradient w.r.t. the N biases and the evolution time for C controllers
 * x [C][N+1]: LBFGS.eval_static_fidelity_gradient (qnewton.py:162-212), the L-BFGS inner call (qnewton.py:497,513).
 * Upstream evaluates N + 1 dense matrix exponentials (N of them on 2N x 2N block matrices); here the derivatives come
 * from the eigendecomposition of the tridiagonal Hamiltonian (divided differences of exp(-i lambda T), csrc/rc_grad.cu),
 * one warp per controller, any N <= 32.  rows [C][2N] (optional): explicit perturbation per controller in the real
 * 2-draw replay layout, sigma = 1 (ham_noisy=True adds structured_perturabation(), qnewton.py:179-180).
 * err [C], grad [C][N+1].  NaN controllers give NaN outputs. */
int rc_fidelity_grad(const double* x_dev, int64_t C, int nspin, int inspin, int outspin, const double* rows_dev, int zz,
                     double* err_dev, double* grad_dev, unsigned long long* nonconv_dev, void* stream);
int rc_fidelity_grad_host(const double* x_host, int64_t C, int nspin, int inspin, int outspin, const double* rows_host,
                          int zz, double* err_host, double* grad_host, void* stream);

/* Dense complex matrix exponential of `batch` M x M matrices (M <= 32), interleaved (re, im) float64,
 * row-major: out = expm(A).  The generality path behind the reference's scipy.linalg.expm calls whose
 * argument is not Hermitian tridiagonal: topo="ring" (noise_model.py:83-85), the complex diagonal
 * entries of directional_perturbation (noise_model.py:196-199), arbitrary perturbation() overrides
 * and the 2N x 2N block matrices of the analytic gradient (qnewton.py:186-196).
 * Scaling and squaring with the [13/13] Pade approximant; NaN/Inf input gives NaN output. */
int rc_expm_batch(const double* A_dev, int64_t batch, int M, double* out_dev, void* stream);

/* Batched Monte-Carlo sweep for the ring topology (noise_model.py:83-85, qnewton.py:145-147: unit couplings between
 * sites 0 and N-1 on top of the chain) under the structured perturbation, which the reference evaluates one call at a
 * time with scipy.linalg.expm (noise_model.py:98-109).  Same arguments, Philox counters, replay layout and fids
 * [S][C][B] output as rc_fidelity_mc; ring = 0 evaluates the open chain through the same dense path (cross-check).
 * Each tile of evaluations gets its dense -iTH built on the device, goes through rc_expm_batch and |U[out,in]|^2 is
 * extracted.  workspace: any size >= one matrix pair; rc_dense_fidelity_mc_workspace_bytes(nspin, tile) for a
 * tile of `tile` evaluations per pass. */
size_t rc_dense_fidelity_mc_workspace_bytes(int nspin, int64_t tile);
int rc_dense_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin, const double* sigma_dev,
                         int S, int64_t B, int model, int zz, int ring, uint64_t seed, int64_t c_offset, int64_t b_offset,
                         const double* replay_dev, double* fids_dev, void* workspace_dev, size_t workspace_bytes,
                         void* stream);

/* Batched Monte-Carlo sweep under directional_perturbation (noise_model.py:150-201): per evaluation ONE of the 3N
 * `directions` (noise_model.py:155-163, same order) is drawn and the pair z[i][j] = v, z[j][i] = conj(v),
 * v = sigma (n0 + i n1), is added to H (noise_model.py:191-199; on a diagonal direction H[i][i] becomes complex, so the
 * evaluation goes through the dense exponential like rc_dense_fidelity_mc).  The reference evaluates these one call at
 * a time (noise_model.py:98-109).  replay_dev: [S][C][B][3] = (direction index as a double, n0, n1) with n0, n1 STANDARD
 * normals — exactly what np.random.randint(0, 3N) and rng(size=2) / sigma return upstream — or NULL: the index comes from
 * Philox block sub-stream 127 of the evaluation's (seed, sigma, controller, draw) counter ((word * 3N) >> 32) and the
 * normals are its primary-stream draws 0 and 1; draws_out_dev (optional, Philox mode) receives the [S][C][B][3] draws
 * that were used, so that a Philox sweep can be replayed through the reference / the oracle.  An index outside
 * [0, 3N) gives a NaN fidelity.  fids [S][C][B]; workspace as for rc_dense_fidelity_mc. */
int rc_directional_fidelity_mc(const double* ctrl_dev, int64_t C, int nspin, int inspin, int outspin,
                               const double* sigma_dev, int S, int64_t B, int zz, int ring, uint64_t seed, int64_t c_offset,
                               int64_t b_offset, const double* replay_dev, double* fids_dev, double* draws_out_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream);

/* FP64 FMA throughput micro-benchmark of the current device (TFLOP/s, 2 flops per DFMA); used as
 * the roofline denominator of the evolution kernel. */
int rc_fp64_peak_tflops(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROBCHAR_B200_H */
