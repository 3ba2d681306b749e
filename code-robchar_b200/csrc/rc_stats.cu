// Statistics stage: per-(sigma, controller) segment sort + fused RIM / 1-Wasserstein, Q-threshold,
// std and worst-case reductions for the centre and the two DKW-shifted variants.
//
// Reference path replaced: wd_from_ideal (wd_sortof_fast_implementation.py:82-116), the metric
// registry Q / std_fids / wc_fids (mcsim.py:144-183) and the DKW loop of
// MCDataSim.get_metrics_dict (mcsim.py:482-498).
//
//   B <= SMEM_SORT_MAX : one CTA per segment, bitonic sort of order-preserving 64-bit keys in
//                        shared memory, statistics straight from shared memory (one HBM read).
//   larger B           : cub::DeviceSegmentedRadixSort in chunks, then a streaming statistics
//                        kernel over the sorted chunk (second pass served by L2).
#include <stdlib.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "rc_common.cuh"
#include "rc_stats.cuh"

namespace rc {

constexpr int SMEM_SORT_MAX = 4096;  // 32 KB of keys per CTA

__device__ __forceinline__ unsigned long long f2key(double f) {
    unsigned long long b = (unsigned long long)__double_as_longlong(f);
    if (f != f) b = 0x7FF8000000000000ull;  // canonical +NaN sorts last (numpy order)
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key2f(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// Deterministic block-wide sum of NV doubles per thread (fixed shuffle tree + fixed warp order).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch /* [32*NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) scratch[warp * NV + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int w = 0; w < nwarp; ++w) t += scratch[w * NV + k];
        v[k] = t;
    }
}

// Statistics of one ascending-sorted segment, read through `at(i)`.  All threads of the CTA call.
template <class At>
__device__ __forceinline__ void sorted_segment_stats(At at, long long B, double eps, long long seg, long long nseg,
                                                     double* __restrict__ stats, unsigned long long* illegal,
                                                     double* scratch) {
    // pass 1: W sums, plain sums, threshold counts, legality
    double acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0;
    unsigned long long bad = 0;
    for (long long i = threadIdx.x; i < B; i += blockDim.x) {
        const double f = at(i);
        const double fn = (i + 1 < B) ? at(i + 1) : 1.0;
        const double cdf = (double)(i + 1) / (double)B;  // np.arange(1, n+1) / n
        const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
        const double vn[3] = {fn, (i + 1 < B) ? clip01(fn - eps) : 1.0, (i + 1 < B) ? clip01(fn + eps) : 1.0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            acc[k] += (vn[k] - v[k]) * cdf;  // intervals * cdf (wd_sortof_fast_implementation.py:108-114)
            acc[3 + k] += v[k];
            acc[6 + k] += (v[k] >= 0.95) ? 1.0 : 0.0;
            acc[9 + k] += (v[k] >= 0.98) ? 1.0 : 0.0;
        }
        if (fabs(f - 1e-8) > 1.0) ++bad;  // check_fidtype (wd_sortof_fast_implementation.py:23)
    }
    block_sum<12>(acc, scratch);
    if (bad && illegal) atomicAdd(illegal, bad);
    // pass 2: population variance about the mean (np.std)
    double m2[3] = {0.0, 0.0, 0.0};
    const double mean[3] = {acc[3] / (double)B, acc[4] / (double)B, acc[5] / (double)B};
    for (long long i = threadIdx.x; i < B; i += blockDim.x) {
        const double f = at(i);
        const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double dlt = v[k] - mean[k];
            m2[k] += dlt * dlt;
        }
    }
    block_sum<3>(m2, scratch);
    if (threadIdx.x == 0) {
        const double f0 = at(0);
        const double mn[3] = {f0, clip01(f0 - eps), clip01(f0 + eps)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            stats[(0 + k) * nseg + seg] = acc[k];
            stats[(3 + k) * nseg + seg] = -1.0 * (acc[6 + k] / (double)B);  // -Q(thr)  (mcsim.py:144-146,170-176)
            stats[(6 + k) * nseg + seg] = -1.0 * (acc[9 + k] / (double)B);
            stats[(9 + k) * nseg + seg] = sqrt(m2[k] / (double)B);
            stats[(12 + k) * nseg + seg] = -mn[k];                         // -min  (mcsim.py:148)
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One WARP per segment, B <= 32*E: keys live in registers (element g = j*32 + lane), bitonic
// network with register exchanges for distances >= 32 and xor-shuffles below; no shared memory,
// no CTA barriers.  Statistics by warp shuffles.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int E>
__global__ void __launch_bounds__(128) sort_stats_warp_kernel(const double* __restrict__ fids, long long nseg, int B,
                                                              double eps, double* __restrict__ stats, double* sorted_out,
                                                              unsigned long long* illegal) {
    __shared__ unsigned long long srow[4 * 32 * E];
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long seg = warp0; seg < nseg; seg += nwarps) {
        const double* src = fids + seg * (long long)B;
        unsigned long long v[E];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int g = j * 32 + lane;
            v[j] = g < B ? f2key(__ldcs(src + g)) : ~0ull;
        }
        // bitonic sort, ascending in g.  Shuffle stages (distance < 32) run as rolled loops with
        // runtime (k, dist) — one copy in the instruction stream; only the register-exchange stages
        // (distance >= 32, compile-time register indices) are unrolled.
        auto shuffle_stage = [&](int k, int dist) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const unsigned long long a = v[j];
                const unsigned long long b = __shfl_xor_sync(0xffffffffu, a, dist);
                const bool up = (((j * 32 + lane) & k) == 0);
                const bool lower = (lane & dist) == 0;
                v[j] = ((a < b) == (lower == up)) ? a : b;
            }
        };
#pragma unroll 1
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll 1
            for (int dist = k >> 1; dist > 0; dist >>= 1) shuffle_stage(k, dist);
        }
#pragma unroll
        for (int k = 64; k <= 32 * E; k <<= 1) {
#pragma unroll
            for (int dist = k >> 1; dist >= 32; dist >>= 1) {
                const int dj = dist >> 5;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if ((j & dj) == 0) {
                        const bool up = ((j * 32) & k) == 0;  // k >= 64: direction depends on j only
                        const unsigned long long a = v[j], b = v[j | dj];
                        const bool sw = (a > b) == up;
                        v[j] = sw ? b : a;
                        v[j | dj] = sw ? a : b;
                    }
                }
            }
#pragma unroll 1
            for (int dist = 16; dist > 0; dist >>= 1) shuffle_stage(k, dist);
        }
        if (sorted_out) {
            double* dst = sorted_out + seg * (long long)B;
#pragma unroll
            for (int j = 0; j < E; ++j) {
                const int g = j * 32 + lane;
                if (g < B) dst[g] = key2f(v[j]);
            }
        }
        // park the sorted keys in this warp's shared-memory row so the statistics run as rolled loops
        // over elements (one copy of the clip / accumulate code in the instruction stream)
        unsigned long long* row = srow + (size_t)(threadIdx.x >> 5) * (32 * E);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < E; ++j) row[j * 32 + lane] = v[j];
        __syncwarp();
        // statistics (same arithmetic as sorted_segment_stats; threshold counts by ballot + popc)
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        int cnt[6] = {0, 0, 0, 0, 0, 0};
        unsigned bad = 0;
#pragma unroll 1
        for (int g0 = 0; g0 < B; g0 += 32) {
            const int g = g0 + lane;
            const bool in = g < B;
            const bool last = (g + 1 >= B);
            const double f = in ? key2f(row[g]) : 0.0;
            const double fn = (in && !last) ? key2f(row[g + 1]) : 1.0;
            const double cdf = (double)(g + 1) / (double)B;
            const double vv[3] = {f, clip01(f - eps), clip01(f + eps)};
            const double vn[3] = {fn, last ? 1.0 : clip01(fn - eps), last ? 1.0 : clip01(fn + eps)};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (in) {
                    acc[k] += (vn[k] - vv[k]) * cdf;
                    acc[3 + k] += vv[k];
                }
                cnt[k] += __popc(__ballot_sync(0xffffffffu, in && vv[k] >= 0.95));
                cnt[3 + k] += __popc(__ballot_sync(0xffffffffu, in && vv[k] >= 0.98));
            }
            if (in && fabs(f - 1e-8) > 1.0) ++bad;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (bad && illegal && lane == 0) atomicAdd(illegal, (unsigned long long)bad);
        double m2[3] = {0.0, 0.0, 0.0};
        const double mean[3] = {acc[3] / (double)B, acc[4] / (double)B, acc[5] / (double)B};
#pragma unroll 1
        for (int g0 = 0; g0 < B; g0 += 32) {
            const int g = g0 + lane;
            if (g < B) {
                const double f = key2f(row[g]);
                const double vv[3] = {f, clip01(f - eps), clip01(f + eps)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double dlt = vv[k] - mean[k];
                    m2[k] += dlt * dlt;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) m2[k] = warp_sum(m2[k]);
        const double f0 = __shfl_sync(0xffffffffu, key2f(v[0]), 0);
        if (lane < 3) {
            const int k = lane;
            const double mn = k == 0 ? f0 : (k == 1 ? clip01(f0 - eps) : clip01(f0 + eps));
            const double a0 = k == 0 ? acc[0] : (k == 1 ? acc[1] : acc[2]);
            const double a6 = (double)(k == 0 ? cnt[0] : (k == 1 ? cnt[1] : cnt[2]));
            const double a9 = (double)(k == 0 ? cnt[3] : (k == 1 ? cnt[4] : cnt[5]));
            const double mm = k == 0 ? m2[0] : (k == 1 ? m2[1] : m2[2]);
            stats[(0 + k) * nseg + seg] = a0;
            stats[(3 + k) * nseg + seg] = -1.0 * (a6 / (double)B);
            stats[(6 + k) * nseg + seg] = -1.0 * (a9 / (double)B);
            stats[(9 + k) * nseg + seg] = sqrt(mm / (double)B);
            stats[(12 + k) * nseg + seg] = -mn;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sort-free statistics (rc_stats_unsorted): the sweep only needs the VALUES of the 15 metrics, and none of
// them depends on the order of the samples once W is written as mean(1 - v) — the value of wd_from_ideal's
// sorted telescoping sum (wd_sortof_fast_implementation.py:105-114) up to rounding (<= B ulp).  Without the
// bitonic network the per-segment work drops from ~2900 to ~300 warp instructions and the kernel becomes a
// single streaming pass over the fidelity tensor (8 B/sample, HBM/L2 bound).
//   B <= 32*E (E <= 16): one warp per segment, samples in registers, two passes from registers
//   longer segments    : one CTA per segment, second pass re-reads (L2)
// Q = -(#v >= thr)/B (mcsim.py:144-146), std = two-pass population std (np.std, mcsim.py:147), worst case
// = -min (mcsim.py:148), each for v = f, clip(f - eps), clip(f + eps) (mcsim.py:483-485); NaN samples give
// NaN W/std/worst case and count as below threshold, as in numpy.  Fixed reduction trees: deterministic.
// ---------------------------------------------------------------------------------------------
#ifndef RC_STATS_MINB
#define RC_STATS_MINB 3
#endif
template <int E>
__global__ void __launch_bounds__(256, E <= 4 ? RC_STATS_MINB : (E <= 8 ? 3 : 2)) stats_unsorted_warp_kernel(const double* __restrict__ fids, long long nseg, int B,
                                                                  double eps, long long stat_stride,
                                                                  double* __restrict__ stats, unsigned long long* illegal) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const double nB = (double)B;
    // software pipeline: the samples of the warp's NEXT segment are requested before the current one is
    // reduced, so the HBM/L2 round trip hides behind the dependent shuffle/divide chain of the reduction
    double fn[E];
    if (warp0 < nseg) {
#pragma unroll
        for (int j = 0; j < E; ++j) fn[j] = (j * 32 + lane < B) ? __ldcs(fids + warp0 * (long long)B + j * 32 + lane) : 1.0;
    }
    for (long long seg = warp0; seg < nseg; seg += nwarps) {
        double f[E];
#pragma unroll
        for (int j = 0; j < E; ++j) f[j] = fn[j];
        if (seg + nwarps < nseg) {
            const double* nsrc = fids + (seg + nwarps) * (long long)B;
#pragma unroll
            for (int j = 0; j < E; ++j) fn[j] = (j * 32 + lane < B) ? __ldcs(nsrc + j * 32 + lane) : 1.0;
        }
        // NaN samples: flagged here and turned into NaN W / std / worst case at the end, so the hot loop can use
        // the plain min/max clip (fmax(NaN, 0) = 0 keeps a NaN sample out of the threshold counts, as numpy does)
        double sv[3] = {0.0, 0.0, 0.0};
        constexpr bool KEEP = E <= 4;          // keep the clipped variants in registers for the second pass
        double vu[KEEP ? E : 1], vl[KEEP ? E : 1];
        int c95[3] = {0, 0, 0}, c98[3] = {0, 0, 0};
        double mn = INFINITY;
        unsigned flags = 0;   // bit 0: NaN seen; bits 1..: illegal samples
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const double u = fmin(fmax(f[j] - eps, 0.0), 1.0), l = fmin(fmax(f[j] + eps, 0.0), 1.0);
            if (KEEP) { vu[j] = u; vl[j] = l; }
            if (j * 32 + lane < B) {
                const double v[3] = {f[j], u, l};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sv[k] += v[k];
                    c95[k] += v[k] >= 0.95;
                    c98[k] += v[k] >= 0.98;
                }
                mn = fmin(mn, f[j]);
                flags |= (f[j] != f[j]) ? 1u : 0u;
                flags += fabs(f[j] - 1e-8) > 1.0 ? 2u : 0u;   // check_fidtype (wd_sortof_fast_implementation.py:23)
            }
        }
        double mean[3], m2[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 3; ++k) { sv[k] = warp_sum(sv[k]); mean[k] = sv[k] / nB; }
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if (j * 32 + lane < B) {
                const double v[3] = {f[j], KEEP ? vu[j] : fmin(fmax(f[j] - eps, 0.0), 1.0),
                                     KEEP ? vl[j] : fmin(fmax(f[j] + eps, 0.0), 1.0)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double dlt = v[k] - mean[k];
                    m2[k] = fma(dlt, dlt, m2[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            m2[k] = warp_sum(m2[k]);
            c95[k] = __reduce_add_sync(0xffffffffu, c95[k]);
            c98[k] = __reduce_add_sync(0xffffffffu, c98[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        const bool anynan = __reduce_or_sync(0xffffffffu, flags & 1u) != 0;
        const unsigned bad = __reduce_add_sync(0xffffffffu, flags >> 1);
        if (bad && illegal && lane == 0) atomicAdd(illegal, (unsigned long long)bad);
        if (lane < 3) {
            const int k = lane;
            const double mk = k == 0 ? mn : fmin(fmax(mn + (k == 1 ? -eps : eps), 0.0), 1.0);
            const double svk = k == 0 ? sv[0] : (k == 1 ? sv[1] : sv[2]);
            const double a95 = (double)(k == 0 ? c95[0] : (k == 1 ? c95[1] : c95[2]));
            const double a98 = (double)(k == 0 ? c98[0] : (k == 1 ? c98[1] : c98[2]));
            const double mm = k == 0 ? m2[0] : (k == 1 ? m2[1] : m2[2]);
            double* out = stats + seg;
            // W = mean(1 - v) = (B - sum v) / B: the subtraction is exact for sum v in [B/2, 2B] (Sterbenz)
            out[(ST_W + k) * stat_stride] = anynan ? NAN : (nB - svk) / nB;
            out[(ST_Q95 + k) * stat_stride] = -1.0 * (a95 / nB);
            out[(ST_Q98 + k) * stat_stride] = -1.0 * (a98 / nB);
            out[(ST_STD + k) * stat_stride] = anynan ? NAN : sqrt(mm / nB);
            out[(ST_WC + k) * stat_stride] = anynan ? NAN : -mk;
        }
    }
}

// Short segments (B <= 128), round 2: G = 4 lanes per segment instead of a whole warp.  The warp-per-segment kernel
// above spends ~1100 warp-instructions per 100-sample segment (four samples per lane, then seven 5-step shuffle
// reductions, three divisions and a square root executed for ONE segment per warp): 0.29 ms for the 2.1e5 segments of
// the paper sweep, 9 % of the HBM rate.  Here a warp works on 8 segments at once: lane q of a group owns samples
// q, q+4, ... (consecutive lanes read consecutive doubles: every 32-byte sector is used in full), accumulates them in
// ONE pass — sums of y = v - shift and y^2 about the segment's first sample (the fused kernels' form; m2 = syy - sy^2/n),
// the six threshold counts in 8-bit fields of two integers, minimum, NaN / illegal flags — and the group reduces with
// two xor-shuffle steps; lanes 0..2 of the group finish one variant (centre, upper, lower) each.  ~170
// warp-instructions per segment.  EXACT: (E-1)*G < B, only the last element of a lane can be out of range.
template <int G, int E, bool EXACT>
__global__ void __launch_bounds__(256, 2) stats_unsorted_group_kernel(const double* __restrict__ fids, long long nseg, int B,
                                                                   double eps, long long stat_stride,
                                                                   double* __restrict__ stats, unsigned long long* illegal) {
    constexpr int SPW = 32 / G;
    const int lane = threadIdx.x & 31, q = lane & (G - 1), grp = lane / G;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const double nB = (double)B;
    for (long long seg0 = warp0 * SPW; seg0 < nseg; seg0 += nwarps * SPW) {
        const long long seg = seg0 + grp;
        const bool live = seg < nseg;
        const double* src = fids + (live ? seg : seg0) * (long long)B;
        double f[E];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int idx = j * G + q;
            f[j] = ((EXACT && j < E - 1) || idx < B) ? __ldcs(src + idx) : 1.0;
        }
        // shift = the segment's first sample (lane 0 of the group), per variant.  NaN samples are flagged and the segment's
        // W / std / worst case become NaN at the end; they stay out of the threshold counts as in numpy (a NaN compares false)
        const double s0 = __shfl_sync(0xffffffffu, f[0], grp * G);
        // clip to [0, 1] on the bit pattern (6 integer instructions; the compiler turns compare + select on doubles
        // back into its 7-instruction fmin / fmax sequence): negative -> +0, high word >= that of 1.0 -> exactly 1.0
        // (a positive NaN becomes 1, a negative one 0: NaN segments are replaced by NaN at the end anyway)
        auto clip = [](double t) {
            int hi = __double2hiint(t), lo = __double2loint(t);
            const int keep = ~(hi >> 31);
            hi &= keep; lo &= keep;
            const bool ge1 = hi >= 0x3FF00000;
            return __hiloint2double(ge1 ? 0x3FF00000 : hi, ge1 ? 0 : lo);
        };
        const double sh[3] = {s0, clip(s0 - eps), clip(s0 + eps)};
        double sy[3] = {0.0, 0.0, 0.0}, syy[3] = {0.0, 0.0, 0.0};
        unsigned c95 = 0, c98 = 0;       // 8-bit fields (B <= 128): centre, upper, lower
        double mn = INFINITY;
        unsigned flags = 0;              // low half: illegal samples, high half: NaN samples
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if ((EXACT && j < E - 1) || j * G + q < B) {
                const double x = f[j], xm = x - eps, xp = x + eps;
                const double v[3] = {x, clip(xm), clip(xp)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double y = v[k] - sh[k];
                    sy[k] += y;
                    syy[k] = fma(y, y, syy[k]);
                }
                // clip to [0, 1] does not change a comparison with a threshold inside (0, 1)
                if (x >= 0.95) c95 += 1u;
                if (xm >= 0.95) c95 += 0x100u;
                if (xp >= 0.95) c95 += 0x10000u;
                if (x >= 0.98) c98 += 1u;
                if (xm >= 0.98) c98 += 0x100u;
                if (xp >= 0.98) c98 += 0x10000u;
                mn = fmin(mn, x);
                if (x != x) flags += 0x10000u;
                if (fabs(x - 1e-8) > 1.0) flags += 1u;   // check_fidtype (wd_sortof_fast_implementation.py:23)
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                sy[k] += __shfl_xor_sync(0xffffffffu, sy[k], o);
                syy[k] += __shfl_xor_sync(0xffffffffu, syy[k], o);
            }
            c95 += __shfl_xor_sync(0xffffffffu, c95, o);
            c98 += __shfl_xor_sync(0xffffffffu, c98, o);
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            flags += __shfl_xor_sync(0xffffffffu, flags, o);
        }
        const unsigned bad = (live && q == 0) ? (flags & 0xffffu) : 0u;
        const unsigned badw = __reduce_add_sync(0xffffffffu, bad);
        if (badw && illegal && lane == 0) atomicAdd(illegal, (unsigned long long)badw);
        if (live && q < 3) {
            const int k = q;
            const bool anynan = (flags >> 16) != 0 || s0 != s0;
            const double syk = k == 0 ? sy[0] : (k == 1 ? sy[1] : sy[2]);
            const double syyk = k == 0 ? syy[0] : (k == 1 ? syy[1] : syy[2]);
            const double shk = k == 0 ? sh[0] : (k == 1 ? sh[1] : sh[2]);
            const double a95 = (double)((c95 >> (8 * k)) & 0xffu);
            const double a98 = (double)((c98 >> (8 * k)) & 0xffu);
            const double mk = k == 0 ? mn : clip(mn + (k == 1 ? -eps : eps));
            const double svk = fma(nB, shk, syk);                        // sum of the samples
            const double m2 = fmax(syyk - syk * syk / nB, 0.0);
            double* out = stats + seg;
            // W = mean(1 - v) = (B - sum v) / B
            out[(ST_W + k) * stat_stride] = anynan ? NAN : (nB - svk) / nB;
            out[(ST_Q95 + k) * stat_stride] = -1.0 * (a95 / nB);
            out[(ST_Q98 + k) * stat_stride] = -1.0 * (a98 / nB);
            out[(ST_STD + k) * stat_stride] = anynan ? NAN : sqrt(m2 / nB);
            out[(ST_WC + k) * stat_stride] = anynan ? NAN : -mk;
        }
    }
}

// One CTA per (long) segment.
__global__ void __launch_bounds__(256) stats_unsorted_block_kernel(const double* __restrict__ fids, long long nseg, long long B,
                                                                   double eps, long long stat_stride,
                                                                   double* __restrict__ stats, unsigned long long* illegal) {
    __shared__ double scratch[32 * 12];
    __shared__ double smn[8];
    __shared__ unsigned sflags[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const double nB = (double)B;
    for (long long seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const double* src = fids + seg * B;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.0;
        double mn = INFINITY;
        unsigned nan = 0;
        unsigned long long bad = 0;
        for (long long i = threadIdx.x; i < B; i += blockDim.x) {
            const double f = __ldg(src + i);
            const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[k] += 1.0 - v[k];
                acc[3 + k] += v[k];
                acc[6 + k] += (v[k] >= 0.95) ? 1.0 : 0.0;
                acc[9 + k] += (v[k] >= 0.98) ? 1.0 : 0.0;
            }
            mn = fmin(mn, f);
            nan |= (f != f) ? 1u : 0u;
            if (fabs(f - 1e-8) > 1.0) ++bad;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        nan = __reduce_or_sync(0xffffffffu, nan);
        __syncthreads();
        if (lane == 0) { smn[warp] = mn; sflags[warp] = nan; }
        block_sum<12>(acc, scratch);   // contains the barriers that publish smn / sflags
        if (bad && illegal) atomicAdd(illegal, bad);
        double m2[3] = {0.0, 0.0, 0.0};
        const double mean[3] = {acc[3] / nB, acc[4] / nB, acc[5] / nB};
        for (long long i = threadIdx.x; i < B; i += blockDim.x) {
            const double f = __ldg(src + i);
            const double v[3] = {f, clip01(f - eps), clip01(f + eps)};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double dlt = v[k] - mean[k];
                m2[k] = fma(dlt, dlt, m2[k]);
            }
        }
        block_sum<3>(m2, scratch);
        if (threadIdx.x < 3) {
            const int k = threadIdx.x;
            double mnb = smn[0];
            unsigned anynan = sflags[0];
            for (int w = 1; w < nwarp; ++w) { mnb = fmin(mnb, smn[w]); anynan |= sflags[w]; }
            if (anynan) mnb = NAN;
            const double mk = k == 0 ? mnb : clip01(mnb + (k == 1 ? -eps : eps));
            double* out = stats + seg;
            out[(ST_W + k) * stat_stride] = acc[k] / nB;
            out[(ST_Q95 + k) * stat_stride] = -1.0 * (acc[6 + k] / nB);
            out[(ST_Q98 + k) * stat_stride] = -1.0 * (acc[9 + k] / nB);
            out[(ST_STD + k) * stat_stride] = sqrt(m2[k] / nB);
            out[(ST_WC + k) * stat_stride] = -mk;
        }
        __syncthreads();
    }
}

template <int E>
static cudaError_t launch_stats_unsorted_warp(const double* fids, long long nseg, int B, double eps, long long stride,
                                              double* stats, unsigned long long* illegal, int sm, cudaStream_t st) {
    int occ = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stats_unsorted_warp_kernel<E>, 256, 0);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long grid = (long long)sm * occ;
    const long long need = (nseg + 7) / 8;
    if (grid > need) grid = need;
    stats_unsorted_warp_kernel<E><<<(unsigned)grid, 256, 0, st>>>(fids, nseg, B, eps, stride, stats, illegal); rc::note_launch();
    return cudaGetLastError();
}

template <int G, int E, bool EXACT>
static cudaError_t launch_stats_unsorted_group(const double* fids, long long nseg, int B, double eps, long long stride,
                                               double* stats, unsigned long long* illegal, int sm, cudaStream_t st) {
    int occ = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stats_unsorted_group_kernel<G, E, EXACT>, 256, 0);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long grid = (long long)sm * occ;
    const long long need = (nseg + 8 * (32 / G) - 1) / (8 * (32 / G));
    if (grid > need) grid = need;
    stats_unsorted_group_kernel<G, E, EXACT><<<(unsigned)grid, 256, 0, st>>>(fids, nseg, B, eps, stride, stats, illegal); rc::note_launch();
    return cudaGetLastError();
}
template <int E>
static cudaError_t launch_stats_unsorted_quad(const double* fids, long long nseg, int B, double eps, long long stride,
                                              double* stats, unsigned long long* illegal, int sm, cudaStream_t st) {
    return (E - 1) * 4 < B ? launch_stats_unsorted_group<4, E, true>(fids, nseg, B, eps, stride, stats, illegal, sm, st)
                           : launch_stats_unsorted_group<4, E, false>(fids, nseg, B, eps, stride, stats, illegal, sm, st);
}

static bool stats_warp_per_segment() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RC_STATS_WARP"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}

int stats_unsorted_impl(const double* fids_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                        int64_t stat_stride, unsigned long long* illegal_dev, cudaStream_t st) {
    if (nseg < 0 || B < 1) return set_error(RC_ERR_BAD_ARG, "rc_stats_unsorted: nseg=%lld B=%lld", (long long)nseg, (long long)B);
    if (nseg == 0) return RC_OK;
    if (!fids_dev || !stats_dev) return set_error(RC_ERR_NULL, "rc_stats_unsorted: null fids/stats pointer");
    const int sm = device_sm_count();
    cudaError_t err;
    const int b = (int)(B <= 512 ? B : 0);
    if (B > 512) {
        long long grid = nseg < (long long)sm * 8 ? nseg : (long long)sm * 8;
        stats_unsorted_block_kernel<<<(unsigned)grid, 256, 0, st>>>(fids_dev, nseg, B, dkw_eps, stat_stride, stats_dev, illegal_dev); rc::note_launch();
        err = cudaGetLastError();
    } else if (b <= 128 && !stats_warp_per_segment()) {
        // four lanes per segment (RC_STATS_WARP=1 in the environment keeps the warp-per-segment kernels: A/B)
        if (b <= 32) err = launch_stats_unsorted_quad<8>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
        else if (b <= 64) err = launch_stats_unsorted_quad<16>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
        else if (b <= 100) err = launch_stats_unsorted_quad<25>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
        else err = launch_stats_unsorted_quad<32>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    } else if (b <= 32) err = launch_stats_unsorted_warp<1>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    else if (b <= 64) err = launch_stats_unsorted_warp<2>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    else if (b <= 128) err = launch_stats_unsorted_warp<4>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    else if (b <= 256) err = launch_stats_unsorted_warp<8>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    else err = launch_stats_unsorted_warp<16>(fids_dev, nseg, b, dkw_eps, stat_stride, stats_dev, illegal_dev, sm, st);
    RC_CUDA_TRY(err);
    return RC_OK;
}

template <int E>
static cudaError_t launch_sort_stats_warp(const double* fids, long long nseg, int B, double eps, double* stats,
                                          double* sorted_out, unsigned long long* illegal, int sm, cudaStream_t st) {
    int occ = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sort_stats_warp_kernel<E>, 128, 0);
    if (err != cudaSuccess) return err;
    if (occ < 1) occ = 1;
    long long grid = (long long)sm * occ;
    const long long need = (nseg + 3) / 4;
    if (grid > need) grid = need;
    sort_stats_warp_kernel<E><<<(unsigned)grid, 128, 0, st>>>(fids, nseg, B, eps, stats, sorted_out, illegal); rc::note_launch();
    return cudaGetLastError();
}

// One CTA per segment: bitonic sort in shared memory, then statistics.
__global__ void __launch_bounds__(512) sort_stats_small_kernel(const double* __restrict__ fids, long long nseg, int B,
                                                               int P /* pow2 >= B */, double eps,
                                                               double* __restrict__ stats, double* sorted_out,
                                                               unsigned long long* illegal) {
    extern __shared__ unsigned long long keys[];
    __shared__ double scratch[16 * 12];
    for (long long seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const double* src = fids + seg * (long long)B;
        __syncthreads();
        for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = i < B ? f2key(src[i]) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
                    int p = i | j;
                    bool up = (i & k) == 0;
                    unsigned long long a = keys[i], b = keys[p];
                    if ((a > b) == up) { keys[i] = b; keys[p] = a; }
                }
                __syncthreads();
            }
        }
        if (sorted_out) {
            double* dst = sorted_out + seg * (long long)B;
            for (int i = threadIdx.x; i < B; i += blockDim.x) dst[i] = key2f(keys[i]);
        }
        sorted_segment_stats([&](long long i) { return key2f(keys[i]); }, B, eps, seg, nseg, stats, illegal, scratch);
    }
}

// Streaming statistics over already-sorted segments in global memory (large B).
__global__ void __launch_bounds__(512) stats_sorted_kernel(const double* __restrict__ sorted, long long seg0,
                                                           long long nseg_chunk, long long nseg, long long B, double eps,
                                                           double* __restrict__ stats, unsigned long long* illegal) {
    __shared__ double scratch[16 * 12];
    for (long long s = blockIdx.x; s < nseg_chunk; s += gridDim.x) {
        const double* src = sorted + s * B;
        sorted_segment_stats([&](long long i) { return __ldg(src + i); }, B, eps, seg0 + s, nseg, stats, illegal, scratch);
    }
}

struct SegOffset {
    int B;
    __host__ __device__ int operator()(int i) const { return i * B; }
};

static long long large_chunk_segments(long long nseg, long long B) {
    const long long max_items = 1ll << 28;  // 2 GiB of doubles per chunk, well inside int offsets
    long long c = max_items / B;
    if (c < 1) c = 1;
    if (c > nseg) c = nseg;
    return c;
}

static size_t cub_temp_bytes(long long chunk_segs, long long B) {
    size_t bytes = 0;
    auto off = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), SegOffset{(int)B});
    cub::DeviceSegmentedRadixSort::SortKeys(nullptr, bytes, (const double*)nullptr, (double*)nullptr,
                                            (int)(chunk_segs * B), (int)chunk_segs, off, off + 1);
    return (bytes + 255) & ~(size_t)255;
}

}  // namespace rc

using namespace rc;

extern "C" int rc_stats_unsorted(const double* fids_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                                 unsigned long long* illegal_dev, void* stream) {
    return stats_unsorted_impl(fids_dev, nseg, B, dkw_eps, stats_dev, nseg, illegal_dev, (cudaStream_t)stream);
}

extern "C" size_t rc_stats_workspace_bytes(int64_t nseg, int64_t B) {
    if (nseg <= 0 || B <= SMEM_SORT_MAX) return 256;
    if (B >= (1ll << 30)) return 0;
    long long cs = large_chunk_segments(nseg, B);
    return (size_t)cs * B * sizeof(double) + cub_temp_bytes(cs, B) + 256;
}

extern "C" int rc_stats(const double* fids_dev, int64_t nseg, int64_t B, double dkw_eps, double* stats_dev,
                        double* sorted_dev, unsigned long long* illegal_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream) {
    if (nseg < 0 || B < 1) return set_error(RC_ERR_BAD_ARG, "rc_stats: nseg=%lld B=%lld", (long long)nseg, (long long)B);
    if (nseg == 0) return RC_OK;
    if (!fids_dev || !stats_dev) return set_error(RC_ERR_NULL, "rc_stats: null fids/stats pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int sm = device_sm_count();
    if (B <= 512 && !getenv("RC_STATS_NO_WARP")) {
        cudaError_t err;
        const int b = (int)B;
        if (b <= 32) err = launch_sort_stats_warp<1>(fids_dev, nseg, b, dkw_eps, stats_dev, sorted_dev, illegal_dev, sm, st);
        else if (b <= 64) err = launch_sort_stats_warp<2>(fids_dev, nseg, b, dkw_eps, stats_dev, sorted_dev, illegal_dev, sm, st);
        else if (b <= 128) err = launch_sort_stats_warp<4>(fids_dev, nseg, b, dkw_eps, stats_dev, sorted_dev, illegal_dev, sm, st);
        else if (b <= 256) err = launch_sort_stats_warp<8>(fids_dev, nseg, b, dkw_eps, stats_dev, sorted_dev, illegal_dev, sm, st);
        else err = launch_sort_stats_warp<16>(fids_dev, nseg, b, dkw_eps, stats_dev, sorted_dev, illegal_dev, sm, st);
        RC_CUDA_TRY(err);
        return RC_OK;
    }
    if (B <= SMEM_SORT_MAX) {
        int P = 1;
        while (P < B) P <<= 1;
        if (P < 2) P = 2;
        // one warp per small segment (CTA barrier == warp barrier), more warps only for long segments
        int threads = P / 8;
        if (const char* e = getenv("RC_STATS_DIV")) threads = P / atoi(e);
        if (threads < 32) threads = 32;
        if (threads > 512) threads = 512;
        size_t smem = (size_t)P * sizeof(unsigned long long);
        int occ = 0;
        RC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sort_stats_small_kernel, threads, smem));
        if (occ < 1) occ = 1;
        long long grid = (long long)sm * occ;
        if (grid > nseg) grid = nseg;
        sort_stats_small_kernel<<<(unsigned)grid, threads, smem, st>>>(fids_dev, nseg, (int)B, P, dkw_eps, stats_dev,
                                                                        sorted_dev, illegal_dev); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        return RC_OK;
    }
    if (B >= (1ll << 30)) return set_error(RC_ERR_BAD_ARG, "rc_stats: B=%lld too large (limit 2^30)", (long long)B);
    const size_t need = rc_stats_workspace_bytes(nseg, B);
    if (!workspace_dev || workspace_bytes < need)
        return set_error(RC_ERR_WORKSPACE, "rc_stats: workspace %zu < required %zu bytes", workspace_bytes, need);
    const long long cs = large_chunk_segments(nseg, B);
    double* chunk = (double*)workspace_dev;
    size_t tmp_bytes = cub_temp_bytes(cs, B);
    void* tmp = (char*)workspace_dev + (((size_t)cs * B * sizeof(double) + 255) & ~(size_t)255);
    for (long long s0 = 0; s0 < nseg; s0 += cs) {
        long long ns = nseg - s0 < cs ? nseg - s0 : cs;
        auto off = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), SegOffset{(int)B});
        size_t tb = tmp_bytes;
        RC_CUDA_TRY(cub::DeviceSegmentedRadixSort::SortKeys(tmp, tb, fids_dev + s0 * B, chunk, (int)(ns * B), (int)ns,
                                                            off, off + 1, 0, 64, st));
        long long grid = ns < (long long)sm * 4 ? ns : (long long)sm * 4;
        stats_sorted_kernel<<<(unsigned)grid, 512, 0, st>>>(chunk, s0, ns, nseg, B, dkw_eps, stats_dev, illegal_dev); rc::note_launch();
        RC_CUDA_TRY(cudaGetLastError());
        if (sorted_dev)
            RC_CUDA_TRY(cudaMemcpyAsync(sorted_dev + s0 * B, chunk, (size_t)ns * B * sizeof(double),
                                        cudaMemcpyDeviceToDevice, st));
    }
    return RC_OK;
}


// ---------------------------------------------------------------------------------------------
// p-RIM: (mean((1 - f)^p))^(1/p) per segment (RIM_p, wd_sortof_fast_implementation.py:147-174; p = 0 gives 1,
// p = 1 is the W row of the statistics).  One warp per segment, one streaming pass, fixed reduction tree.
// ---------------------------------------------------------------------------------------------
namespace rc {
__global__ void __launch_bounds__(256) rim_p_kernel(const double* __restrict__ fids, long long nseg, long long B, double p,
                                                    double* __restrict__ out, unsigned long long* illegal) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long seg = warp0; seg < nseg; seg += nwarps) {
        const double* src = fids + seg * B;
        double acc = 0.0;
        unsigned bad = 0;
        for (long long j = lane; j < B; j += 32) {
            const double f = __ldcs(src + j);
            bad += fabs(f - 1e-8) > 1.0 ? 1u : 0u;          // check_fidtype (wd_sortof_fast_implementation.py:23)
            acc += pow(1.0 - f, p);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) {
            out[seg] = p == 0.0 ? 1.0 : pow(acc / (double)B, 1.0 / p);
            if (bad && illegal) atomicAdd(illegal, (unsigned long long)bad);
        }
    }
}
}  // namespace rc

extern "C" int rc_rim_p(const double* fids_dev, int64_t nseg, int64_t B, double p, double* out_dev,
                        unsigned long long* illegal_dev, void* stream) {
    if (nseg < 0 || B < 1) return rc::set_error(RC_ERR_BAD_ARG, "rc_rim_p: nseg=%lld B=%lld", (long long)nseg, (long long)B);
    if (nseg == 0) return RC_OK;
    if (!fids_dev || !out_dev) return rc::set_error(RC_ERR_NULL, "rc_rim_p: null pointer");
    long long blocks = (nseg + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rc::rim_p_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(fids_dev, nseg, B, p, out_dev, illegal_dev); rc::note_launch();
    RC_CUDA_TRY(cudaGetLastError());
    return RC_OK;
}
