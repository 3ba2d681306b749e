// Counter-based noise generation: Philox4x32-10 -> 64 random bits per draw -> 1024-layer ziggurat.
//
// Replaces the reference's global-state `np.random.normal(scale=sigma)` scalar calls
// (noise_model.py:114-115,135-147).  The counter is a pure function of the GLOBAL
// (sigma index, controller index, draw index, sub-stream) so results do not depend on how the
// sweep is sharded over GPUs:
//     ctr = { draw_lo32, controller_lo32, sigma_idx | sub << 16, draw_hi16 | controller_hi16 << 16 }
//     key = { seed_lo32, seed_hi32 }
//     sub = p                      primary stream: Philox block p yields the draws with compact indices 2p, 2p+1
//     sub = jc | attempt << 7      completion stream of draw jc (attempt >= 1), only used when the
//                                  ziggurat fast path misses (0.43 % of draws)
// The compact order is the reference draw order with the two discarded site-0 coupling draws removed.
//
// Ziggurat (Marsaglia & Tsang construction, tables from tools/gen_zig_table.py): 64 bits per draw =
// 10-bit layer index | sign | 53-bit magnitude (disjoint bit fields).  Fast path: x = mag * W[i] is
// accepted when mag < K[i] — one table load, one integer compare, one conversion, one multiply.
// Misses are NOT resolved in line (a warp would pay the wedge code whenever any of its 32 lanes
// missed): the raw bits are parked in the draw's slot, its index is pushed on a small per-lane list,
// and the list is drained after the block loop.  Completion of a missed draw: wedge test against the
// density (one log), tail sampling for layer 0 (Marsaglia's exponential rejection), and when the wedge
// test rejects, a fresh Box-Muller normal — a rejection sampler's restart may use any exact N(0,1)
// generator, since first-attempt acceptances are already exactly normal.
#pragma once
#ifndef RC_ZIG_UNROLL2
#define RC_ZIG_UNROLL2 0
#endif
#include <stdint.h>
#include "rc_ql.cuh"
#include "rc_zig_const.h"

namespace rc {

struct Philox4 { uint32_t x, y, z, w; };

RC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

RC_HD Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
        Philox4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// (0,1) uniform from 53 random bits
RC_HD double u53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return fma((double)v, 0x1.0p-53, 0x1.0p-54);
}

struct NoiseKey {
    uint32_t seed_lo, seed_hi, sidx;
    uint64_t cidx, bidx;   // GLOBAL controller / draw indices
};

RC_HD Philox4 philox_block(const NoiseKey& k, uint32_t sub) {
    Philox4 c;
    c.x = (uint32_t)k.bidx;
    c.y = (uint32_t)k.cidx;
    c.z = (k.sidx & 0xFFFFu) | (sub << 16);
    c.w = (uint32_t)((k.bidx >> 32) & 0xFFFFu) | ((uint32_t)((k.cidx >> 32) & 0xFFFFu) << 16);
    return philox4x32_10(c, k.seed_lo, k.seed_hi);
}

// ------------------------------------------------------------------------------------------------
// Ziggurat
// ------------------------------------------------------------------------------------------------
constexpr int ZIG_BITS = 10;
constexpr int ZIG_LAYERS = 1 << ZIG_BITS;
struct
#if defined(__CUDACC__)
    __align__(16)
#endif
    ZigEntry { unsigned long long k; double w; };   // fast-accept threshold, X[i] * 2^-53
struct ZigTables { const ZigEntry* kw; const double* y; };   // kw may point to shared memory; y[1025] = f(X[i])

RC_HD double bits_as_double(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
RC_HD uint64_t double_as_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
RC_HD double flip_sign_if(double x, uint32_t lo) {   // sign = bit ZIG_BITS of the low word
    return bits_as_double(double_as_bits(x) ^ ((uint64_t)((lo << (31 - ZIG_BITS)) & 0x80000000u) << 32));
}

// One draw, fast path.  Returns the value to store in the draw's slot: the normal when accepted,
// the raw random bits otherwise (then *miss = true).
RC_HD double zig_try(uint32_t hi, uint32_t lo, const ZigEntry* __restrict__ kw, bool* miss) {
    const uint64_t bits = ((uint64_t)hi << 32) | lo;
    const ZigEntry en = kw[lo & (ZIG_LAYERS - 1)];
    const uint64_t mag = bits >> 11;
    const double x = flip_sign_if((double)(long long)mag * en.w, lo);
    *miss = !(mag < en.k);
    return *miss ? bits_as_double(bits) : x;
}

// Completion of a missed draw (see the header comment).  `parked` = the raw bits zig_try returned.
RC_HD double zig_complete(const NoiseKey& key, uint32_t jc, double parked, const ZigTables& t) {
    const uint64_t bits = double_as_bits(parked);
    const uint32_t lo = (uint32_t)bits, idx = lo & (ZIG_LAYERS - 1);
    const double x = (double)(long long)(bits >> 11) * t.kw[idx].w;
    int mode = idx == 0 ? 1 : 0;                 // 0 wedge, 1 tail, 2 fresh normal
    double ylo = 0.0, ydif = 1.0;
    if (mode == 0) { ylo = t.y[idx]; ydif = t.y[idx + 1] - ylo; }
    double val = 0.0;
    uint32_t attempt = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    while (true) {
        const Philox4 r = philox_block(key, jc | (attempt << 7));
        attempt = attempt < 500 ? attempt + 1 : 1;   // 9-bit field; a wrap needs ~500 consecutive tail rejections
        const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
        const double la = rc_log01(mode == 0 ? fma(u1, ydif, ylo) : u1);
        if (mode == 0) {
            if (la < -0.5 * x * x) { val = flip_sign_if(x, lo); break; }
            mode = 2;
        } else if (mode == 1) {
            const double xt = -la * RC_ZIG_RINV;
            if (-2.0 * rc_log01(u2) > xt * xt) { val = flip_sign_if(RC_ZIG_R + xt, lo); break; }
        } else {
            double sn, cs;
            rc_sincos_2pi(u2, &sn, &cs);
            val = rc_sqrt(-2.0 * la) * cs;
            break;
        }
    }
    return val;
}

// All `nc` standard normals of one evaluation.  slot(jc) returns a reference to the storage of the
// draw with compact index jc (any addressable memory: shared, global).  The block loop handles two
// Philox blocks per trip (two independent integer chains and four table lookups in flight — the
// fast path is one long dependency chain otherwise) and is deliberately not unrolled further: code
// size matters, see rc_ql.cuh.
template <class Slot>
RC_HD void normals_fill(const NoiseKey& key, int nc, const ZigTables& t, Slot&& slot) {
#if RC_ZIG_UNROLL2
    const int np = (nc + 1) / 2;
    uint32_t pend = 0;      // up to four parked draw indices (jc + 1), 8 bits each
    int p = 0;
    while (true) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (; p < np; p += 2) {
            if (pend) break;         // a trip can park up to four draws: start it with an empty list (rare)
            const Philox4 ra = philox_block(key, (uint32_t)p);
            const Philox4 rb = philox_block(key, (uint32_t)p + 1u);
            bool m0, m1, m2, m3;
            const int j0 = 2 * p;
            const double v0 = zig_try(ra.x, ra.y, t.kw, &m0);
            const double v1 = zig_try(ra.z, ra.w, t.kw, &m1);
            const double v2 = zig_try(rb.x, rb.y, t.kw, &m2);
            const double v3 = zig_try(rb.z, rb.w, t.kw, &m3);
            slot(j0) = v0;
            if (m0) pend = (uint32_t)(j0 + 1);
            if (j0 + 1 < nc) { slot(j0 + 1) = v1; if (m1) pend = (pend << 8) | (uint32_t)(j0 + 2); }
            if (j0 + 2 < nc) { slot(j0 + 2) = v2; if (m2) pend = (pend << 8) | (uint32_t)(j0 + 3); }
            if (j0 + 3 < nc) { slot(j0 + 3) = v3; if (m3) pend = (pend << 8) | (uint32_t)(j0 + 4); }
        }
        if (!pend && p >= np) break;
        while (pend) {
            const uint32_t jc = (pend & 0xFFu) - 1u;
            pend >>= 8;
            slot((int)jc) = zig_complete(key, jc, slot((int)jc), t);
        }
        if (p >= np) break;
    }
#else
    const int np = (nc + 1) / 2;
    uint32_t pend = 0;      // up to four parked draw indices (jc + 1), 8 bits each
    int p = 0;
    // structured loops only (conditions instead of `break`): the warp reconverges after the drain of every trip
    do {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (; p < np && !(pend >> 16); ++p) {   // fewer than two free entries: drain first (rare)
            const Philox4 r = philox_block(key, (uint32_t)p);
            bool miss;
            const int j0 = 2 * p;
            slot(j0) = zig_try(r.x, r.y, t.kw, &miss);
            if (miss) pend = (pend << 8) | (uint32_t)(j0 + 1);
            if (j0 + 1 < nc) {
                slot(j0 + 1) = zig_try(r.z, r.w, t.kw, &miss);
                if (miss) pend = (pend << 8) | (uint32_t)(j0 + 2);
            }
        }
        while (pend) {
            const uint32_t jc = (pend & 0xFFu) - 1u;
            pend >>= 8;
            slot((int)jc) = zig_complete(key, jc, slot((int)jc), t);
        }
    } while (p < np);
#endif
}

}  // namespace rc
