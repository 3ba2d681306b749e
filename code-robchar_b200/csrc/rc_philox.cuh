// Counter-based noise generation: Philox4x32-10 -> two 53-bit uniforms -> Box-Muller pair.
//
// Replaces the reference's global-state `np.random.normal(scale=sigma)` scalar calls
// (noise_model.py:114-115,135-147).  The counter is a pure function of the GLOBAL
// (sigma index, controller index, draw index, pair index) so results do not depend on how the
// sweep is sharded over GPUs:
//     ctr = { draw_lo32, controller_lo32, sigma_idx | pair_idx << 16, draw_hi16 | controller_hi16 << 16 }
//     key = { seed_lo32, seed_hi32 }
// Pair p yields the standard normals with compact indices 2p and 2p+1, where the compact order is
// the reference draw order with the two discarded site-0 coupling draws removed.
#pragma once
#include <stdint.h>
#include "rc_ql.cuh"

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

RC_HD void normal_pair(uint32_t seed_lo, uint32_t seed_hi, uint32_t sidx, uint64_t cidx, uint64_t bidx,
                       uint32_t pair, double& z0, double& z1) {
    Philox4 c;
    c.x = (uint32_t)bidx;
    c.y = (uint32_t)cidx;
    c.z = (sidx & 0xFFFFu) | (pair << 16);
    c.w = (uint32_t)((bidx >> 32) & 0xFFFFu) | ((uint32_t)((cidx >> 32) & 0xFFFFu) << 16);
    Philox4 r = philox4x32_10(c, seed_lo, seed_hi);
    double u1 = u53(r.x, r.y);
    double u2 = u53(r.z, r.w);
    double rad = rc_sqrt(-2.0 * rc_log01(u1));
    double sn, cs;
    rc_sincos_2pi(u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}

}  // namespace rc
