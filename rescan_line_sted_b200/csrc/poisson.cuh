// Counter-based Poisson sampler fused into the forward-model epilogue.
//
// Replaces `np.random.poisson(m)` at figure_generation/line_sted_tools.py:510
// (legacy MT19937 stream -> here Philox4x32-10, one independent counter per
// (seed, image, pixel, attempt), so the noise field does not depend on the
// launch geometry).  The samplers follow the same two published algorithms
// numpy's legacy generator uses: Knuth's multiplication method for
// lambda < 10 and Hoermann's PTRS transformed rejection for lambda >= 10
// (W. Hoermann, Insurance: Mathematics and Economics 12 (1993) 39-45), all
// in double precision because figure-2 fluxes reach 1.8e7 counts per pixel
// (> 2^24).  The stream differs from MT19937, so parity with the reference
// is statistical (tests/test_gpu_poisson.py); bit-exact comparisons inject
// the oracle's noise field instead.
#pragma once
#include <stdint.h>
#include <math.h>
#include "fft_core.cuh"

namespace lsted {

struct Philox4 { uint32_t v[4]; };

LSTED_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

LSTED_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                               uint32_t k1) {
    LSTED_UNROLL
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, hi0, lo0);
        mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

// 53-bit uniform in [0, 1)
LSTED_HD double u01(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

// One Poisson(lam) variate.  Counter = (pixel lo, pixel hi, image, attempt).
LSTED_HD double poisson_sample(double lam, unsigned long long seed, unsigned long long pixel,
                               uint32_t image) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t c0 = (uint32_t)pixel, c1 = (uint32_t)(pixel >> 32);
    if (!(lam > 0.0)) return 0.0;  // lam == 0 (or NaN/negative): no events
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        double prod = 1.0, x = 0.0;
        for (uint32_t att = 0;; ++att) {
            const Philox4 r = philox4x32_10(c0, c1, image, att, k0, k1);
            prod *= u01(r.v[0], r.v[1]);
            if (!(prod > enlam)) return x;
            x += 1.0;
            prod *= u01(r.v[2], r.v[3]);
            if (!(prod > enlam)) return x;
            x += 1.0;
        }
    }
    // PTRS.  The squeeze `us >= 0.07 && V <= vr` accepts ~90 % of the attempts, so
    // the logarithms (and invalpha) are evaluated only on the slow path.
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double aa = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t att = 0;; ++att) {
        const Philox4 r = philox4x32_10(c0, c1, image, att, k0, k1);
        const double U = u01(r.v[0], r.v[1]) - 0.5;
        const double V = u01(r.v[2], r.v[3]);
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * aa / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
        if (log(V) + log(invalpha) - log(aa / (us * us) + b) <=
            -lam + k * log(lam) - lgamma(k + 1.0))
            return k;
    }
}

// First PTRS attempt only (lam >= 10): true + the variate when the squeeze accepts it
// (~90 % of the pixels), false when the pixel needs the full sampler.  Splitting the
// two keeps the logarithm / lgamma path out of the common instruction stream: a warp
// would otherwise run it whenever any of its 32 lanes misses the squeeze (97 % of the
// time).  `poisson_sample` on a rejected pixel replays the same counters, so the field
// is identical to calling `poisson_sample` everywhere.
LSTED_HD bool poisson_fast_ptrs(double lam, unsigned long long seed, unsigned long long pixel,
                                uint32_t image, double& out) {
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double aa = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    const Philox4 r = philox4x32_10((uint32_t)pixel, (uint32_t)(pixel >> 32), image, 0u,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const double U = u01(r.v[0], r.v[1]) - 0.5;
    const double V = u01(r.v[2], r.v[3]);
    const double us = 0.5 - fabs(U);
    out = floor((2.0 * aa / us + b) * U + lam + 0.43);
    return us >= 0.07 && V <= vr;
}

}  // namespace lsted
