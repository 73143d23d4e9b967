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

// One attempt `att` of Hoermann's PTRS for lam >= 10: true + the variate when it is accepted.
// FULL = false stops after the squeeze `us >= 0.07 && V <= vr` (~76 % of the attempts at
// large lam: no logarithm); FULL = true goes on to the exact test
//     log(V * invalpha / (a / us^2 + b)) <= -lam + k log(lam) - lgamma(k + 1).
// Both tests are evaluated in forms that cost less than the textbook ones and decide the same
// inequality: the squeeze multiplied through by (b - 2) > 0 (no division for vr), and, for
// k + 1 >= 16, Stirling's series for lgamma with its 0.5 log(k + 1) moved into the left-hand
// logarithm and k log(lam) - k log(k + 1) taken as k log(lam / (k + 1)) -- two logarithms,
// one quotient, a reciprocal and a square root instead of four logarithms, three quotients
// and a library lgamma (series through x^-9: below 2e-16 at x = 16).
template <bool FULL>
LSTED_HD bool ptrs_attempt(double lam, unsigned long long seed, unsigned long long pixel,
                           uint32_t image, uint32_t att, double& out) {
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double aa = -0.059 + 0.02483 * b;
    const Philox4 r = philox4x32_10((uint32_t)pixel, (uint32_t)(pixel >> 32), image, att,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const double U = u01(r.v[0], r.v[1]) - 0.5;
    const double V = u01(r.v[2], r.v[3]);
    const double us = 0.5 - fabs(U);
    const double k = floor((2.0 * aa / us + b) * U + lam + 0.43);
    out = k;
    if (us >= 0.07 && V * (b - 2.0) <= 0.9277 * (b - 2.0) - 3.6224) return true;
    if (!FULL) return false;
    if (k < 0.0 || (us < 0.013 && V > us)) return false;
    // V * invalpha / (a / us^2 + b) with invalpha = 1.1239 + 1.1328 / (b - 3.4), as ONE quotient
    const double us2 = us * us;
    const double num = V * (1.1239 * (b - 3.4) + 1.1328) * us2;
    const double den = (b - 3.4) * (aa + b * us2);
    const double x = k + 1.0;
    if (x >= 16.0) {
        const double xi = 1.0 / x, xi2 = xi * xi;
        const double series = xi * (8.3333333333333329e-02 + xi2 * (-2.7777777777777779e-03 + xi2 *
                              (7.9365079365079365e-04 + xi2 * (-5.9523809523809529e-04 + xi2 *
                              8.4175084175084182e-04))));
        return log(num * sqrt(x) / den) <=
               (x - lam) + k * log(lam * xi) - 0.91893853320467278 - series;
    }
    return log(num / den) <= -lam + k * log(lam) - lgamma(x);
}

// One Poisson(lam) variate.  Counter = (pixel lo, pixel hi, image, attempt).
LSTED_HD double poisson_sample(double lam, unsigned long long seed, unsigned long long pixel,
                               uint32_t image, uint32_t first_attempt = 0u) {
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
    // PTRS: attempts first_attempt, first_attempt + 1, ... until one is accepted (the fast
    // path's dense rounds have already spent the earlier ones)
    for (uint32_t att = first_attempt;; ++att) {
        double k;
        if (ptrs_attempt<true>(lam, seed, pixel, image, att, k)) return k;
    }
}

// First PTRS attempt only (lam >= 10): true + the variate when the squeeze accepts it,
// false when the pixel needs the exact test.  Splitting the two keeps the logarithms out
// of the common instruction stream: a warp would otherwise run them whenever any of its
// 32 lanes misses the squeeze (practically always).  The exact test replays the same
// counters, so the field is identical to calling `poisson_sample` everywhere.
LSTED_HD bool poisson_fast_ptrs(double lam, unsigned long long seed, unsigned long long pixel,
                                uint32_t image, double& out) {
    return ptrs_attempt<false>(lam, seed, pixel, image, 0u, out);
}

}  // namespace lsted
