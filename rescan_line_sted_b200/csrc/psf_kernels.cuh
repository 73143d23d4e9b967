// PSF synthesis kernels (fp64): illumination patterns, pulse saturation and
// the rescan / descan system PSFs.  One CTA per operating point of a sweep.
//
// Replaces generate_psfs(), figure_generation/line_sted_tools.py:168-363.
//
// Structure exploited (verified against the reference loop by the oracle
// tests): a Gaussian-filtered delta is an outer product of 1-D profiles.
// With taps w[0..2r] (scipy `_gaussian_kernel1d`, mode='reflect'), c = n//2:
//   k1[i] = w[r + c - i]                      blur of the 1-D delta   (:185,:191,:200,:211,:260)
//   g2[i] = sum_t w[t] k1[reflect(i+t-r)]     second blur             (:202,:213)
//   point: inner = k1 (x) k1, outer = g2 (x) g2;  line: rows = k1, g2
//   depletion = outer/max - inner/max, both scaled to peak brightness (:204-207,:215-218)
//   fractions 1 - 2^-exc, 2^-dep, product                            (:223-243)
// Rescan (:258-310), with w_rev[s] = sted_row[n-1-s], E = k1 (x) k1, W = ratio*n:
//   descan[y][s] = w_rev[s] * k1[y] * sum_x k1[x]
//   rescan[y][j] = k1[y] * sum_s w_rev[s] * sum_{r<ratio, x<n} k1[(x - c + s) mod n],
//                  x = ((j - s)*ratio + r - ratio//2 + c) mod W
#pragma once
#include <math.h>
#include "fft_core.cuh"

namespace lsted {

enum { kPsfThreads = 256, kPsfMaxN = 1024, kPsfMaxTaps = 2048 };

struct PsfIlluminationArgs {
    int psf_type;  // 0 point, 1 line
    int n, radius;
    const double* taps;
    const double* exc_brightness;  // [batch]
    const double* dep_brightness;  // [batch]
    double* out;                   // [batch][5][n][n]
};

struct PsfRescanArgs {
    int n, radius;
    const double* taps;
    const double* sted_rows;  // [batch][n]
    const int* ratios;        // [batch]
    double* out;              // [batch][3][n][n]: emission, rescan, descan
    double* wide;             // [n][ratio*n] or null (batch 1)
};

struct PsfSmem {
    double k1[kPsfMaxN];
    double g2[kPsfMaxN];   // illumination: second blur; rescan: q[j]
    double red[kPsfThreads];
    double sc[4];
};

LSTED_HD int reflect_idx(int i, int n) {
    const int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - 1 - m;
}

template <class Ctx> LSTED_HD void psf_profiles(Ctx& cx, int n, int radius, const double* taps,
                                                PsfSmem* sm, bool second) {
    const int c = n / 2;
    cx.parallel_for(n, [&](int i) {
        const int t = radius + c - i;
        sm->k1[i] = (t >= 0 && t <= 2 * radius) ? taps[t] : 0.0;
    });
    if (!second) return;
    cx.parallel_for(n, [&](int i) {
        double s = 0;
        for (int t = 0; t <= 2 * radius; ++t) s += taps[t] * sm->k1[reflect_idx(i + t - radius, n)];
        sm->g2[i] = s;
    });
}

template <class Ctx>
LSTED_HD void psf_illumination_body(Ctx& cx, int b, const PsfIlluminationArgs& a, PsfSmem* sm) {
    const int n = a.n;
    const size_t img = (size_t)n * n;
    psf_profiles(cx, n, a.radius, a.taps, sm, true);
    cx.parallel_for(1, [&](int) {
        double m1 = sm->k1[0], m2 = sm->g2[0];
        for (int i = 1; i < n; ++i) {
            m1 = sm->k1[i] > m1 ? sm->k1[i] : m1;
            m2 = sm->g2[i] > m2 ? sm->g2[i] : m2;
        }
        sm->sc[0] = a.psf_type == 0 ? m1 * m1 : m1;  // inner.max()
        sm->sc[1] = a.psf_type == 0 ? m2 * m2 : m2;  // outer.max()
    });
    const double inner_max = sm->sc[0], outer_max = sm->sc[1];
    const bool point = a.psf_type == 0;
    // max of the un-scaled difference of Gaussians
    const size_t count = point ? img : (size_t)n;
    cx.parallel_for(kPsfThreads, [&](int lane) {
        double m = -1e300;
        for (size_t p = lane; p < count; p += kPsfThreads) {
            const int y = (int)(p / n), x = (int)(p - (size_t)y * n);
            const double inner = point ? sm->k1[y] * sm->k1[x] : sm->k1[x];
            const double outer = point ? sm->g2[y] * sm->g2[x] : sm->g2[x];
            const double d = outer / outer_max - inner / inner_max;
            m = d > m ? d : m;
        }
        sm->red[lane] = m;
    });
    cx.parallel_for(1, [&](int) {
        double m = sm->red[0];
        for (int i = 1; i < kPsfThreads; ++i) m = sm->red[i] > m ? sm->red[i] : m;
        sm->sc[2] = m;
    });
    const double exc_scale = a.exc_brightness[b] / inner_max;
    const double dep_scale = a.dep_brightness[b] / sm->sc[2];
    double* out = a.out + img * 5 * (size_t)b;
    cx.parallel_for((int)img, [&](int p) {
        const int y = p / n, x = p - y * n;
        const double inner = point ? sm->k1[y] * sm->k1[x] : sm->k1[x];
        const double outer = point ? sm->g2[y] * sm->g2[x] : sm->g2[x];
        const double exc = inner * exc_scale;
        const double dep = (outer / outer_max - inner / inner_max) * dep_scale;
        const double ef = 1.0 - exp2(-exc);
        const double df = exp2(-dep);
        out[p] = exc;
        out[img + p] = dep;
        out[2 * img + p] = ef;
        out[3 * img + p] = df;
        out[4 * img + p] = ef * df;
    });
}

// sum of k1[(x - c + s) mod n] over x in [x0, x1)
LSTED_HD double cyclic_run_sum(const double* k1, int n, int x0, int x1, int shift) {
    double acc = 0;
    int src = (x0 + shift) % n;
    if (src < 0) src += n;
    for (int x = x0; x < x1; ++x) {
        acc += k1[src];
        if (++src == n) src = 0;
    }
    return acc;
}

template <class Ctx>
LSTED_HD void psf_rescan_body(Ctx& cx, int b, const PsfRescanArgs& a, PsfSmem* sm) {
    const int n = a.n, c = n / 2;
    const int R = a.ratios[b];
    const int W = R * n;
    const size_t img = (size_t)n * n;
    const double* row = a.sted_rows + (size_t)n * b;
    psf_profiles(cx, n, a.radius, a.taps, sm, false);
    cx.parallel_for(1, [&](int) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += sm->k1[i];
        sm->sc[0] = s;
    });
    cx.parallel_for(n, [&](int j) {
        double q = 0;
        for (int s = 0; s < n; ++s) {
            int base = ((j - s) * R + c - R / 2) % W;
            if (base < 0) base += W;
            double acc = 0;
            if (base < n) {  // x = base + r
                const int r1 = (n - base) < R ? (n - base) : R;
                acc += cyclic_run_sum(sm->k1, n, base, base + r1, s - c);
            }
            if (W - base < R) {  // wrapped: x = base + r - W
                const int r0 = W - base;
                const int r1 = (r0 + n) < R ? (r0 + n) : R;
                acc += cyclic_run_sum(sm->k1, n, 0, r1 - r0, s - c);
            }
            q += row[n - 1 - s] * acc;
        }
        sm->g2[j] = q;
    });
    const double ksum = sm->sc[0];
    double* out = a.out + img * 3 * (size_t)b;
    cx.parallel_for((int)img, [&](int p) {
        const int y = p / n, x = p - y * n;
        out[p] = sm->k1[y] * sm->k1[x];
        out[img + p] = sm->k1[y] * sm->g2[x];
        out[2 * img + p] = row[n - 1 - x] * (sm->k1[y] * ksum);
    });
    if (a.wide) {
        // u[X] = sum_s [x<n] w_rev[s] k1[(x - c + s) mod n], x = (X - s*R + c) mod W,
        // parked in the last row, then wide[y][X] = k1[y] * u[X].
        double* u = a.wide + (size_t)(n - 1) * W;
        cx.parallel_for(W, [&](int X) {
            double acc = 0;
            for (int s = 0; s < n; ++s) {
                int x = (X - s * R + c) % W;
                if (x < 0) x += W;
                if (x < n) {
                    int src = (x - c + s) % n;
                    if (src < 0) src += n;
                    acc += row[n - 1 - s] * sm->k1[src];
                }
            }
            u[X] = acc;
        });
        cx.parallel_for(W, [&](int X) {
            const double ux = u[X];
            for (int y = 0; y < n; ++y) a.wide[(size_t)y * W + X] = sm->k1[y] * ux;
        });
    }
}

// ---------------------------------------------------------------------------
// Orientation step: the caller of Deconvolver rotates one
// system PSF to K line orientations with scipy.ndimage.rotate(order=3,
// mode='constant', reshape=False) and clips to [0, 1.1*max]
// (line_sted_figure_2.py:264-272, used at :244-247).  One CTA per orientation:
//   1. cubic B-spline prefilter of the plane, axis 0 then axis 1 (gain 6, pole
//      sqrt(3)-2, mirror boundary with the exact whole-line causal start),
//   2. four-tap interpolation at matrix*(i,j)+offset, 0 outside [0, n-1],
//      coefficient indices mirrored at the edges, 3. clip.
// ---------------------------------------------------------------------------
struct PsfRotateArgs {
    int n0, n1;            // plane shape
    const double* plane;   // [n0][n1], shared by all orientations
    const double* xform;   // [batch][6]: m00 m01 m10 m11 offset0 offset1
    double clip_hi;
    double* coef;          // [batch][n0][n1] scratch (spline coefficients)
    double* out;           // [batch][n0][n1]
};

// in-place prefilter of one line c[0], c[stride], ... (n elements)
LSTED_HD void spline_prefilter_line(double* c, int n, size_t stride) {
    if (n < 2) return;
    const double z = sqrt(3.0) - 2.0;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    for (int i = 0; i < n; ++i) c[i * stride] *= gain;
    double z_n_1 = 1.0;
    for (int i = 0; i < n - 1; ++i) z_n_1 *= z;
    double s = c[0] + z_n_1 * c[(n - 1) * stride];
    double z_i = z;
    for (int i = 1; i < n - 1; ++i) {
        s += z_i * (c[i * stride] + z_n_1 * c[(n - 1 - i) * stride]);
        z_i *= z;
    }
    c[0] = s / (1.0 - z_n_1 * z_n_1);
    for (int i = 1; i < n; ++i) c[i * stride] += z * c[(i - 1) * stride];
    c[(n - 1) * stride] = (z * c[(n - 2) * stride] + c[(n - 1) * stride]) * z / (z * z - 1.0);
    for (int i = n - 2; i >= 0; --i) c[i * stride] = z * (c[(i + 1) * stride] - c[i * stride]);
}
LSTED_HD int spline_mirror(int idx, int n) {
    if (n <= 1) return 0;
    const int s2 = 2 * n - 2;
    if (idx < 0) {
        idx = s2 * (-idx / s2) + idx;
        idx = idx <= 1 - n ? idx + s2 : -idx;
    } else if (idx >= n) {
        idx -= s2 * (idx / s2);
        if (idx >= n) idx = s2 - idx;
    }
    return idx;
}
LSTED_HD void cubic_weights(double x, double* w) {
    const double y = x - floor(x), z = 1.0 - y;
    w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
    w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
    w[0] = z * z * z / 6.0;
    w[3] = 1.0 - w[0] - w[1] - w[2];
}

template <class Ctx> LSTED_HD void psf_rotate_body(Ctx& cx, int b, const PsfRotateArgs& a) {
    const int n0 = a.n0, n1 = a.n1;
    const size_t img = (size_t)n0 * n1;
    double* co = a.coef + img * b;
    double* out = a.out + img * b;
    const double* m = a.xform + 6 * (size_t)b;
    cx.parallel_for(n1, [&](int j) {          // axis 0: one column per thread
        for (int i = 0; i < n0; ++i) co[(size_t)i * n1 + j] = a.plane[(size_t)i * n1 + j];
        spline_prefilter_line(co + j, n0, (size_t)n1);
    });
    cx.parallel_for(n0, [&](int i) {          // axis 1: one row per thread
        spline_prefilter_line(co + (size_t)i * n1, n1, 1);
    });
    cx.parallel_for((int)img, [&](int e) {
        const int i = e / n1, j = e - i * n1;
        const double y = m[0] * i + m[1] * j + m[4];
        const double x = m[2] * i + m[3] * j + m[5];
        double v = 0.0;
        if (!(y < 0.0 || y > (double)(n0 - 1) || x < 0.0 || x > (double)(n1 - 1))) {
            double wy[4], wx[4];
            cubic_weights(y, wy);
            cubic_weights(x, wx);
            const int sy = (int)floor(y) - 1, sx = (int)floor(x) - 1;
            int xx[4];
            for (int q = 0; q < 4; ++q) xx[q] = spline_mirror(sx + q, n1);
            for (int p = 0; p < 4; ++p) {
                const double* row = co + (size_t)spline_mirror(sy + p, n0) * n1;
                for (int q = 0; q < 4; ++q) v += wy[p] * wx[q] * row[xx[q]];
            }
        }
        out[e] = v < 0.0 ? 0.0 : (v > a.clip_hi ? a.clip_hi : v);
    });
}

#ifdef __CUDACC__
struct PsfDeviceCtx {
    template <class F> __device__ __forceinline__ void parallel_for(int n, F f) {
        for (int w = threadIdx.x; w < n; w += blockDim.x) f(w);
        __syncthreads();
    }
};
__global__ void __launch_bounds__(kPsfThreads) psf_illumination_kernel(PsfIlluminationArgs a) {
    __shared__ PsfSmem sm;
    PsfDeviceCtx cx;
    psf_illumination_body(cx, blockIdx.x, a, &sm);
}
__global__ void __launch_bounds__(kPsfThreads) psf_rescan_kernel(PsfRescanArgs a) {
    __shared__ PsfSmem sm;
    PsfDeviceCtx cx;
    psf_rescan_body(cx, blockIdx.x, a, &sm);
}
__global__ void __launch_bounds__(kPsfThreads) psf_rotate_kernel(PsfRotateArgs a) {
    PsfDeviceCtx cx;
    psf_rotate_body(cx, blockIdx.x, a);
}
#endif

}  // namespace lsted
