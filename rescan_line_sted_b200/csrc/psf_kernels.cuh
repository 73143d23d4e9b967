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
// One operating point end to end on the device (psf_report, ref:75-166, with
// generate_psfs :168-363 and get_width :653-668 inside): illumination, the
// Gaussian fits, the integer rescan ratio the fit decides (:252-256), the
// rescan / descan system PSFs, doses -- one CTA per sweep point, each with its
// own grid size, nothing returns to the host in between (the reference and the
// two-call path above fit on the host: ~1.5 ms of scipy per fit around ~35 us
// of kernels).
// ---------------------------------------------------------------------------
enum { kFitLanes = 64, kFitSums = 5, kPsfReportScalars = 16 };   // (kFitSums * kFitLanes >= 2 * 5 * kMpLanes)
// scalar slots of one report
enum { PR_EX_SIGMA = 0, PR_STED_SIGMA, PR_RESCAN_SIGMA, PR_RATIO_REAL, PR_RATIO, PR_EXC_DOSE,
       PR_DEP_DOSE, PR_EMISSION, PR_FIT_ITERS, PR_FIT_STATUS, PR_EX_MAX_OK, PR_STED_MAX_OK,
       PR_RESCAN_MAX_OK };

struct PsfReportArgs {
    int psf_type;              // 0 point, 1 line
    int nmax;                  // largest grid of the batch (stride of the per-point buffers)
    int tap_stride;
    const int* n;              // [batch] grid size (num_steps, :92)
    const int* radius;         // [batch] FIR radius
    const double* taps;        // [batch][tap_stride]
    const double* blur_sigma;  // [batch]
    const double* exc_brightness;
    const double* dep_brightness;
    double* scalars;           // [batch][kPsfReportScalars]
    double* psfs;              // null, or [batch][7][nmax*nmax]: excitation, depletion, excitation_fraction,
                               // depletion_fraction, sted, rescan_sted, descan_sted (n*n packed at the front)
};

// per-CTA scratch (dynamic shared memory on the device): five rows of nmax + the fit sums
struct PsfReportSmem {
    double *k1, *g2, *q, *row_ex, *row_sted;
    double *fvec, *wa4, *jac;   // the fit: residuals, a work vector, 3 Jacobian columns
    double* red;      // [kFitSums][kFitLanes] partial sums of the reductions
    double* sc;       // [8] report scalars
    double* fit;      // [8] result of the last fit
    int nmax;
    LSTED_HD static size_t doubles(int nmax) { return 10 * (size_t)nmax + kFitSums * kFitLanes + 16; }
    LSTED_HD void carve(double* base, int nmax_) {
        nmax = nmax_;
        k1 = base; g2 = k1 + nmax; q = g2 + nmax; row_ex = q + nmax; row_sted = row_ex + nmax;
        fvec = row_sted + nmax; wa4 = fvec + nmax; jac = wa4 + nmax;
        red = jac + 3 * (size_t)nmax; sc = red + kFitSums * kFitLanes; fit = sc + 8;
    }
};

// ---------------------------------------------------------------------------
// get_width (ref:653-668) = scipy.optimize.curve_fit(gauss, range(n), y, p0=[1, n/2, 1]),
// i.e. MINPACK's lmdif (Levenberg-Marquardt with a forward-difference Jacobian, QR with
// column pivoting, the More' trust-region parameter) at scipy's defaults ftol = xtol =
// 1.49012e-8, gtol = 0, maxfev = 200*(n+1), epsfcn = eps, factor = 100, mode 1.  The fitted
// width is only defined by where THAT iteration stops (up to 7e-6 away from the true
// least-squares minimum on saturated profiles, anywhere on sub-pixel ones), and it decides an
// integer array shape (:252-256) -- so the device fit restates lmdif step for step, with its
// summation orders (the published algorithm: More', Garbow, Hillstrom, MINPACK-1, 1980;
// routines lmdif, lmpar, qrfac, qrsolv, fdjac2, enorm).  Given the same profile the CPU replay
// reproduces scipy bit for bit (tests/test_host_mirror.py); on the GPU only exp() can differ
// in the last place.
// ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
LSTED_HD double nofma_mul(double a, double b) { return __dmul_rn(a, b); }
LSTED_HD double nofma_sub(double a, double b) { return __dsub_rn(a, b); }
LSTED_HD double nofma_add(double a, double b) { return __dadd_rn(a, b); }
#else
LSTED_HD double nofma_mul(double a, double b) { return a * b; }
LSTED_HD double nofma_sub(double a, double b) { return a - b; }
LSTED_HD double nofma_add(double a, double b) { return a + b; }
#endif

LSTED_HD double mp_enorm(const double* x, int n) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)n;
    for (int i = 0; i < n; ++i) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 = nofma_add(s2, nofma_mul(xabs, xabs));
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                const double t = x3max / xabs;
                s3 = nofma_add(1.0, nofma_mul(s3, nofma_mul(t, t)));
                x3max = xabs;
            } else if (xabs != 0) {
                const double t = xabs / x3max;
                s3 = nofma_add(s3, nofma_mul(t, t));
            }
        } else {
            if (xabs > x1max) {
                const double t = x1max / xabs;
                s1 = nofma_add(1.0, nofma_mul(s1, nofma_mul(t, t)));
                x1max = xabs;
            } else {
                const double t = xabs / x1max;
                s1 = nofma_add(s1, nofma_mul(t, t));
            }
        }
    }
    if (s1 != 0) return nofma_mul(x1max, sqrt(nofma_add(s1, (s2 / x1max) / x1max)));
    if (s2 != 0) {
        if (s2 >= x3max) return sqrt(nofma_mul(s2, nofma_add(1.0, nofma_mul(x3max / s2, nofma_mul(x3max, s3)))));
        return sqrt(nofma_mul(x3max, nofma_add(s2 / x3max, nofma_mul(x3max, s3))));
    }
    return nofma_mul(x3max, sqrt(s3));
}
enum { MP_N = 3, kMpLanes = 32 };

// The O(m) loops of the algorithm, spread over the CTA: element-wise steps by all threads,
// sums as kMpLanes strided partial sums that every thread then combines in lane order (so all
// threads hold the same scalar and the control flow stays uniform; the summation order differs
// from MINPACK's serial loops by rounding only -- 1e-16, against the 1e-8 that exp()'s last
// place puts into the forward-difference Jacobian).  Partial sums ping-pong between two
// buffers: a thread can run at most one barrier ahead of the slowest.
template <class Ctx> struct MpVec {
    Ctx& cx;
    double* red;    // [2][5][kMpLanes]
    int flip;
    LSTED_HD MpVec(Ctx& c, double* r) : cx(c), red(r), flip(0) {}
    LSTED_HD void sync() { cx.parallel_for(0, [](int) {}); }
    LSTED_HD double dot(const double* a, const double* b, int from, int to) {
        double* buf = red + (flip ^= 1) * 5 * kMpLanes;
        cx.parallel_for(kMpLanes, [&](int lane) {
            double s = 0;
            for (int i = from + lane; i < to; i += kMpLanes) s = nofma_add(s, nofma_mul(a[i], b[i]));
            buf[lane] = s;
        });
        double s = 0;
        for (int l = 0; l < kMpLanes; ++l) s = nofma_add(s, buf[l]);
        return s;
    }
    // enorm: the three magnitude classes of MINPACK's scaled sum of squares, merged over lanes
    LSTED_HD double enorm(const double* x, int from, int to) {
        const double rdwarf = 3.834e-20, rgiant = 1.304e19;
        const int n = to - from;
        if (n <= 0) return 0.0;
        const double agiant = rgiant / (double)n;
        double* buf = red + (flip ^= 1) * 5 * kMpLanes;
        cx.parallel_for(kMpLanes, [&](int lane) {
            double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
            for (int i = from + lane; i < to; i += kMpLanes) {
                const double xabs = fabs(x[i]);
                if (xabs > rdwarf && xabs < agiant) {
                    s2 = nofma_add(s2, nofma_mul(xabs, xabs));
                } else if (xabs <= rdwarf) {
                    if (xabs > x3max) {
                        const double t = x3max / xabs;
                        s3 = nofma_add(1.0, nofma_mul(s3, nofma_mul(t, t)));
                        x3max = xabs;
                    } else if (xabs != 0) {
                        const double t = xabs / x3max;
                        s3 = nofma_add(s3, nofma_mul(t, t));
                    }
                } else if (xabs > x1max) {
                    const double t = x1max / xabs;
                    s1 = nofma_add(1.0, nofma_mul(s1, nofma_mul(t, t)));
                    x1max = xabs;
                } else {
                    const double t = xabs / x1max;
                    s1 = nofma_add(s1, nofma_mul(t, t));
                }
            }
            buf[lane] = s1; buf[kMpLanes + lane] = x1max; buf[2 * kMpLanes + lane] = s2;
            buf[3 * kMpLanes + lane] = s3; buf[4 * kMpLanes + lane] = x3max;
        });
        double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
        for (int l = 0; l < kMpLanes; ++l) {
            const double p1 = buf[l], m1 = buf[kMpLanes + l], p3 = buf[3 * kMpLanes + l],
                         m3 = buf[4 * kMpLanes + l];
            s2 = nofma_add(s2, buf[2 * kMpLanes + l]);
            if (m1 > x1max) { const double t = x1max / m1; s1 = nofma_add(p1, nofma_mul(s1, nofma_mul(t, t))); x1max = m1; }
            else if (m1 != 0) { const double t = m1 / x1max; s1 = nofma_add(s1, nofma_mul(p1, nofma_mul(t, t))); }
            if (m3 > x3max) { const double t = x3max / m3; s3 = nofma_add(p3, nofma_mul(s3, nofma_mul(t, t))); x3max = m3; }
            else if (m3 != 0) { const double t = m3 / x3max; s3 = nofma_add(s3, nofma_mul(p3, nofma_mul(t, t))); }
        }
        if (s1 != 0) return nofma_mul(x1max, sqrt(nofma_add(s1, (s2 / x1max) / x1max)));
        if (s2 != 0) {
            if (s2 >= x3max) return sqrt(nofma_mul(s2, nofma_add(1.0, nofma_mul(x3max / s2, nofma_mul(x3max, s3)))));
            return sqrt(nofma_mul(x3max, nofma_add(s2 / x3max, nofma_mul(x3max, s3))));
        }
        return nofma_mul(x3max, sqrt(s3));
    }
};

// qrsolv on the 3 x 3 upper triangle r (row i, column j); the lower triangle is scratch
LSTED_HD void mp_qrsolv(double r[MP_N][MP_N], const int* ipvt, const double* diag, const double* qtb,
                        double* x, double* sdiag) {
    double wa[MP_N];
    for (int j = 0; j < MP_N; ++j) {
        for (int i = j; i < MP_N; ++i) r[i][j] = r[j][i];
        x[j] = r[j][j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < MP_N; ++j) {
        const int l = ipvt[j];
        if (diag[l] != 0) {
            for (int k = j; k < MP_N; ++k) sdiag[k] = 0;
            sdiag[j] = diag[l];
            double qtbpj = 0;
            for (int k = j; k < MP_N; ++k) {
                if (sdiag[k] == 0) continue;
                double sn, cs;
                if (fabs(r[k][k]) < fabs(sdiag[k])) {
                    const double cotan = r[k][k] / sdiag[k];
                    sn = 0.5 / sqrt(nofma_add(0.25, nofma_mul(0.25, nofma_mul(cotan, cotan))));
                    cs = nofma_mul(sn, cotan);
                } else {
                    const double tn = sdiag[k] / r[k][k];
                    cs = 0.5 / sqrt(nofma_add(0.25, nofma_mul(0.25, nofma_mul(tn, tn))));
                    sn = nofma_mul(cs, tn);
                }
                r[k][k] = nofma_add(nofma_mul(cs, r[k][k]), nofma_mul(sn, sdiag[k]));
                const double temp = nofma_add(nofma_mul(cs, wa[k]), nofma_mul(sn, qtbpj));
                qtbpj = nofma_add(nofma_mul(-sn, wa[k]), nofma_mul(cs, qtbpj));
                wa[k] = temp;
                for (int i = k + 1; i < MP_N; ++i) {
                    const double t2 = nofma_add(nofma_mul(cs, r[i][k]), nofma_mul(sn, sdiag[i]));
                    sdiag[i] = nofma_add(nofma_mul(-sn, r[i][k]), nofma_mul(cs, sdiag[i]));
                    r[i][k] = t2;
                }
            }
        }
        sdiag[j] = r[j][j];
        r[j][j] = x[j];
    }
    int nsing = MP_N;
    for (int j = 0; j < MP_N; ++j) {
        if (sdiag[j] == 0 && nsing == MP_N) nsing = j;
        if (nsing < MP_N) wa[j] = 0;
    }
    for (int k = 0; k < nsing; ++k) {
        const int j = nsing - k - 1;
        double sum = 0;
        for (int i = j + 1; i < nsing; ++i) sum = nofma_add(sum, nofma_mul(r[i][j], wa[i]));
        wa[j] = nofma_sub(wa[j], sum) / sdiag[j];
    }
    for (int j = 0; j < MP_N; ++j) x[ipvt[j]] = wa[j];
}

// lmpar: the Levenberg-Marquardt parameter for the trust region `delta`; returns par, x
LSTED_HD double mp_lmpar(double r[MP_N][MP_N], const int* ipvt, const double* diag, const double* qtb,
                         double delta, double par, double* x) {
    const double dwarf = 2.2250738585072014e-308;
    double wa1[MP_N], wa2[MP_N], sdiag[MP_N];
    int nsing = MP_N;
    for (int j = 0; j < MP_N; ++j) {
        wa1[j] = qtb[j];
        if (r[j][j] == 0 && nsing == MP_N) nsing = j;
        if (nsing < MP_N) wa1[j] = 0;
    }
    for (int k = 0; k < nsing; ++k) {
        const int j = nsing - k - 1;
        wa1[j] /= r[j][j];
        const double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] = nofma_sub(wa1[i], nofma_mul(r[i][j], temp));
    }
    for (int j = 0; j < MP_N; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < MP_N; ++j) wa2[j] = nofma_mul(diag[j], x[j]);
    double dxnorm = mp_enorm(wa2, MP_N);
    double fp = nofma_sub(dxnorm, delta);
    if (fp <= nofma_mul(0.1, delta)) return 0.0;
    double parl = 0;
    if (nsing >= MP_N) {
        for (int j = 0; j < MP_N; ++j) {
            const int l = ipvt[j];
            wa1[j] = nofma_mul(diag[l], wa2[l] / dxnorm);
        }
        for (int j = 0; j < MP_N; ++j) {
            double sum = 0;
            for (int i = 0; i < j; ++i) sum = nofma_add(sum, nofma_mul(r[i][j], wa1[i]));
            wa1[j] = nofma_sub(wa1[j], sum) / r[j][j];
        }
        const double temp = mp_enorm(wa1, MP_N);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < MP_N; ++j) {
        double sum = 0;
        for (int i = 0; i <= j; ++i) sum = nofma_add(sum, nofma_mul(r[i][j], qtb[i]));
        wa1[j] = sum / diag[ipvt[j]];
    }
    const double gnorm = mp_enorm(wa1, MP_N);
    double paru = gnorm / delta;
    if (paru == 0) paru = dwarf / (delta < 0.1 ? delta : 0.1);
    par = par > parl ? par : parl;
    par = par < paru ? par : paru;
    if (par == 0) par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (par == 0) { const double t = nofma_mul(0.001, paru); par = dwarf > t ? dwarf : t; }
        double temp = sqrt(par);
        for (int j = 0; j < MP_N; ++j) wa1[j] = nofma_mul(temp, diag[j]);
        mp_qrsolv(r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < MP_N; ++j) wa2[j] = nofma_mul(diag[j], x[j]);
        dxnorm = mp_enorm(wa2, MP_N);
        temp = fp;
        fp = nofma_sub(dxnorm, delta);
        if (fabs(fp) <= nofma_mul(0.1, delta) || (parl == 0 && fp <= temp && temp < 0) || iter == 10) break;
        for (int j = 0; j < MP_N; ++j) {
            const int l = ipvt[j];
            wa1[j] = nofma_mul(diag[l], wa2[l] / dxnorm);
        }
        for (int j = 0; j < MP_N; ++j) {
            wa1[j] /= sdiag[j];
            const double t = wa1[j];
            for (int i = j + 1; i < MP_N; ++i) wa1[i] = nofma_sub(wa1[i], nofma_mul(r[i][j], t));
        }
        temp = mp_enorm(wa1, MP_N);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0) parl = parl > par ? parl : par;
        if (fp < 0) paru = paru < par ? paru : par;
        const double t = nofma_add(par, parc);
        par = parl > t ? parl : t;
    }
    return par;
}

// out: A, mu, sigma, function evaluations, MINPACK info (1-4 = converged like scipy accepts).
// All scalars of the algorithm live in registers, identical in every thread.
template <class Ctx>
LSTED_HD void gauss_fit_body(Ctx& cx, const double* y, int n, PsfReportSmem* sm, double* out) {
    const int m = n;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
    const double epsmch = 2.220446049250313e-16;
    const int maxfev = 200 * (MP_N + 1);
    double* const fvec = sm->fvec;
    double* const wa4 = sm->wa4;
    double* const col[MP_N] = {sm->jac, sm->jac + sm->nmax, sm->jac + 2 * (size_t)sm->nmax};
    MpVec<Ctx> v(cx, sm->red);
    // residual gauss(i; p) - y[i] (curve_fit's wrapped function), numpy's operation order
    auto residual = [&](const double* p, int i) {
        const double d = nofma_sub((double)i, p[1]);
        const double den = nofma_mul(2.0, nofma_mul(p[2], p[2]));
        const double e = exp(-nofma_mul(d, d) / den);
        return nofma_sub(nofma_mul(p[0], e), y[i]);
    };
    double x[MP_N] = {1.0, (double)n / 2.0, 1.0};
    double diag[MP_N] = {0, 0, 0}, qtf[MP_N], rdiag[MP_N], acnorm[MP_N], wa[MP_N], trial[MP_N];
    double r[MP_N][MP_N];
    int ipvt[MP_N];
    double par = 0, delta = 0, xnorm = 0, gnorm = 0;
    int iter = 1, info = 0, nfev = 1;
    v.sync();   // (the caller's writes of y)
    cx.parallel_for(m, [&](int i) { fvec[i] = residual(x, i); });
    double fnorm = v.enorm(fvec, 0, m);
    const double eps = sqrt(epsmch);
    for (;;) {   // outer loop
        // forward-difference Jacobian (fdjac2)
        v.sync();
        for (int j = 0; j < MP_N; ++j) {
            for (int k = 0; k < MP_N; ++k) trial[k] = x[k];
            double h = nofma_mul(eps, fabs(x[j]));
            if (h == 0) h = eps;
            trial[j] = nofma_add(x[j], h);
            cx.parallel_for(m, [&](int i) { col[j][i] = nofma_sub(residual(trial, i), fvec[i]) / h; });
        }
        nfev += MP_N;
        // qrfac: Householder QR with column pivoting
        for (int j = 0; j < MP_N; ++j) {
            acnorm[j] = v.enorm(col[j], 0, m);
            rdiag[j] = acnorm[j]; wa[j] = rdiag[j]; ipvt[j] = j;
        }
        for (int j = 0; j < MP_N && j < m; ++j) {
            int kmax = j;
            for (int k = j; k < MP_N; ++k)
                if (rdiag[k] > rdiag[kmax]) kmax = k;
            if (kmax != j) {
                v.sync();
                cx.parallel_for(m, [&](int i) { const double t = col[j][i]; col[j][i] = col[kmax][i]; col[kmax][i] = t; });
                rdiag[kmax] = rdiag[j]; wa[kmax] = wa[j];
                const int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
            }
            double ajnorm = v.enorm(col[j], j, m);
            if (ajnorm != 0) {
                if (col[j][j] < 0) ajnorm = -ajnorm;
                v.sync();
                cx.parallel_for(m - j, [&](int w) {
                    const int i = j + w;
                    double t = col[j][i] / ajnorm;
                    if (i == j) t = nofma_add(t, 1.0);
                    col[j][i] = t;
                });
                const double ajj = col[j][j];
                for (int k = j + 1; k < MP_N; ++k) {
                    const double sum = v.dot(col[j], col[k], j, m);
                    const double temp = sum / ajj;
                    v.sync();
                    cx.parallel_for(m - j, [&](int w) {
                        const int i = j + w;
                        col[k][i] = nofma_sub(col[k][i], nofma_mul(temp, col[j][i]));
                    });
                    if (rdiag[k] != 0) {
                        const double t = col[k][j] / rdiag[k];
                        const double u = nofma_sub(1.0, nofma_mul(t, t));
                        rdiag[k] = nofma_mul(rdiag[k], sqrt(u > 0 ? u : 0.0));
                        const double q = rdiag[k] / wa[k];
                        if (nofma_mul(0.05, nofma_mul(q, q)) <= epsmch) {
                            rdiag[k] = v.enorm(col[k], j + 1, m);
                            wa[k] = rdiag[k];
                        }
                    }
                }
            }
            rdiag[j] = -ajnorm;
        }
        if (iter == 1) {
            double wa3[MP_N];
            for (int j = 0; j < MP_N; ++j) {
                diag[j] = acnorm[j];
                if (acnorm[j] == 0) diag[j] = 1.0;
                wa3[j] = nofma_mul(diag[j], x[j]);
            }
            xnorm = mp_enorm(wa3, MP_N);
            delta = nofma_mul(factor, xnorm);
            if (delta == 0) delta = factor;
        }
        // first n components of Q^T fvec
        v.sync();
        cx.parallel_for(m, [&](int i) { wa4[i] = fvec[i]; });
        for (int j = 0; j < MP_N; ++j) {
            const double cjj = col[j][j];
            if (cjj != 0) {
                const double sum = v.dot(col[j], wa4, j, m);
                const double temp = -sum / cjj;
                v.sync();
                cx.parallel_for(m - j, [&](int w) {
                    const int i = j + w;
                    wa4[i] = nofma_add(wa4[i], nofma_mul(col[j][i], temp));
                });
            }
            qtf[j] = wa4[j];
        }
        // R: the strict upper triangle sits in the columns, the diagonal in rdiag
        for (int i = 0; i < MP_N; ++i)
            for (int j = 0; j < MP_N; ++j) r[i][j] = i < j ? col[j][i] : (i == j ? rdiag[j] : 0.0);
        gnorm = 0;
        if (fnorm != 0) {
            for (int j = 0; j < MP_N; ++j) {
                const int l = ipvt[j];
                if (acnorm[l] != 0) {
                    double sum = 0;
                    for (int i = 0; i <= j; ++i) sum = nofma_add(sum, nofma_mul(r[i][j], qtf[i] / fnorm));
                    const double g = fabs(sum / acnorm[l]);
                    gnorm = g > gnorm ? g : gnorm;
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < MP_N; ++j) diag[j] = diag[j] > acnorm[j] ? diag[j] : acnorm[j];
        for (;;) {   // inner loop
            double rr[MP_N][MP_N], p[MP_N], wa1[MP_N], wa3[MP_N];
            for (int i = 0; i < MP_N; ++i)
                for (int j = 0; j < MP_N; ++j) rr[i][j] = r[i][j];
            par = mp_lmpar(rr, ipvt, diag, qtf, delta, par, p);
            for (int j = 0; j < MP_N; ++j) {
                wa1[j] = -p[j];
                trial[j] = nofma_add(x[j], wa1[j]);
                wa3[j] = nofma_mul(diag[j], wa1[j]);
            }
            const double pnorm = mp_enorm(wa3, MP_N);
            if (iter == 1) delta = delta < pnorm ? delta : pnorm;
            v.sync();
            cx.parallel_for(m, [&](int i) { wa4[i] = residual(trial, i); });
            nfev += 1;
            const double fnorm1 = v.enorm(wa4, 0, m);
            double actred = -1.0;
            if (nofma_mul(0.1, fnorm1) < fnorm) {
                const double t = fnorm1 / fnorm;
                actred = nofma_sub(1.0, nofma_mul(t, t));
            }
            for (int i = 0; i < MP_N; ++i) wa3[i] = 0;
            for (int j = 0; j < MP_N; ++j) {
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] = nofma_add(wa3[i], nofma_mul(r[i][j], temp));
            }
            const double temp1 = mp_enorm(wa3, MP_N) / fnorm;
            const double temp2 = nofma_mul(sqrt(par), pnorm) / fnorm;
            const double t11 = nofma_mul(temp1, temp1), t22 = nofma_mul(temp2, temp2);
            const double prered = nofma_add(t11, t22 / 0.5);
            const double dirder = -nofma_add(t11, t22);
            double ratio = 0;
            if (prered != 0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp;
                if (actred >= 0) temp = 0.5;
                else temp = nofma_mul(0.5, dirder) / nofma_add(dirder, nofma_mul(0.5, actred));
                if (nofma_mul(0.1, fnorm1) >= fnorm || temp < 0.1) temp = 0.1;
                const double pn = pnorm / 0.1;
                delta = nofma_mul(temp, delta < pn ? delta : pn);
                par = par / temp;
            } else if (par == 0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = nofma_mul(0.5, par);
            }
            if (ratio >= 1e-4) {   // successful iteration
                double wa2[MP_N];
                for (int j = 0; j < MP_N; ++j) {
                    x[j] = trial[j];
                    wa2[j] = nofma_mul(diag[j], x[j]);
                }
                v.sync();
                cx.parallel_for(m, [&](int i) { fvec[i] = wa4[i]; });
                xnorm = mp_enorm(wa2, MP_N);
                fnorm = fnorm1;
                iter += 1;
            }
            const bool small = fabs(actred) <= ftol && prered <= ftol && nofma_mul(0.5, ratio) <= 1.0;
            if (small) info = 1;
            if (delta <= nofma_mul(xtol, xnorm)) info = 2;
            if (small && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && nofma_mul(0.5, ratio) <= 1.0) info = 6;
            if (delta <= nofma_mul(epsmch, xnorm)) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) break;
            if (ratio >= 1e-4) break;
        }
        if (info != 0) break;
    }
    v.sync();
    cx.parallel_for(1, [&](int) {
        out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = (double)nfev; out[4] = (double)info;
    });
}

// get_width for a batch of profiles of one length: one CTA per row
struct GaussFitArgs {
    int n;
    const double* rows;   // [batch][n]
    double* out;          // [batch][5]: A, mu, sigma, function evaluations, MINPACK info
};
template <class Ctx>
LSTED_HD void gauss_fit_rows_body(Ctx& cx, int b, const GaussFitArgs& a, PsfReportSmem* sm) {
    const double* src = a.rows + (size_t)a.n * b;
    cx.parallel_for(a.n, [&](int i) { sm->row_ex[i] = src[i]; });
    gauss_fit_body(cx, sm->row_ex, a.n, sm, sm->fit);
    cx.parallel_for(5, [&](int i) { a.out[5 * (size_t)b + i] = sm->fit[i]; });
}

// q[j] of the rescan PSF (see the header of this file); rescan[y][j] = k1[y] * q[j]
template <class Ctx>
LSTED_HD void rescan_profile(Ctx& cx, int n, int R, const double* k1, const double* row, double* q) {
    const int c = n / 2, W = R * n;
    cx.parallel_for(n, [&](int j) {
        double acc_q = 0;
        for (int s = 0; s < n; ++s) {
            int base = (int)((((long long)(j - s) * R + c - R / 2) % W + W) % W);
            double acc = 0;
            if (base < n) {
                const int r1 = (n - base) < R ? (n - base) : R;
                acc += cyclic_run_sum(k1, n, base, base + r1, s - c);
            }
            if (W - base < R) {
                const int r0 = W - base;
                const int r1 = (r0 + n) < R ? (r0 + n) : R;
                acc += cyclic_run_sum(k1, n, 0, r1 - r0, s - c);
            }
            acc_q += row[n - 1 - s] * acc;
        }
        q[j] = acc_q;
    });
}

template <class Ctx>
LSTED_HD void psf_report_body(Ctx& cx, int b, const PsfReportArgs& a, PsfReportSmem* sm) {
    const int n = a.n[b], radius = a.radius[b], mid = n / 2;
    const double* taps = a.taps + (size_t)a.tap_stride * b;
    const bool point = a.psf_type == 0;
    const size_t img = (size_t)n * n;
    double* sc = a.scalars + (size_t)kPsfReportScalars * b;
    // 1-D profiles: first and second blur of the delta (as psf_profiles)
    cx.parallel_for(n, [&](int i) {
        const int t = radius + mid - i;
        sm->k1[i] = (t >= 0 && t <= 2 * radius) ? taps[t] : 0.0;
    });
    cx.parallel_for(n, [&](int i) {
        double s = 0;
        for (int t = 0; t <= 2 * radius; ++t) s += taps[t] * sm->k1[reflect_idx(i + t - radius, n)];
        sm->g2[i] = s;
    });
    cx.parallel_for(1, [&](int) {
        double m1 = sm->k1[0], m2 = sm->g2[0], ks = 0;
        for (int i = 0; i < n; ++i) {
            m1 = sm->k1[i] > m1 ? sm->k1[i] : m1;
            m2 = sm->g2[i] > m2 ? sm->g2[i] : m2;
            ks += sm->k1[i];
        }
        sm->sc[0] = point ? m1 * m1 : m1;
        sm->sc[1] = point ? m2 * m2 : m2;
        sm->sc[3] = ks;
    });
    const double inner_max = sm->sc[0], outer_max = sm->sc[1];
    const size_t count = point ? img : (size_t)n;
    cx.parallel_for(kFitLanes, [&](int lane) {
        double m = -1e300;
        for (size_t p = lane; p < count; p += kFitLanes) {
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
        for (int i = 1; i < kFitLanes; ++i) m = sm->red[i] > m ? sm->red[i] : m;
        sm->sc[2] = m;
    });
    const double exc_scale = a.exc_brightness[b] / inner_max;
    const double dep_scale = a.dep_brightness[b] / sm->sc[2];
    // the five illumination values of pixel (y, x)
    auto pixel = [&](int y, int x, double* v) {
        const double inner = point ? sm->k1[y] * sm->k1[x] : sm->k1[x];
        const double outer = point ? sm->g2[y] * sm->g2[x] : sm->g2[x];
        v[0] = inner * exc_scale;
        v[1] = (outer / outer_max - inner / inner_max) * dep_scale;
        v[2] = 1.0 - exp2(-v[0]);
        v[3] = exp2(-v[1]);
        v[4] = v[2] * v[3];
    };
    // centre rows (:103-104) and the doses (:134-144): sums over the area (point) or over the
    // centre row (line); also: is the centre-row maximum the global one (the asserts :105-106)
    cx.parallel_for(n, [&](int x) {
        double v[5];
        pixel(mid, x, v);
        sm->row_ex[x] = v[0];
        sm->row_sted[x] = v[4];
    });
    cx.parallel_for(kFitLanes, [&](int lane) {
        double se = 0, sd = 0, ss = 0, me = -1e300, ms = -1e300;
        for (size_t p = lane; p < count; p += kFitLanes) {
            const int y = point ? (int)(p / n) : mid, x = (int)(p - (size_t)(point ? y : 0) * n);
            double v[5];
            pixel(y, x, v);
            se += v[0]; sd += v[1]; ss += v[4];
            me = v[0] > me ? v[0] : me;
            ms = v[4] > ms ? v[4] : ms;
        }
        sm->red[lane] = se; sm->red[kFitLanes + lane] = sd; sm->red[2 * kFitLanes + lane] = ss;
        sm->red[3 * kFitLanes + lane] = me; sm->red[4 * kFitLanes + lane] = ms;
    });
    cx.parallel_for(1, [&](int) {
        double se = 0, sd = 0, ss = 0, me = -1e300, ms = -1e300, re = -1e300, rs = -1e300;
        for (int l = 0; l < kFitLanes; ++l) {
            se += sm->red[l]; sd += sm->red[kFitLanes + l]; ss += sm->red[2 * kFitLanes + l];
            me = sm->red[3 * kFitLanes + l] > me ? sm->red[3 * kFitLanes + l] : me;
            ms = sm->red[4 * kFitLanes + l] > ms ? sm->red[4 * kFitLanes + l] : ms;
        }
        for (int x = 0; x < n; ++x) {
            re = sm->row_ex[x] > re ? sm->row_ex[x] : re;
            rs = sm->row_sted[x] > rs ? sm->row_sted[x] : rs;
        }
        sc[PR_EXC_DOSE] = se; sc[PR_DEP_DOSE] = sd; sc[PR_EMISSION] = ss;
        sc[PR_EX_MAX_OK] = (re == me) ? 1.0 : 0.0;
        sc[PR_STED_MAX_OK] = (rs == ms) ? 1.0 : 0.0;
        sc[PR_RESCAN_MAX_OK] = 1.0;
        sc[PR_RESCAN_SIGMA] = 0.0; sc[PR_RATIO_REAL] = 0.0; sc[PR_RATIO] = 0.0;
    });
    double* const fit = sm->fit;
    gauss_fit_body(cx, sm->row_ex, n, sm, fit);
    // scipy accepts MINPACK's info 1..4 and raises otherwise: anything else is reported
    auto fit_status = [&](double info) { return (info >= 1.0 && info <= 4.0) ? 0.0 : 100.0 + info; };
    cx.parallel_for(1, [&](int) {
        sc[PR_EX_SIGMA] = fit[2]; sc[PR_FIT_ITERS] = fit[3]; sc[PR_FIT_STATUS] = fit_status(fit[4]);
    });
    gauss_fit_body(cx, sm->row_sted, n, sm, fit);
    cx.parallel_for(1, [&](int) {
        sc[PR_STED_SIGMA] = fit[2];
        sc[PR_FIT_ITERS] += fit[3];
        if (fit_status(fit[4]) != 0.0) sc[PR_FIT_STATUS] = fit_status(fit[4]);
    });
    int R = 0;
    if (!point) {
        // :252-256: the integer rescan ratio comes from the fitted width
        cx.parallel_for(1, [&](int) {
            const double w = a.blur_sigma[b] / sc[PR_STED_SIGMA];
            const double rr = w * w + 1.0;
            double ri = rint(rr);            // numpy.round: half to even
            if (!(ri >= 1.0)) ri = 1.0;
            if (ri * n > 16777216.0) { ri = 1.0; sc[PR_FIT_STATUS] = 99.0; }
            sc[PR_RATIO_REAL] = rr;
            sc[PR_RATIO] = ri;
        });
        R = (int)sc[PR_RATIO];
        rescan_profile(cx, n, R, sm->k1, sm->row_sted, sm->q);
        // centre row of the rescan PSF and its fit (:119-121); its maximum against the plane's
        cx.parallel_for(n, [&](int x) { sm->row_ex[x] = sm->k1[mid] * sm->q[x]; });
        cx.parallel_for(1, [&](int) {
            double kmax = sm->k1[0], rmax = -1e300;
            for (int i = 0; i < n; ++i) kmax = sm->k1[i] > kmax ? sm->k1[i] : kmax;
            bool ok = sm->k1[mid] == kmax;
            for (int x = 0; x < n; ++x) rmax = sm->row_ex[x] > rmax ? sm->row_ex[x] : rmax;
            sc[PR_RESCAN_MAX_OK] = (ok && rmax >= 0.0) ? 1.0 : 0.0;
        });
        gauss_fit_body(cx, sm->row_ex, n, sm, fit);
        cx.parallel_for(1, [&](int) {
            sc[PR_RESCAN_SIGMA] = fit[2];
            sc[PR_FIT_ITERS] += fit[3];
            if (fit_status(fit[4]) != 0.0) sc[PR_FIT_STATUS] = fit_status(fit[4]);
        });
    }
    if (!a.psfs) return;
    double* out = a.psfs + (size_t)7 * a.nmax * a.nmax * b;
    const size_t plane = (size_t)a.nmax * a.nmax;
    const double ksum = sm->sc[3];
    cx.parallel_for((int)img, [&](int p) {
        const int y = p / n, x = p - y * n;
        double v[5];
        pixel(y, x, v);
        for (int i = 0; i < 5; ++i) out[plane * i + p] = v[i];
        if (!point) {
            out[plane * 5 + p] = sm->k1[y] * sm->q[x];
            out[plane * 6 + p] = sm->row_sted[n - 1 - x] * (sm->k1[y] * ksum);
        }
    });
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
__global__ void __launch_bounds__(kPsfThreads) psf_report_kernel(PsfReportArgs a) {
    extern __shared__ double psf_report_dyn[];
    PsfReportSmem sm;
    sm.carve(psf_report_dyn, a.nmax);
    PsfDeviceCtx cx;
    psf_report_body(cx, blockIdx.x, a, &sm);
}
__global__ void __launch_bounds__(kPsfThreads) gauss_fit_kernel(GaussFitArgs a) {
    extern __shared__ double psf_report_dyn[];
    PsfReportSmem sm;
    sm.carve(psf_report_dyn, a.n);
    PsfDeviceCtx cx;
    gauss_fit_rows_body(cx, blockIdx.x, a, &sm);
}
__global__ void __launch_bounds__(kPsfThreads) psf_rotate_kernel(PsfRotateArgs a) {
    PsfDeviceCtx cx;
    psf_rotate_body(cx, blockIdx.x, a);
}
#endif

}  // namespace lsted
