// Host-side planning for the convolution engine: FFT lengths, radix
// schedules, twiddle tables, shared-memory pitches.  Pure C++ (no CUDA).
#pragma once
#include <math.h>
#include <stddef.h>
#include <string.h>
#include <vector>
#include "conv_bodies.cuh"

namespace lsted {

enum { kSmemLimit = 227 * 1024, kMinFftLen = 8 };

// Smallest L >= n of the form 2^a 3^b 5^c.
inline int next_smooth_len(int n) {
    if (n < kMinFftLen) n = kMinFftLen;
    for (int L = n;; ++L) {
        int m = L;
        while (m % 2 == 0) m /= 2;
        while (m % 3 == 0) m /= 3;
        while (m % 5 == 0) m /= 5;
        if (m == 1) return L;
    }
}

// Radix schedule, largest radix first (the first pass needs no twiddles).
inline bool make_fft_plan(int L, FftPlan* p) {
    memset(p, 0, sizeof(*p));
    p->L = L;
    int a = 0, b = 0, c = 0, m = L;
    while (m % 2 == 0) { m /= 2; ++a; }
    while (m % 3 == 0) { m /= 3; ++b; }
    while (m % 5 == 0) { m /= 5; ++c; }
    if (m != 1 || L < 2) return false;
    int n = 0;
    while (a >= 4) { p->radix[n++] = 16; a -= 4; }
    while (b >= 2) { p->radix[n++] = 9; b -= 2; }
    if (a == 3) { p->radix[n++] = 8; a = 0; }
    while (c >= 1) { p->radix[n++] = 5; --c; }
    if (a == 2) { p->radix[n++] = 4; a = 0; }
    if (b == 1) { p->radix[n++] = 3; b = 0; }
    if (a == 1) { p->radix[n++] = 2; a = 0; }
    if (n > kMaxPasses) return false;
    p->npass = n;
    return true;
}

template <typename T> inline void fill_twiddles(int L, cplx<T>* out) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int m = 0; m < L; ++m) {
        const long double t = two_pi * (long double)m / (long double)L;
        out[m].x = (T)cosl(t);
        out[m].y = (T)(-sinl(t));
    }
}

inline int smem_pitch(int L, int elem_bytes, int C) {
    const int shift = elem_bytes == 8 ? 4 : 3;      // must match PadShift<T>
    const int units = 128 / elem_bytes;             // elements per 128-byte wavefront
    int lp = (L - 1) + ((L - 1) >> shift) + 1;
    lp = (lp + units - 1) / units * units;
    if (C > 1 && C <= units) lp += units / C;       // spreads c*Lp over the banks
    return lp;
}

// Geometry for an Ny x Nx image and ny x nx PSFs.  `cplx_bytes` = 8 (fp32)
// or 16 (fp64).  Returns an empty string on success, else the reason.
// `fast_C(L)` > 0 means the backend has a compile-time plan for column length L
// that wants C columns per block (it then needs only the generic OTF kernel to
// fit: two buffers instead of three).
// `force_L` > 0 pins both transform lengths (overlap-save tiles: the window is as
// long as the transform, only its interior is alias-free and only that is kept).
inline const char* make_geom(int Ny, int Nx, int ny, int nx, int cplx_bytes, ConvGeom* g,
                             int (*fast_C)(int L, int cplx_bytes) = 0, int force_L = 0) {
    memset(g, 0, sizeof(*g));
    if (Ny < 1 || Nx < 1 || ny < 1 || nx < 1) return "empty image or PSF";
    g->Ny = Ny; g->Nx = Nx;
    g->sy = (ny - 1) / 2; g->sx = (nx - 1) / 2;
    const int hy = g->sy > ny - 1 - g->sy ? g->sy : ny - 1 - g->sy;
    const int hx = g->sx > nx - 1 - g->sx ? g->sx : nx - 1 - g->sx;
    g->Ly = next_smooth_len(Ny + hy);
    g->Lx = next_smooth_len(Nx + hx);
    if (g->Ly < ny) g->Ly = next_smooth_len(ny);  // the whole PSF must fit
    if (g->Lx < nx) g->Lx = next_smooth_len(nx);
    if (force_L > 0) {
        if (force_L < Ny || force_L < Nx || force_L < ny || force_L < nx) return "forced FFT length too short";
        g->Ly = g->Lx = force_L;
    }
    if (!make_fft_plan(g->Ly, &g->py) || !make_fft_plan(g->Lx, &g->px)) return "no FFT plan";
    g->Lxh = g->Lx / 2 + 1;
    g->C = cplx_bytes == 8 ? 4 : 2;
    const int want = fast_C ? fast_C(g->Ly, cplx_bytes) : 0;
    if (want > 0 && (size_t)2 * want * smem_pitch(g->Ly, cplx_bytes, want) * cplx_bytes <= kSmemLimit) {
        g->C = want;
        g->Lpy = smem_pitch(g->Ly, cplx_bytes, g->C);
    } else {
        for (;; g->C /= 2) {
            g->Lpy = smem_pitch(g->Ly, cplx_bytes, g->C);
            if ((size_t)3 * g->C * g->Lpy * cplx_bytes <= kSmemLimit) break;
            if (g->C == 1) return "column transform does not fit in shared memory; tile the object";
        }
    }
    g->nxb = (g->Lxh + g->C - 1) / g->C;
    g->PR = 2;
    for (;; g->PR /= 2) {
        g->Lpx = smem_pitch(g->Lx, cplx_bytes, 1);
        if ((size_t)2 * g->PR * g->Lpx * cplx_bytes <= kSmemLimit) break;
        if (g->PR == 1) return "row transform does not fit in shared memory; tile the object";
    }
    return "";
}

// elements of one spectrum array with `rows` rows: row spectra (XB2) round the rows up to even
inline size_t spec_elems(const ConvGeom& g, int rows) { return (size_t)g.nxb * g.C * even_rows(rows); }
inline size_t otf_elems(const ConvGeom& g) { return (size_t)g.nxb * g.C * g.Ly; }   // XB, all Ly rows
inline int row_blocks(const ConvGeom& g) { return (((g.Ny + 1) / 2) + g.PR - 1) / g.PR; }
inline size_t row_smem_bytes(const ConvGeom& g, int cplx_bytes) {
    return (size_t)2 * g.PR * g.Lpx * cplx_bytes;
}
inline size_t col_smem_bytes(const ConvGeom& g, int cplx_bytes, int nbuf = 3) {
    return (size_t)nbuf * g.C * g.Lpy * cplx_bytes;
}

}  // namespace lsted
