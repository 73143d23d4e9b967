// Scan-position engine (fp64): the explicit acquisition simulation of figure 3 -- descan point,
// non-descanned multipoint, descan line and rescan line STED -- one excitation position at a
// time, for every position of a scan at once.
//
// Replaces simulate_imaging(), figure_generation/line_sted_figure_3.py:76-273, and its helpers
// rotate :382-391, shift :393-396, scale_y :398-409.
//
// What the reference does per scan position (shift_y, shift_x), on the zero-padded object:
//   exc   = shift(centered_exc, s)                 cubic-spline shift by INTEGERS, clipped >= 0
//   glow  = rot_obj * exc
//   desc  = shift(glow, -s)                        ("descanned" back to the optical axis)
//   descan line / point:  inst = gaussian_filter(desc, psf_sigma); one image row / pixel block
//                         of the reconstruction = column sums / total of inst
//   multipoint:           inst = gaussian_filter(glow, psf_sigma); one block per spot = the sum
//                         of the detector region around that spot
//   rescan line:          inst = shift(scale_y(gaussian_filter(desc), 1/(R^2+1)), s), summed on
//                         the camera over the whole scan (one exposure per orientation)
// Structure used here:
//   * a cubic B-spline interpolates its samples, so scipy's shift by integers is a translation
//     with zero fill (to 1e-16 of the image maximum: the prefilter's round-off); exc, glow and
//     desc are therefore never materialised -- desc(y,x) = rot_obj(y+sy, x+sx) * centered_exc(y,x)
//     is formed inside the first blur pass;
//   * all scan positions are independent up to the camera sum, so every kernel runs over
//     (position, pixel) and the sum over positions is one ordered pass per pixel;
//   * scale_y's zoom acts on y only (zoom 1 along x reproduces the samples): one column
//     prefilter + four taps per output pixel, fused with the rescan shift and the clip;
//   * reconstruction and new_signal are assembled on demand from the per-position sums
//     (column sums / totals / region sums): which position wrote a pixel is index arithmetic.
// The figure's display rotations (rotate(exc, -rot), rotate(glow, -rot)) and the object
// rotation use the general spline operator below, which restates scipy.ndimage's
// affine_transform / zoom / shift for order 3: prefilter with mirror ('constant') or reflect
// ('nearest', after a 12-sample edge pad) boundary start, coefficient indices mirrored
// ('constant') or clamped ('nearest') at the ends.
//
// Every kernel is an element functor run by Backend::for_each (a grid-stride CUDA kernel on the
// GPU; tests/host_emul replays the same functors serially on the CPU).
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>
#include <algorithm>
#include "psf_kernels.cuh"

namespace lsted {

enum { SCAN_DESCAN_POINT = 0, SCAN_MULTIPOINT = 1, SCAN_DESCAN_LINE = 2, SCAN_RESCAN_LINE = 3 };
enum { SPLINE_CONSTANT = 0, SPLINE_NEAREST = 1 };
enum { kSplinePrepad = 12, kReduceSegments = 256 };

// ---------------------------------------------------------------------------------------------
// General plane operators, batched over planes [b][n0][n1]
// ---------------------------------------------------------------------------------------------

// scipy.ndimage.correlate1d with a symmetric kernel, mode='reflect': centre tap first, then
// the pairs from the outside in (the order of NI_Correlate1D's symmetric branch).
struct ImgFirFn {
    const double* in; double* out;
    int n0, n1, axis, radius;
    const double* taps;   // [2*radius+1], symmetric
    LSTED_HD void operator()(size_t e) const {
        const int x = (int)(e % n1);
        const size_t t = e / n1;
        const int y = (int)(t % n0);
        const double* p = in + (t / n0) * (size_t)n0 * n1;
        double acc;
        if (axis == 0) {
            acc = p[(size_t)y * n1 + x] * taps[radius];
            for (int d = radius; d >= 1; --d)
                acc += (p[(size_t)reflect_idx(y - d, n0) * n1 + x] +
                        p[(size_t)reflect_idx(y + d, n0) * n1 + x]) * taps[radius - d];
        } else {
            const double* row = p + (size_t)y * n1;
            acc = row[x] * taps[radius];
            for (int d = radius; d >= 1; --d)
                acc += (row[reflect_idx(x - d, n1)] + row[reflect_idx(x + d, n1)]) * taps[radius - d];
        }
        out[e] = acc;
    }
};

// NI_SplineFilter1D, order 3, in place: boundary == 0: mirror start (modes 'constant',
// 'mirror'), 1: reflect start (modes 'nearest', 'reflect').
LSTED_HD void spline_prefilter_line_mode(double* c, int n, size_t stride, int boundary) {
    if (boundary == 0) { spline_prefilter_line(c, n, stride); return; }
    if (n < 2) return;
    const double z = sqrt(3.0) - 2.0;
    const double gain = (1.0 - z) * (1.0 - 1.0 / z);
    for (int i = 0; i < n; ++i) c[i * stride] *= gain;
    double z_n = 1.0;
    for (int i = 0; i < n; ++i) z_n *= z;
    const double c0 = c[0];
    double s = c[0] + z_n * c[(n - 1) * stride];
    double z_i = z;
    for (int i = 1; i < n; ++i) {
        s += z_i * (c[i * stride] + z_n * c[(n - 1 - i) * stride]);
        z_i *= z;
    }
    c[0] = s * (z / (1.0 - z_n * z_n)) + c0;
    for (int i = 1; i < n; ++i) c[i * stride] += z * c[(i - 1) * stride];
    c[(n - 1) * stride] *= z / (z - 1.0);
    for (int i = n - 2; i >= 0; --i) c[i * stride] = z * (c[(i + 1) * stride] - c[i * stride]);
}

struct ImgPrefilterFn {   // one line per element: e = plane * lines + line
    double* buf;
    int n0, n1, axis, boundary;
    LSTED_HD void operator()(size_t e) const {
        const int lines = axis == 0 ? n1 : n0;
        double* p = buf + (e / lines) * (size_t)n0 * n1;
        const int l = (int)(e % lines);
        if (axis == 0) spline_prefilter_line_mode(p + l, n0, (size_t)n1, boundary);
        else spline_prefilter_line_mode(p + (size_t)l * n1, n1, 1, boundary);
    }
};

// Blocked form of the same prefilter for long lines.  The recursions c+[i] = x[i] + z c+[i-1]
// and c[i] = z (c[i+1] - c+[i]) forget their start as z^k (z = sqrt(3) - 2; 0.268^48 = 4e-28),
// so a thread can produce kPfBlock outputs after a warm-up of kPfWarm samples from a zero state
// instead of walking the whole line: n / kPfBlock threads per line, two passes (causal into a
// second buffer, anti-causal back).  Line starts / ends use the boundary formulas of
// spline_prefilter_line_mode with their sums cut at kPfInit terms (z^64 = 3e-37).
enum { kPfBlock = 32, kPfWarm = 48, kPfInit = 64, kPfMinLine = 160 };
struct PrefilterLines {
    int n;                 // samples per line
    size_t stride;         // between consecutive samples of a line
    size_t nlines;         // lines per plane
    size_t line_step;      // between the first samples of consecutive lines
    size_t plane_step;     // between planes
    size_t offset;         // first sample of line 0 of plane 0
    int nblocks;           // ceil(n / kPfBlock)
    LSTED_HD size_t base(size_t line) const {
        return offset + (line / nlines) * plane_step + (line % nlines) * line_step;
    }
};
struct SplineCausalBlockFn {     // e = block * total_lines + line  (lanes across lines)
    const double* in; double* tmp;
    PrefilterLines L;
    size_t total_lines;
    int boundary;                // 0 mirror, 1 reflect
    LSTED_HD void operator()(size_t e) const {
        const double z = sqrt(3.0) - 2.0;
        const double gain = (1.0 - z) * (1.0 - 1.0 / z);
        const size_t b0 = L.base(e % total_lines);
        const double* x = in + b0;
        double* c = tmp + b0;
        const int s = (int)(e / total_lines) * kPfBlock;
        const int end = s + kPfBlock < L.n ? s + kPfBlock : L.n;
        double state;
        int i;
        if (s <= kPfWarm) {      // from the start of the line: boundary sum (cut at kPfInit terms)
            double sum = 0.0, zi = 1.0;
            for (int k = 0; k < kPfInit && k < L.n; ++k) { sum += zi * (gain * x[(size_t)k * L.stride]); zi *= z; }
            state = boundary == 0 ? sum : gain * x[0] + z * sum;
            if (s == 0) c[0] = state;
            i = 1;
        } else {                 // zero history kPfWarm samples before the block
            i = s - kPfWarm;
            state = gain * x[(size_t)i * L.stride];
            ++i;
        }
        for (; i < end; ++i) {
            state = gain * x[(size_t)i * L.stride] + z * state;
            if (i >= s) c[(size_t)i * L.stride] = state;
        }
    }
};
struct SplineAnticausalBlockFn {
    const double* tmp; double* out;
    PrefilterLines L;
    size_t total_lines;
    int boundary;
    LSTED_HD void operator()(size_t e) const {
        const double z = sqrt(3.0) - 2.0;
        const size_t b0 = L.base(e % total_lines);
        const double* cp = tmp + b0;
        double* c = out + b0;
        const int s = (int)(e / total_lines) * kPfBlock;
        const int last = (s + kPfBlock < L.n ? s + kPfBlock : L.n) - 1;
        const int n = L.n;
        double state;
        int i;
        if (last + kPfWarm >= n - 1) {   // from the end of the line
            state = boundary == 0
                ? (z * cp[(size_t)(n - 2) * L.stride] + cp[(size_t)(n - 1) * L.stride]) * z / (z * z - 1.0)
                : cp[(size_t)(n - 1) * L.stride] * (z / (z - 1.0));
            if (last == n - 1) c[(size_t)(n - 1) * L.stride] = state;
            i = n - 2;
        } else {
            i = last + kPfWarm;
            state = -z * cp[(size_t)i * L.stride];
            --i;
        }
        for (; i >= s; --i) {
            state = z * (state - cp[(size_t)i * L.stride]);
            if (i <= last) c[(size_t)i * L.stride] = state;
        }
    }
};

// Column prefilter (mirror start) of rows [y0, y1) only: the data are zero outside a band well
// inside that range, and the recursion's memory decays as 0.268^k (40 rows: 1e-23)
struct ImgPrefilterRowsFn {
    double* buf;
    int n0, n1, y0, y1;
    LSTED_HD void operator()(size_t e) const {
        double* p = buf + (e / n1) * (size_t)n0 * n1 + (size_t)y0 * n1 + (e % n1);
        spline_prefilter_line(p, y1 - y0, (size_t)n1);
    }
};

// np.pad(plane, npad, 'edge') (scipy's _prepad_for_spline_filter for mode='nearest')
struct ImgEdgePadFn {
    const double* in; double* out;
    int n0, n1, npad;
    LSTED_HD void operator()(size_t e) const {
        const int p0 = n0 + 2 * npad, p1 = n1 + 2 * npad;
        const int x = (int)(e % p1);
        const size_t t = e / p1;
        const int y = (int)(t % p0);
        int sy = y - npad, sx = x - npad;
        sy = sy < 0 ? 0 : (sy >= n0 ? n0 - 1 : sy);
        sx = sx < 0 ? 0 : (sx >= n1 ? n1 - 1 : sx);
        out[e] = in[(t / p0) * (size_t)n0 * n1 + (size_t)sy * n1 + sx];
    }
};

// One interpolated value from spline coefficients co[c0][c1] at (y, x) (already offset by the
// prepad).  'constant': 0 outside [0, len-1], indices mirrored; 'nearest': no coordinate
// mapping, indices clamped (what NI_GeometricTransform does for these two modes).
LSTED_HD double spline_value(const double* co, int c0, int c1, double y, double x, int mode) {
    if (mode == SPLINE_CONSTANT &&
        (y < 0.0 || y > (double)(c0 - 1) || x < 0.0 || x > (double)(c1 - 1)))
        return 0.0;
    double wy[4], wx[4];
    cubic_weights(y, wy);
    cubic_weights(x, wx);
    const double fy = floor(y), fx = floor(x);
    // clamp before the int conversion: 'nearest' coordinates may lie far outside
    const int sy = (int)(fy < -4.0 ? -4.0 : (fy > c0 + 4.0 ? c0 + 4.0 : fy)) - 1;
    const int sx = (int)(fx < -4.0 ? -4.0 : (fx > c1 + 4.0 ? c1 + 4.0 : fx)) - 1;
    int xx[4];
    for (int q = 0; q < 4; ++q) {
        const int i = sx + q;
        xx[q] = mode == SPLINE_CONSTANT ? spline_mirror(i, c1) : (i < 0 ? 0 : (i >= c1 ? c1 - 1 : i));
    }
    double v = 0.0;
    for (int p = 0; p < 4; ++p) {
        const int i = sy + p;
        const int yy = mode == SPLINE_CONSTANT ? spline_mirror(i, c0) : (i < 0 ? 0 : (i >= c0 ? c0 - 1 : i));
        LSTED_DCHECK(yy >= 0 && yy < c0 && xx[0] >= 0 && xx[3] < c1 && xx[0] < c1 && xx[3] >= 0);
        const double* row = co + (size_t)yy * c1;
        for (int q = 0; q < 4; ++q) v += wy[p] * wx[q] * row[xx[q]];
    }
    return v;
}

struct ImgSplineFn {
    const double* coef;     // [batch][c0][c1] prefiltered (c = n + 2*npad)
    const double* xform;    // [batch][6]: m00 m01 m10 m11 offset0 offset1 (output -> input)
    const double* clip_hi;  // [batch] or null: clip to [0, clip_hi[b]]
    double* out;            // [batch][m0][m1]
    int c0, c1, m0, m1, npad, mode;
    LSTED_HD void operator()(size_t e) const {
        const int j = (int)(e % m1);
        const size_t t = e / m1;
        const int i = (int)(t % m0);
        const size_t b = t / m0;
        const double* m = xform + 6 * b;
        const double y = m[0] * i + m[1] * j + m[4] + npad;
        const double x = m[2] * i + m[3] * j + m[5] + npad;
        double v = spline_value(coef + b * (size_t)c0 * c1, c0, c1, y, x, mode);
        if (clip_hi) v = v < 0.0 ? 0.0 : (v > clip_hi[b] ? clip_hi[b] : v);
        out[e] = v;
    }
};

// Deterministic two-level plane reductions: kReduceSegments interleaved partials per plane.
struct PlaneReducePartialFn {
    const double* in; double* partial;   // partial[plane][kReduceSegments]
    size_t plane_elems;
    int is_max;                           // 1: maximum, 0: sum
    // optional rectangle [y0, y1) x [x0, x1) of planes with `pitch` columns (pitch 0: the
    // whole plane): everything outside is known to be zero
    int y0, y1, x0, x1, pitch;
    LSTED_HD void operator()(size_t e) const {
        const size_t b = e / kReduceSegments, s = e % kReduceSegments;
        const double* p = in + b * plane_elems;
        double acc = is_max ? -INFINITY : 0.0;
        if (pitch == 0) {
            for (size_t k = s; k < plane_elems; k += kReduceSegments) {
                const double v = p[k];
                if (is_max) acc = v > acc ? v : acc; else acc += v;
            }
        } else {
            const int w = x1 - x0;
            const size_t n = (size_t)(y1 - y0) * w;
            for (size_t k = s; k < n; k += kReduceSegments) {
                const double v = p[(size_t)(y0 + (int)(k / w)) * pitch + x0 + (int)(k % w)];
                if (is_max) acc = v > acc ? v : acc; else acc += v;
            }
            if (is_max && acc < 0.0 && n < plane_elems) acc = 0.0;   // the zeros outside the rectangle
        }
        partial[e] = acc;
    }
};
struct PlaneReduceFinalFn {
    const double* partial; double* out;   // out[plane * out_stride]
    size_t out_stride;
    int is_max;
    double scale;                          // result multiplied by this (1.1 for clip bounds)
    LSTED_HD void operator()(size_t b) const {
        const double* p = partial + b * kReduceSegments;
        double acc = p[0];
        for (int s = 1; s < kReduceSegments; ++s) acc = is_max ? (p[s] > acc ? p[s] : acc) : acc + p[s];
        out[b * out_stride] = acc * scale;
    }
};

// ---------------------------------------------------------------------------------------------
// The scan itself
// ---------------------------------------------------------------------------------------------
struct ScanGeom {
    int type;
    int n0, n1;            // padded image
    int n_y, n_x, pad;     // unpadded object, padding
    int step, exc_sep;
    int num_pos;
    int pos_y, pos_x;      // positions per axis (point / multipoint), num_pos = pos_y * pos_x
    int base_y, base_x;    // first reconstruction row / column written by position 0
    int spots_y, spots_x;  // multipoint: spots per axis
    int zoom_rows, zoom_top;  // rescan: rows of the scaled image, rows of zero padding above it
    double zoom;              // rescan: input rows per output row, (n0-1)/(zoom_rows-1)
    const double* rot_obj;    // [n0][n1]
    const double* cexc;       // [n0][n1] centred excitation
    const int* pos;           // [num_pos][2]
};

// glow of position p at (y, x): rot_obj * excitation translated by the scan position (:165-166)
LSTED_HD double scan_glow(const ScanGeom& g, int p, int y, int x) {
    const int ey = y - g.pos[2 * p], ex = x - g.pos[2 * p + 1];
    if (ey < 0 || ey >= g.n0 || ex < 0 || ex >= g.n1) return 0.0;
    return g.rot_obj[(size_t)y * g.n1 + x] * g.cexc[(size_t)ey * g.n1 + ex];
}
LSTED_HD double scan_exc(const ScanGeom& g, int p, int y, int x) {
    const int ey = y - g.pos[2 * p], ex = x - g.pos[2 * p + 1];
    if (ey < 0 || ey >= g.n0 || ex < 0 || ex >= g.n1) return 0.0;
    return g.cexc[(size_t)ey * g.n1 + ex];
}
// what the detector blur acts on: the glow itself (multipoint :202) or the descanned glow (:167)
LSTED_HD double scan_src(const ScanGeom& g, int p, int y, int x) {
    if (g.type == SCAN_MULTIPOINT) return scan_glow(g, p, y, x);
    const int oy = y + g.pos[2 * p], ox = x + g.pos[2 * p + 1];
    if (oy < 0 || oy >= g.n0 || ox < 0 || ox >= g.n1) return 0.0;
    return g.rot_obj[(size_t)oy * g.n1 + ox] * g.cexc[(size_t)y * g.n1 + x];
}

// scipy 'reflect' index without the integer division of reflect_idx (|i| is within a few n here)
LSTED_HD int reflect_near(int i, int n) {
    while ((unsigned)i >= (unsigned)n) i = i < 0 ? -1 - i : 2 * n - 1 - i;
    return i;
}

struct ScanBlur0Fn {   // first pass of the detector blur (axis 0) of positions p0 .. p0+count-1
    ScanGeom g;
    int p0, radius;
    const double* taps;
    double* out;        // [count][n0][n1]
    LSTED_HD void operator()(size_t e) const {
        const int x = (int)(e % g.n1);
        const size_t t = e / g.n1;
        const int y = (int)(t % g.n0);
        const int p = p0 + (int)(t / g.n0);
        double acc = scan_src(g, p, y, x) * taps[radius];
        for (int d = radius; d >= 1; --d)
            acc += (scan_src(g, p, reflect_near(y - d, g.n0), x) +
                    scan_src(g, p, reflect_near(y + d, g.n0), x)) * taps[radius - d];
        out[e] = acc;
    }
};

// The excitation of a line / point scan is confined to a band (box) around the optical axis,
// so the descanned glow is, and after the blur its support has only grown by the blur radius.
// When that support stays clear of the image edges (the reflect boundary never folds it) both
// blur passes work on the support alone: `src` = rows / columns where the blur's input can be
// non-zero, `dst` = where its output can be.
struct ScanSupport {
    int src_y0, src_y1, src_x0, src_x1;
    int dst_y0, dst_y1;          // rows after pass 0 (and after pass 1)
    int dst_x0, dst_x1;          // columns after pass 1
};
enum { kBlurStrip = 8 };

// One pass of the detector blur as strips: a thread owns kBlurStrip consecutive outputs along
// the filter axis at one coordinate across it; the lanes of a warp run across (coalesced
// loads).  Every source sample is loaded once for the whole strip and the taps it meets slide
// through registers (taps_ext = the taps with kBlurStrip zeros before and 3 kBlurStrip after).
//   PASS 0 filters along rows; its input is the descanned glow formed on the fly; its output is
//          written TRANSPOSED, mid[pl][column][row], so that
//   PASS 1 filters along columns with the lanes across rows again, and writes out[pl][row][col].
// Support mode (ScanSupport): samples outside the source support are zero and only the
// destination support is computed; pass 1 writes rows [wr_y0, wr_y1) (zeros where the blur
// cannot reach) and never touches the rest of `out`, which was zeroed once.  Frame mode
// (fold != 0): everything is computed, samples beyond the frame come from the reflected index.
template <int PASS> struct ScanBlurStripFn {
    ScanGeom g;
    ScanSupport sp;
    int p0, radius, nstrips, fold;
    int wr_y0, wr_y1;            // PASS 1: rows written
    const double* taps_ext;
    const double* mid_in;        // PASS 1 input
    double* out;                 // PASS 0: mid[count][n1][n0]; PASS 1: [count][n0][n1]
    LSTED_HD void operator()(size_t e) const {
        const int across0 = PASS == 0 ? sp.src_x0 : wr_y0;
        const int across_n = PASS == 0 ? sp.src_x1 - sp.src_x0 : wr_y1 - wr_y0;
        const int along0 = PASS == 0 ? sp.dst_y0 : sp.dst_x0, along1 = PASS == 0 ? sp.dst_y1 : sp.dst_x1;
        const int src0 = PASS == 0 ? sp.src_y0 : sp.src_x0, src1 = PASS == 0 ? sp.src_y1 : sp.src_x1;
        const int n_along = PASS == 0 ? g.n0 : g.n1;
        const int a = across0 + (int)(e % across_n);
        const size_t t = e / across_n;
        const int first_out = along0 + (int)(t % nstrips) * kBlurStrip;
        const size_t pl = t / nstrips;
        const int p = p0 + (int)pl;
        double acc[kBlurStrip];
#pragma unroll
        for (int j = 0; j < kBlurStrip; ++j) acc[j] = 0.0;
        const bool live = PASS == 0 || (a >= sp.dst_y0 && a < sp.dst_y1);
        if (live) {
            // sample u sits at first_out - radius + u and meets output j with tap u - j
            const int first = first_out - radius;
            const int nsamp = kBlurStrip + 2 * radius;
            const double* mid = PASS == 1 ? mid_in + pl * (size_t)g.n0 * g.n1 + a : 0;
            for (int u0 = 0; u0 < nsamp; u0 += kBlurStrip) {
                double w[2 * kBlurStrip - 1];   // taps u0 - (kBlurStrip-1) .. u0 + kBlurStrip - 1
#pragma unroll
                for (int i = 0; i < 2 * kBlurStrip - 1; ++i) w[i] = taps_ext[u0 + 1 + i];
#pragma unroll
                for (int q = 0; q < kBlurStrip; ++q) {
                    int idx = first + u0 + q;
                    if (fold) idx = reflect_near(idx, n_along);
                    double v = 0.0;
                    if (idx >= src0 && idx < src1)
                        v = PASS == 0 ? scan_src(g, p, idx, a) : mid[(size_t)idx * g.n0];
#pragma unroll
                    for (int j = 0; j < kBlurStrip; ++j) acc[j] += v * w[q - j + kBlurStrip - 1];
                }
            }
        }
        LSTED_DCHECK(first_out >= 0 && first_out < n_along && a >= 0 && p < g.num_pos);
        double* o = PASS == 0 ? out + (pl * g.n1 + a) * (size_t)g.n0 + first_out
                              : out + (pl * g.n0 + a) * (size_t)g.n1 + first_out;
#pragma unroll
        for (int j = 0; j < kBlurStrip; ++j)
            if (first_out + j < along1) o[j] = acc[j];
    }
};

struct ScanGlowMaxPartialFn {   // partial maxima of glow (:231)
    ScanGeom g;
    int p0;
    double* partial;             // [count][kReduceSegments]
    // support of the centred excitation (rows / columns; the whole frame when unknown): the
    // glow of position p is zero outside that rectangle shifted by the scan position
    int ey0, ey1, ex0, ex1;
    LSTED_HD void operator()(size_t e) const {
        const int p = p0 + (int)(e / kReduceSegments);
        int y0 = ey0 + g.pos[2 * p], y1 = ey1 + g.pos[2 * p];
        int x0 = ex0 + g.pos[2 * p + 1], x1 = ex1 + g.pos[2 * p + 1];
        y0 = y0 < 0 ? 0 : y0; y1 = y1 > g.n0 ? g.n0 : y1;
        x0 = x0 < 0 ? 0 : x0; x1 = x1 > g.n1 ? g.n1 : x1;
        const int w = x1 - x0;
        const size_t n = y1 > y0 && w > 0 ? (size_t)(y1 - y0) * w : 0;
        double acc = n < (size_t)g.n0 * g.n1 ? 0.0 : -INFINITY;   // zeros outside the rectangle
        for (size_t k = e % kReduceSegments; k < n; k += kReduceSegments) {
            const double v = scan_glow(g, p, y0 + (int)(k / w), x0 + (int)(k % w));
            acc = v > acc ? v : acc;
        }
        partial[e] = acc;
    }
};

struct ScanColSumFn {   // descan line: inst.sum(axis=1), rows added in order (:179-181)
    const double* inst; double* colsum;   // colsum[(p0 + pl)][n1]
    int n0, n1, p0;
    int y0, y1;                           // rows that can be non-zero
    LSTED_HD void operator()(size_t e) const {
        const size_t pl = e / n1;
        const int x = (int)(e % n1);
        const double* p = inst + pl * (size_t)n0 * n1 + x;
        double acc = 0.0;
        for (int y = y0; y < y1; ++y) acc += p[(size_t)y * n1];
        colsum[(size_t)(p0 + pl) * n1 + x] = acc;
    }
};

struct ScanRegionSumFn {   // multipoint: detector region around every spot (:206-217)
    ScanGeom g;
    const double* inst; double* regsum;   // regsum[p][spots_y][spots_x]
    int p0;
    LSTED_HD void operator()(size_t e) const {
        const int kx = (int)(e % g.spots_x);
        const size_t t = e / g.spots_x;
        const int ky = (int)(t % g.spots_y);
        const size_t pl = t / g.spots_y;
        const int p = p0 + (int)pl;
        const int third = g.exc_sep / 3;
        const int y_sp = g.pad + g.pos[2 * p] + ky * g.exc_sep;
        const int x_sp = g.pad + g.pos[2 * p + 1] + kx * g.exc_sep;
        const int ya = y_sp - third < 0 ? 0 : y_sp - third, yb = y_sp + third > g.n0 ? g.n0 : y_sp + third;
        const int xa = x_sp - third < 0 ? 0 : x_sp - third, xb = x_sp + third > g.n1 ? g.n1 : x_sp + third;
        const double* plane = inst + pl * (size_t)g.n0 * g.n1;
        double acc = 0.0;
        for (int y = ya; y < yb; ++y)
            for (int x = xa; x < xb; ++x) acc += plane[(size_t)y * g.n1 + x];
        regsum[((size_t)p * g.spots_y + ky) * g.spots_x + kx] = acc;
    }
};

// rescan line (:218-225): coef = column-prefiltered blurred descanned glow of the chunk;
// out = shift(scale_y(.), s): row y of the output comes from row y - sy of the padded scaled
// image, i.e. from zoom output row o = y - sy - zoom_top, interpolated at o * zoom.
struct ScanRescanFn {
    ScanGeom g;
    const double* coef; double* out;
    int p0;
    int cy0, cy1;   // rows of `coef` that were prefiltered (coefficients are zero elsewhere)
    LSTED_HD void operator()(size_t e) const {
        const int x = (int)(e % g.n1);
        const size_t t = e / g.n1;
        const int y = (int)(t % g.n0);
        const size_t pl = t / g.n0;
        const int p = p0 + (int)pl;
        const int yy = y - g.pos[2 * p], xx = x - g.pos[2 * p + 1];
        double v = 0.0;
        const int o = yy - g.zoom_top;
        if (yy >= 0 && yy < g.n0 && xx >= 0 && xx < g.n1 && o >= 0 && o < g.zoom_rows) {
            const double cc = (double)o * g.zoom;
            if (!(cc < 0.0 || cc > (double)(g.n0 - 1))) {
                double w[4];
                cubic_weights(cc, w);
                const int s = (int)floor(cc) - 1;
                const double* col = coef + pl * (size_t)g.n0 * g.n1 + xx;
                for (int k = 0; k < 4; ++k) {
                    const int r = spline_mirror(s + k, g.n0);
                    if (r >= cy0 && r < cy1) v += w[k] * col[(size_t)r * g.n1];
                }
            }
        }
        out[e] = v < 0.0 ? 0.0 : v;
    }
};

struct ScanCumFn {   // camera integration over the scan, positions in order (:226)
    const double* inst; double* cum_frames; double* cum;   // cum[n0*n1] carried between chunks
    size_t plane;
    int count;
    LSTED_HD void operator()(size_t e) const {
        double acc = cum[e];
        for (int pl = 0; pl < count; ++pl) {
            acc += inst[(size_t)pl * plane + e];
            cum_frames[(size_t)pl * plane + e] = acc;
        }
        cum[e] = acc;
    }
};

struct GatherPlanesFn {   // dst[slot[k]] = src[k] for the planes of a chunk that are kept (slot >= 0)
    const double* src; double* dst;
    const int* slot;       // [count]
    size_t plane;
    LSTED_HD void operator()(size_t e) const {
        const size_t pl = e / plane;
        const int s = slot[pl];
        if (s >= 0) dst[(size_t)s * plane + e % plane] = src[e];
    }
};

// Per-position sums the reconstruction is assembled from.
struct ScanSums {
    const double* colsum;   // descan line  [P][n1]
    const double* total;    // descan point [P]
    const double* regsum;   // multipoint   [P][spots_y][spots_x]
    const double* cum;      // rescan line  [n0][n1] (camera image after the whole scan)
};

// reconstruction[y][x] (padded coordinates) after scan positions 0 .. p_last (:177-229)
LSTED_HD double scan_reconstruction(const ScanGeom& g, const ScanSums& s, int p_last, int y, int x) {
    if (p_last < 0) return 0.0;
    switch (g.type) {
    case SCAN_DESCAN_LINE: {
        if (y < g.base_y) return 0.0;
        const int q = (y - g.base_y) / g.step;
        return q <= p_last && q < g.num_pos ? s.colsum[(size_t)q * g.n1 + x] : 0.0;
    }
    case SCAN_DESCAN_POINT: {
        if (y < g.base_y || x < g.base_x) return 0.0;
        const int qy = (y - g.base_y) / g.step, qx = (x - g.base_x) / g.step;
        if (qy >= g.pos_y || qx >= g.pos_x) return 0.0;
        const int q = qy * g.pos_x + qx;
        return q <= p_last ? s.total[q] : 0.0;
    }
    case SCAN_MULTIPOINT: {
        const int uy = y - g.pad + g.step / 2, ux = x - g.pad + g.step / 2;
        if (uy < 0 || ux < 0) return 0.0;
        const int ky = uy / g.exc_sep, kx = ux / g.exc_sep;
        if (ky >= g.spots_y || kx >= g.spots_x) return 0.0;
        const int q = ((uy % g.exc_sep) / g.step) * g.pos_x + (ux % g.exc_sep) / g.step;
        return q <= p_last ? s.regsum[((size_t)q * g.spots_y + ky) * g.spots_x + kx] : 0.0;
    }
    default:
        return p_last == g.num_pos - 1 ? s.cum[(size_t)y * g.n1 + x] : 0.0;
    }
}

struct ScanReconFn {   // the full padded reconstruction after the last position
    ScanGeom g; ScanSums s;
    double* out;
    LSTED_HD void operator()(size_t e) const {
        out[e] = scan_reconstruction(g, s, g.num_pos - 1, (int)(e / g.n1), (int)(e % g.n1));
    }
};

// Planes of exc / glow of the kept positions, edge-padded for the display rotation (:169-170)
struct ScanPadExcGlowFn {
    ScanGeom g;
    const int* frame_pos;   // [count]
    double* out;            // [count][2][n0+24][n1+24]
    LSTED_HD void operator()(size_t e) const {
        const int p0 = g.n0 + 2 * kSplinePrepad, p1 = g.n1 + 2 * kSplinePrepad;
        const int x = (int)(e % p1);
        size_t t = e / p1;
        const int y = (int)(t % p0);
        t /= p0;
        const int which = (int)(t % 2);
        const int p = frame_pos[t / 2];
        int sy = y - kSplinePrepad, sx = x - kSplinePrepad;
        sy = sy < 0 ? 0 : (sy >= g.n0 ? g.n0 - 1 : sy);
        sx = sx < 0 ? 0 : (sx >= g.n1 ? g.n1 - 1 : sx);
        out[e] = which == 0 ? scan_exc(g, p, sy, sx) : scan_glow(g, p, sy, sx);
    }
};

// The six cropped, display-scaled planes simulate_imaging hands to generate_figure (:262-271)
struct ScanFrameFn {
    ScanGeom g; ScanSums s;
    const int* frame_pos;        // [count] scan position of every frame
    const double* inst_store;    // [count][n0][n1]
    const double* cum_store;     // [count][n0][n1] (rescan line) or null: cum == inst
    const double* rotated;       // [count][2][n_y][n_x] de-rotated exc / glow, or null (rot == 0)
    double disp_max[6];          // exc, glow, inst, cum, new_sig, reconstruction
    double* out;                 // [count][6][n_y][n_x]
    LSTED_HD void operator()(size_t e) const {
        const int x = (int)(e % g.n_x);
        size_t t = e / g.n_x;
        const int y = (int)(t % g.n_y);
        t /= g.n_y;
        const int which = (int)(t % 6);
        const size_t f = t / 6;
        const int p = frame_pos[f];
        const int Y = y + g.pad, X = x + g.pad;
        LSTED_DCHECK(p >= 0 && p < g.num_pos && Y < g.n0 && X < g.n1);
        const size_t pix = (size_t)Y * g.n1 + X, plane = (size_t)g.n0 * g.n1;
        double v;
        switch (which) {
        case 0: v = rotated ? rotated[((f * 2 + 0) * g.n_y + y) * g.n_x + x] : scan_exc(g, p, Y, X); break;
        case 1: v = rotated ? rotated[((f * 2 + 1) * g.n_y + y) * g.n_x + x] : scan_glow(g, p, Y, X); break;
        case 2: v = inst_store[f * plane + pix]; break;
        case 3: v = (cum_store ? cum_store : inst_store)[f * plane + pix]; break;
        case 4: v = scan_reconstruction(g, s, p, Y, X) - scan_reconstruction(g, s, p - 1, Y, X); break;
        default: v = scan_reconstruction(g, s, p, Y, X); break;
        }
        out[e] = v / disp_max[which];
    }
};

// ---------------------------------------------------------------------------------------------
// Orchestration (backend = CUDA stream + device memory, or the CPU replay of the tests)
// ---------------------------------------------------------------------------------------------
struct ScanParams {
    int type;
    int n_y, n_x, pad;
    int step, exc_sep;
    int num_pos;
    double zoom_factor;        // rescan: 1 / (R^2 + 1)
    int blur_radius;           // detector blur taps (gaussian_filter(., psf_sigma), truncate 4)
    int exc_radius;            // excitation taps (sted sigma, truncate 8)
    size_t chunk_bytes;        // device memory the per-position planes of one chunk may take
};

template <class BK> class ScanEngine {
public:
    ScanGeom g;
    ScanParams prm;
    BK& bk;
    std::vector<int> h_pos;
    int* d_pos = nullptr;
    double *d_blur = nullptr, *d_blur_ext = nullptr, *d_exc_taps = nullptr;
    ScanSupport sup;             // blur supports (the whole frame when use_support is false)
    bool use_support = false;
    int ey0 = 0, ey1 = 0, ex0 = 0, ex1 = 0;   // support of the centred excitation
    int cy0 = 0, cy1 = 0;                     // rescan: rows of the blurred image that are prefiltered
    double *d_obj = nullptr, *d_rot = nullptr, *d_cexc = nullptr;
    double *d_colsum = nullptr, *d_total = nullptr, *d_regsum = nullptr, *d_cum = nullptr;
    double *d_partial = nullptr, *d_max = nullptr;   // [chunk][segments], [P][3]: glow, inst, cum
    double *d_a = nullptr, *d_b = nullptr;           // chunk planes
    double *d_inst_store = nullptr, *d_cum_store = nullptr;
    int *d_frame_pos = nullptr, *d_slot = nullptr;
    double *d_rot_xf = nullptr, *d_rot_pad = nullptr;   // object rotation: transform + clip bound, padded plane
    double last_ms = 0.0;                                // device time of the last run() (CUDA events)
    std::vector<int> h_frame_pos;
    int chunk = 1, frames_cap = 0;
    size_t plane;
    bool have_run = false;

    ScanEngine(BK& backend, const ScanParams& p, const int* positions, const double* blur_taps,
               const double* exc_taps) : prm(p), bk(backend) {
        g.type = p.type;
        g.n_y = p.n_y; g.n_x = p.n_x; g.pad = p.pad;
        g.n0 = p.n_y + 2 * p.pad; g.n1 = p.n_x + 2 * p.pad;
        g.step = p.step; g.exc_sep = p.exc_sep > 0 ? p.exc_sep : 1;
        g.num_pos = p.num_pos;
        plane = (size_t)g.n0 * g.n1;
        h_pos.assign(positions, positions + 2 * (size_t)p.num_pos);
        // positions per axis and the first reconstruction row / column
        g.pos_y = g.pos_x = 1;
        if (p.type == SCAN_DESCAN_LINE || p.type == SCAN_RESCAN_LINE) {
            g.pos_y = p.num_pos;
        } else {
            int px = 1;
            while (px < p.num_pos && positions[2 * px] == positions[0]) ++px;
            g.pos_x = px; g.pos_y = p.num_pos / px;
        }
        g.base_y = positions[0] + p.n_y / 2 + p.pad;
        g.base_x = positions[1] + p.n_x / 2 + p.pad;
        g.spots_y = (p.n_y + g.exc_sep - 1) / g.exc_sep;
        g.spots_x = (p.n_x + g.exc_sep - 1) / g.exc_sep;
        g.zoom_rows = (int)nearbyint((double)g.n0 * p.zoom_factor);   // Python round(): half to even
        g.zoom_top = (g.n0 - g.zoom_rows) / 2;
        g.zoom = g.zoom_rows > 1 ? (double)(g.n0 - 1) / (double)(g.zoom_rows - 1) : 1.0;
        const int planes_per_pos = p.type == SCAN_RESCAN_LINE ? 3 : 2;
        size_t c = p.chunk_bytes / (planes_per_pos * plane * sizeof(double));
        chunk = (int)std::max<size_t>(1, std::min<size_t>(c, (size_t)p.num_pos));

        d_pos = bk.template alloc<int>(2 * (size_t)p.num_pos);
        bk.upload(d_pos, positions, sizeof(int) * 2 * (size_t)p.num_pos);
        d_blur = bk.template alloc<double>(2 * p.blur_radius + 1);
        bk.upload(d_blur, blur_taps, sizeof(double) * (2 * p.blur_radius + 1));
        {   // taps with kBlurStrip zeros on either side (strip kernel)
            std::vector<double> ext(2 * p.blur_radius + 1 + 4 * kBlurStrip, 0.0);
            for (int i = 0; i <= 2 * p.blur_radius; ++i) ext[kBlurStrip + i] = blur_taps[i];
            d_blur_ext = bk.template alloc<double>(ext.size());
            bk.upload(d_blur_ext, ext.data(), sizeof(double) * ext.size());
        }
        find_support();
        d_exc_taps = bk.template alloc<double>(2 * p.exc_radius + 1);
        bk.upload(d_exc_taps, exc_taps, sizeof(double) * (2 * p.exc_radius + 1));
        d_obj = bk.template alloc<double>(plane);
        d_rot = bk.template alloc<double>(plane);
        d_cexc = bk.template alloc<double>(plane);
        d_cum = bk.template alloc<double>(plane);
        d_partial = bk.template alloc<double>((size_t)std::max(chunk, 2) * kReduceSegments);
        d_max = bk.template alloc<double>(3 * (size_t)p.num_pos);
        if (p.type == SCAN_DESCAN_LINE) d_colsum = bk.template alloc<double>((size_t)p.num_pos * g.n1);
        if (p.type == SCAN_DESCAN_POINT) d_total = bk.template alloc<double>(p.num_pos);
        if (p.type == SCAN_MULTIPOINT)
            d_regsum = bk.template alloc<double>((size_t)p.num_pos * g.spots_y * g.spots_x);
        d_a = bk.template alloc<double>((size_t)chunk * plane);
        d_b = bk.template alloc<double>((size_t)chunk * plane * (planes_per_pos - 1));
        // support mode writes the blur's destination rectangle only: the rest stays zero
        bk.zero(d_b, sizeof(double) * (size_t)chunk * plane);
        d_slot = bk.template alloc<int>(chunk);
        g.pos = d_pos; g.rot_obj = d_rot; g.cexc = d_cexc;
        make_excitation();
    }
    ~ScanEngine() {
        void* all[] = {d_pos, d_blur, d_blur_ext, d_exc_taps, d_obj, d_rot, d_cexc, d_colsum, d_total, d_regsum,
                       d_cum, d_partial, d_max, d_a, d_b, d_inst_store, d_cum_store, d_frame_pos, d_slot,
                       d_rot_xf, d_rot_pad};
        for (void* p : all) if (p) bk.free(p);
    }

    // Where the descanned glow and its blur can be non-zero (see ScanSupport).  Multipoint
    // excitation covers the frame, and a support that reaches an edge may be folded back by the
    // reflect boundary: both keep the general kernels.
    void find_support() {
        use_support = false;
        const int re = prm.exc_radius, rb = prm.blur_radius;
        ScanSupport full = {0, g.n0, 0, g.n1, 0, g.n0, 0, g.n1};
        sup = full;
        ey0 = 0; ey1 = g.n0; ex0 = 0; ex1 = g.n1;
        cy0 = 0; cy1 = g.n0;
        if (g.type == SCAN_MULTIPOINT) return;
        // excitation: a delta line / point blurred with radius re (reflections stay inside)
        ey0 = std::max(0, g.n0 / 2 - re); ey1 = std::min(g.n0, g.n0 / 2 + re + 1);
        if (g.type == SCAN_DESCAN_POINT) { ex0 = std::max(0, g.n1 / 2 - re); ex1 = std::min(g.n1, g.n1 / 2 + re + 1); }
        ScanSupport t = full;
        t.src_y0 = g.n0 / 2 - re; t.src_y1 = g.n0 / 2 + re + 1;
        t.dst_y0 = t.src_y0 - rb; t.dst_y1 = t.src_y1 + rb;
        if (t.dst_y0 < 0 || t.dst_y1 > g.n0) return;
        if (g.type == SCAN_DESCAN_POINT) {
            t.src_x0 = g.n1 / 2 - re; t.src_x1 = g.n1 / 2 + re + 1;
            t.dst_x0 = t.src_x0 - rb; t.dst_x1 = t.src_x1 + rb;
            if (t.dst_x0 < 0 || t.dst_x1 > g.n1) return;
        }   // a line runs edge to edge: pass 1 covers the full width and reflects at the row ends
        sup = t;
        use_support = true;
        if (g.type == SCAN_RESCAN_LINE) { cy0 = std::max(0, t.dst_y0 - 40); cy1 = std::min(g.n0, t.dst_y1 + 40); }
    }

    // centred excitation (:104-140): delta line / point / spot lattice, blurred by the STED width
    void make_excitation() {
        std::vector<double> e(plane, 0.0);
        if (g.type == SCAN_DESCAN_LINE || g.type == SCAN_RESCAN_LINE) {
            for (int x = 0; x < g.n1; ++x) e[(size_t)(g.n0 / 2) * g.n1 + x] = 1.0;
        } else if (g.type == SCAN_DESCAN_POINT) {
            e[(size_t)(g.n0 / 2) * g.n1 + g.n1 / 2] = 1.0;
        } else {
            for (int y = g.pad; y < g.n0 - g.pad; y += g.exc_sep)
                for (int x = g.pad; x < g.n1 - g.pad; x += g.exc_sep) e[(size_t)y * g.n1 + x] = 1.0;
        }
        bk.upload(d_a, e.data(), sizeof(double) * plane);
        const bool line = g.type == SCAN_DESCAN_LINE || g.type == SCAN_RESCAN_LINE;
        ImgFirFn f;
        f.n0 = g.n0; f.n1 = g.n1; f.radius = prm.exc_radius; f.taps = d_exc_taps;
        f.in = d_a; f.out = line ? d_cexc : d_b; f.axis = 0;
        bk.for_each(plane, f);
        if (!line) {   // sted_sigma = (0, s, s): axis 1 then axis 2
            f.in = d_b; f.out = d_cexc; f.axis = 1;
            bk.for_each(plane, f);
        }
    }
    void get_excitation(double* out) { bk.download(out, d_cexc, sizeof(double) * plane); }

    void plane_reduce(const double* in, size_t planes, size_t elems, int is_max, double* out,
                      size_t out_stride, double scale = 1.0, bool on_support = false) {
        for (size_t b0 = 0; b0 < planes; b0 += (size_t)std::max(chunk, 2)) {
            const size_t nb = std::min<size_t>(std::max(chunk, 2), planes - b0);
            PlaneReducePartialFn a{in + b0 * elems, d_partial, elems, is_max};
            if (on_support && use_support) {
                a.y0 = sup.dst_y0; a.y1 = sup.dst_y1; a.x0 = sup.dst_x0; a.x1 = sup.dst_x1; a.pitch = g.n1;
            }
            bk.for_each(nb * kReduceSegments, a);
            PlaneReduceFinalFn b{d_partial, out + b0 * out_stride, out_stride, is_max, scale};
            bk.for_each(nb, b);
        }
    }

    // cubic-spline prefilter of `planes` planes [c0][c1] along both axes, in place; `tmp` has the
    // same size (used by the blocked passes for lines of kPfMinLine samples and more)
    void prefilter_planes(double* buf, double* tmp, size_t planes, int c0, int c1, int boundary) {
        for (int axis = 0; axis < 2; ++axis) {
            const int n = axis == 0 ? c0 : c1;
            const size_t nlines = axis == 0 ? c1 : c0;
            if (n < kPfMinLine || !tmp) {
                ImgPrefilterFn pf{buf, c0, c1, axis, boundary};
                bk.for_each(planes * nlines, pf);
                continue;
            }
            PrefilterLines L;
            L.n = n; L.stride = axis == 0 ? (size_t)c1 : 1; L.nlines = nlines;
            L.line_step = axis == 0 ? 1 : (size_t)c1; L.plane_step = (size_t)c0 * c1; L.offset = 0;
            L.nblocks = (n + kPfBlock - 1) / kPfBlock;
            SplineCausalBlockFn ca{buf, tmp, L, planes * nlines, boundary};
            bk.for_each(planes * nlines * L.nblocks, ca);
            SplineAnticausalBlockFn an{tmp, buf, L, planes * nlines, boundary};
            bk.for_each(planes * nlines * L.nblocks, an);
        }
    }

    // rotate(obj, rot) (:382-391): scipy.ndimage.rotate, order 3, mode='nearest', clipped to
    // [0, 1.1 max]; xform = the matrix/offset scipy builds for the angle; null: rot == 0.
    void rotate_planes(const double* in, size_t planes, int n0, int n1, const double* d_xform,
                       const double* d_clip, double* padded, double* tmp, double* out, int m0, int m1) {
        const int c0 = n0 + 2 * kSplinePrepad, c1 = n1 + 2 * kSplinePrepad;
        if (in) {
            ImgEdgePadFn pad{in, padded, n0, n1, kSplinePrepad};
            bk.for_each(planes * c0 * c1, pad);
        }
        prefilter_planes(padded, tmp, planes, c0, c1, 1);
        ImgSplineFn sp{padded, d_xform, d_clip, out, c0, c1, m0, m1, kSplinePrepad, SPLINE_NEAREST};
        bk.for_each(planes * m0 * m1, sp);
    }

    // One orientation: the whole scan.  maxima[P][5] = per-position maxima of glow, inst, cum,
    // reconstruction, new_signal (:231-235); reconstruction / cum_detector_sig [n0][n1] after the
    // last position; frame_positions = the scan positions whose planes are kept for frames().
    void run(const double* obj_padded, const double* rot_xform, const int* frame_positions,
             int num_frames, double* maxima, double* reconstruction, double* cum_detector_sig) {
        const int P = g.num_pos;
        // memory first (kept for the next orientation), then the timed device work
        if (rot_xform && !d_rot_pad) {
            const int c0 = g.n0 + 2 * kSplinePrepad, c1 = g.n1 + 2 * kSplinePrepad;
            d_rot_xf = bk.template alloc<double>(8);
            d_rot_pad = bk.template alloc<double>(2 * (size_t)c0 * c1);   // padded plane + prefilter buffer
        }
        h_frame_pos.assign(frame_positions, frame_positions + num_frames);
        std::vector<int> slot_of(P, -1);
        for (int f = 0; f < num_frames; ++f) slot_of[h_frame_pos[f]] = f;
        if (num_frames > frames_cap) {
            if (d_inst_store) bk.free(d_inst_store);
            if (d_cum_store) bk.free(d_cum_store);
            if (d_frame_pos) bk.free(d_frame_pos);
            d_inst_store = bk.template alloc<double>((size_t)num_frames * plane);
            d_cum_store = g.type == SCAN_RESCAN_LINE ? bk.template alloc<double>((size_t)num_frames * plane) : nullptr;
            d_frame_pos = bk.template alloc<int>(num_frames);
            frames_cap = num_frames;
        }
        bk.upload(d_obj, obj_padded, sizeof(double) * plane);
        if (rot_xform) bk.upload(d_rot_xf, rot_xform, sizeof(double) * 6);
        bk.timer_start();
        if (rot_xform) {
            plane_reduce(d_obj, 1, plane, 1, d_rot_xf + 6, 1, 1.1);
            rotate_planes(d_obj, 1, g.n0, g.n1, d_rot_xf, d_rot_xf + 6, d_rot_pad,
                          d_rot_pad + (size_t)(g.n0 + 2 * kSplinePrepad) * (g.n1 + 2 * kSplinePrepad),
                          d_rot, g.n0, g.n1);
        } else {
            bk.copy(d_rot, d_obj, sizeof(double) * plane);
        }
        if (num_frames) bk.upload(d_frame_pos, h_frame_pos.data(), sizeof(int) * num_frames);
        bk.zero(d_cum, sizeof(double) * plane);

        for (int p0 = 0; p0 < P; p0 += chunk) {
            const int cnt = std::min(chunk, P - p0);
            const size_t elems = (size_t)cnt * plane;
            ScanGlowMaxPartialFn gm{g, p0, d_partial, ey0, ey1, ex0, ex1};
            bk.for_each((size_t)cnt * kReduceSegments, gm);
            PlaneReduceFinalFn gf{d_partial, d_max + 3 * (size_t)p0 + 0, 3, 1, 1.0};
            bk.for_each(cnt, gf);
            {   // detector blur: rows (output transposed into d_a), then columns into d_b
                const bool rescan = g.type == SCAN_RESCAN_LINE;
                const int fold = use_support ? 0 : 1;
                ScanBlurStripFn<0> b0;
                b0.g = g; b0.sp = sup; b0.p0 = p0; b0.radius = prm.blur_radius; b0.fold = fold;
                b0.nstrips = (sup.dst_y1 - sup.dst_y0 + kBlurStrip - 1) / kBlurStrip;
                b0.wr_y0 = b0.wr_y1 = 0; b0.taps_ext = d_blur_ext; b0.mid_in = nullptr; b0.out = d_a;
                bk.for_each((size_t)cnt * b0.nstrips * (sup.src_x1 - sup.src_x0), b0);
                ScanBlurStripFn<1> b1;
                b1.g = g; b1.sp = sup; b1.p0 = p0; b1.radius = prm.blur_radius;
                b1.fold = (use_support && g.type == SCAN_DESCAN_POINT) ? 0 : 1;   // lines: reflect at the row ends
                b1.nstrips = (sup.dst_x1 - sup.dst_x0 + kBlurStrip - 1) / kBlurStrip;
                b1.wr_y0 = rescan ? cy0 : sup.dst_y0; b1.wr_y1 = rescan ? cy1 : sup.dst_y1;
                b1.taps_ext = d_blur_ext; b1.mid_in = d_a; b1.out = d_b;
                bk.for_each((size_t)cnt * b1.nstrips * (b1.wr_y1 - b1.wr_y0), b1);
            }
            double* inst = d_b;
            if (g.type == SCAN_RESCAN_LINE) {
                if (cy1 - cy0 >= kPfMinLine) {   // blocked column prefilter, d_a (free here) as second buffer
                    PrefilterLines L;
                    L.n = cy1 - cy0; L.stride = (size_t)g.n1; L.nlines = (size_t)g.n1; L.line_step = 1;
                    L.plane_step = plane; L.offset = (size_t)cy0 * g.n1;
                    L.nblocks = (L.n + kPfBlock - 1) / kPfBlock;
                    SplineCausalBlockFn ca{d_b, d_a, L, (size_t)cnt * g.n1, 0};
                    bk.for_each((size_t)cnt * g.n1 * L.nblocks, ca);
                    SplineAnticausalBlockFn an{d_a, d_b, L, (size_t)cnt * g.n1, 0};
                    bk.for_each((size_t)cnt * g.n1 * L.nblocks, an);
                } else {
                    ImgPrefilterRowsFn pf{d_b, g.n0, g.n1, cy0, cy1};
                    bk.for_each((size_t)cnt * g.n1, pf);
                }
                ScanRescanFn rs{g, d_b, d_a, p0, cy0, cy1};
                bk.for_each(elems, rs);
                inst = d_a;
                double* cumf = d_b + (size_t)chunk * plane;
                ScanCumFn cf{inst, cumf, d_cum, plane, cnt};
                bk.for_each(plane, cf);
                plane_reduce(cumf, cnt, plane, 1, d_max + 3 * (size_t)p0 + 2, 3);
                if (num_frames) {
                    bk.upload(d_slot, slot_of.data() + p0, sizeof(int) * cnt);
                    GatherPlanesFn ga{cumf, d_cum_store, d_slot, plane};
                    bk.for_each(elems, ga);
                }
            } else if (g.type == SCAN_DESCAN_LINE) {
                ScanColSumFn cs{inst, d_colsum, g.n0, g.n1, p0, sup.dst_y0, sup.dst_y1};
                bk.for_each((size_t)cnt * g.n1, cs);
            } else if (g.type == SCAN_DESCAN_POINT) {
                plane_reduce(inst, cnt, plane, 0, d_total + p0, 1, 1.0, true);
            } else {
                ScanRegionSumFn rg{g, inst, d_regsum, p0};
                bk.for_each((size_t)cnt * g.spots_y * g.spots_x, rg);
            }
            plane_reduce(inst, cnt, plane, 1, d_max + 3 * (size_t)p0 + 1, 3, 1.0, g.type != SCAN_RESCAN_LINE);
            if (num_frames) {
                bk.upload(d_slot, slot_of.data() + p0, sizeof(int) * cnt);
                GatherPlanesFn ga{inst, d_inst_store, d_slot, plane};
                bk.for_each(elems, ga);
            }
            bk.sync();   // d_slot / chunk planes are reused by the next chunk
        }
        // results
        ScanReconFn rc{g, sums(), d_a};
        bk.for_each(plane, rc);
        last_ms = bk.timer_stop();
        std::vector<double> mx(3 * (size_t)P);
        bk.download(mx.data(), d_max, sizeof(double) * 3 * P);
        if (reconstruction) bk.download(reconstruction, d_a, sizeof(double) * plane);
        if (cum_detector_sig) {
            if (g.type == SCAN_RESCAN_LINE) bk.download(cum_detector_sig, d_cum, sizeof(double) * plane);
            else bk.download(cum_detector_sig, d_b + (size_t)((P - 1) % chunk) * plane, sizeof(double) * plane);
        }
        if (maxima) position_maxima(mx, maxima);
        have_run = true;
    }

    ScanSums sums() const { return ScanSums{d_colsum, d_total, d_regsum, d_cum}; }

    // maxima of reconstruction / new_signal per position from the small per-position sums
    // (bookkeeping on P x n1 numbers at most, done on the host)
    void position_maxima(const std::vector<double>& mx, double* maxima) {
        const int P = g.num_pos;
        std::vector<double> written(P, 0.0);   // maximum of what position p wrote
        if (g.type == SCAN_DESCAN_LINE) {
            std::vector<double> cs((size_t)P * g.n1);
            bk.download(cs.data(), d_colsum, sizeof(double) * cs.size());
            for (int p = 0; p < P; ++p) {
                const bool inside = g.base_y + p * g.step < g.n0;
                double m = -INFINITY;
                for (int x = 0; x < g.n1; ++x) m = std::max(m, cs[(size_t)p * g.n1 + x]);
                written[p] = inside ? m : -INFINITY;
            }
        } else if (g.type == SCAN_DESCAN_POINT) {
            bk.download(written.data(), d_total, sizeof(double) * P);
            for (int p = 0; p < P; ++p)
                if (g.base_y + (p / g.pos_x) * g.step >= g.n0 || g.base_x + (p % g.pos_x) * g.step >= g.n1)
                    written[p] = -INFINITY;
        } else if (g.type == SCAN_MULTIPOINT) {
            const size_t spots = (size_t)g.spots_y * g.spots_x;
            std::vector<double> rs((size_t)P * spots);
            bk.download(rs.data(), d_regsum, sizeof(double) * rs.size());
            for (int p = 0; p < P; ++p) {
                double m = -INFINITY;
                for (size_t k = 0; k < spots; ++k) m = std::max(m, rs[p * spots + k]);
                written[p] = m;
            }
        }
        double recon_max = 0.0;
        for (int p = 0; p < P; ++p) {
            double* o = maxima + 5 * (size_t)p;
            o[0] = mx[3 * (size_t)p + 0];
            o[1] = mx[3 * (size_t)p + 1];
            if (g.type == SCAN_RESCAN_LINE) {
                o[2] = mx[3 * (size_t)p + 2];
                o[3] = o[4] = p == P - 1 ? std::max(o[2], 0.0) : 0.0;
            } else {
                o[2] = o[1];                                  // cum_detector_sig is inst (:176,:204)
                recon_max = std::max(recon_max, written[p]);
                o[3] = recon_max;
                o[4] = std::max(written[p], 0.0);             // the rest of new_signal is 0
            }
        }
    }

    // Frames first .. first+count-1 of the kept positions: out[count][6][n_y][n_x] =
    // excitation, glow, instantaneous / cumulative detector signal, new signal, reconstruction,
    // cropped and divided by disp_max[6]; inv_xform = scipy's matrix/offset for rotate(., -rot),
    // null when rot == 0.
    void frames(int first, int count, const double* disp_max, const double* inv_xform, double* out) {
        const size_t crop = (size_t)g.n_y * g.n_x;
        double* d_out = bk.template alloc<double>((size_t)count * 6 * crop);
        double* d_rotated = nullptr;
        double *d_padded = nullptr, *d_xf = nullptr, *d_clip = nullptr;
        if (inv_xform) {
            const int c0 = g.n0 + 2 * kSplinePrepad, c1 = g.n1 + 2 * kSplinePrepad;
            const size_t planes = 2 * (size_t)count;
            d_padded = bk.template alloc<double>(2 * planes * c0 * c1);   // + the prefilter's second buffer
            d_rotated = bk.template alloc<double>(planes * crop);
            d_xf = bk.template alloc<double>(6 * planes);
            d_clip = bk.template alloc<double>(planes);
            // the crop window [pad, pad + n) of the rotated plane: fold it into the offset
            std::vector<double> xf(6 * planes);
            for (size_t b = 0; b < planes; ++b) {
                double* m = xf.data() + 6 * b;
                for (int k = 0; k < 4; ++k) m[k] = inv_xform[k];
                m[4] = inv_xform[4] + inv_xform[0] * g.pad + inv_xform[1] * g.pad;
                m[5] = inv_xform[5] + inv_xform[2] * g.pad + inv_xform[3] * g.pad;
            }
            bk.upload(d_xf, xf.data(), sizeof(double) * xf.size());
            ScanPadExcGlowFn pe{g, d_frame_pos + first, d_padded};
            bk.for_each(planes * c0 * c1, pe);
            plane_reduce_any(d_padded, planes, (size_t)c0 * c1, d_clip);
            rotate_planes(nullptr, planes, g.n0, g.n1, d_xf, d_clip, d_padded, d_padded + planes * c0 * c1,
                          d_rotated, g.n_y, g.n_x);
        }
        ScanFrameFn ff;
        ff.g = g; ff.s = sums(); ff.frame_pos = d_frame_pos + first;
        ff.inst_store = d_inst_store + (size_t)first * plane;
        ff.cum_store = d_cum_store ? d_cum_store + (size_t)first * plane : nullptr;
        ff.rotated = d_rotated;
        for (int k = 0; k < 6; ++k) ff.disp_max[k] = disp_max[k];
        ff.out = d_out;
        bk.for_each((size_t)count * 6 * crop, ff);
        bk.download(out, d_out, sizeof(double) * (size_t)count * 6 * crop);
        void* tmp[] = {d_out, d_rotated, d_padded, d_xf, d_clip};
        for (void* p : tmp) if (p) bk.free(p);
    }

    // clip bounds 1.1 * max for any number of planes (own partial buffer: may exceed the chunk)
    void plane_reduce_any(const double* in, size_t planes, size_t elems, double* out) {
        double* part = bk.template alloc<double>(planes * kReduceSegments);
        PlaneReducePartialFn a{in, part, elems, 1};
        bk.for_each(planes * kReduceSegments, a);
        PlaneReduceFinalFn b{part, out, 1, 1, 1.1};
        bk.for_each(planes, b);
        bk.sync();
        bk.free(part);
    }
};

}  // namespace lsted
