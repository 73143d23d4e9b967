// Kernel bodies of the 2-D linear ('same') FFT convolution engine and the
// fused Richardson-Lucy passes, written against an abstract CTA context
// (see fft_core.cuh): `cx.parallel_for(n, f)` runs f(0..n-1) across the
// CTA's threads followed by a barrier.  conv_kernels.cu instantiates them
// as sm_100a kernels; tests/host_emul replays them on the CPU.
//
// Operators replaced (reference figure_generation/line_sted_tools.py):
//   H    :567-577   x -> [clip(fftconvolve(x, psf_k, 'same'))]_k
//   H_t  :579-594   y -> sum_k clip(fftconvolve(y_k, psf_k, 'same')) / norm
//   iterate :520-531, create_data_from_object :496-512 (with Poisson :510)
//
// Geometry.  A linear convolution of an Ny x Nx image with an ny x nx PSF,
// cropped like scipy's mode='same' (offset s = (n-1)//2), is alias-free on
// the crop when the circular length is L >= N + max(s, n-1-s).  Rows are
// transformed real-to-complex (two real rows packed into one complex FFT of
// length Lx), columns complex-to-complex (length Ly) on the Lx/2+1
// non-redundant columns.
//
// Spectrum layout in HBM: columns are grouped in blocks of C; block xb holds
// rows x C complex values contiguously, OTFs as [y][c] ("XB"), row spectra as
// [y/2][c][y%2] ("XB2", see xb2_index), so a column CTA streams one fully
// contiguous slab, and a row CTA that owns 2*PR adjacent rows moves
// 2*PR*C*sizeof(complex) contiguous bytes per block.
#pragma once
#include "fft_core.cuh"
#include "poisson.cuh"

namespace lsted {

struct ConvGeom {
    int Ny, Nx;    // image rows / columns
    int Ly, Lx;    // column / row FFT lengths
    int Lxh;       // Lx/2 + 1 non-redundant columns
    int nxb;       // number of column blocks = ceil(Lxh / C)
    int C;         // columns per block (and per column CTA)
    int PR;        // row pairs per row CTA
    int sy, sx;    // 'same' crop offsets
    int Lpx, Lpy;  // shared-memory pitch of one row / column sequence
    FftPlan px, py;
};

LSTED_HD size_t xb_index(int y, int x, int rows, int C) {   // OTF arrays: [xb][y][c]
    return ((size_t)(x / C) * rows + y) * C + (x % C);
}
// Row-spectrum arrays ("XB2"): inside a column block the two rows of a row pair sit
// side by side per column, [xb][y/2][c][y%2], so a row thread moves (A[k], B[k]) of its
// pair with ONE 16-byte access (half the L1 wavefronts of two 8-byte ones 32 bytes
// apart) while a column CTA still streams a fully contiguous slab.  Images with an odd
// number of rows are stored as if they had one more (never read).
LSTED_HD int even_rows(int rows) { return (rows + 1) & ~1; }
// Position of cropped output i inside a transform of length L.  Inside overlap-save tiles the
// window is as long as the transform (N == L, crop offset s > 0): the last s outputs wrap
// around -- they lie outside the alias-free interior and are discarded by the caller, but the
// read must stay inside the sequence.
LSTED_HD int crop_pos(int s, int i, int L) {
    const int p = s + i;
    LSTED_DCHECK(p >= 0 && p < 2 * L);
    return p >= L ? p - L : p;
}
LSTED_HD size_t slab2_index(int y, int c, int C) { return (size_t)(y & ~1) * C + 2 * c + (y & 1); }
LSTED_HD size_t xb2_index(int xb, int y, int c, int rows, int C) {
    return (size_t)xb * even_rows(rows) * C + slab2_index(y, c, C);
}

enum RowMode {
    ROW_FWD = 0,       // real rows -> row spectra
    ROW_INV_STORE = 1, // row spectra -> real rows (optional clip)
    ROW_INV_SIM = 2,   // ... -> clipped noiseless image + Poisson-noisy image
    ROW_MID = 3,       // ... -> ratio = measurement / clip(expected) -> row spectra
    ROW_FINAL = 4      // ... -> estimate *= clip(.) / norm -> row spectra of estimate
};

template <typename T> struct RowArgs {
    ConvGeom g;
    const cplx<T>* tw;       // exp(-2 pi i m / Lx)
    int nimg;
    const T* real_in;        // ROW_FWD: [nimg][Ny][Nx]
    const cplx<T>* spec_in;  // [nimg] XB(rows = Ny)
    cplx<T>* spec_out;       // [nimg] XB(rows = Ny)
    T* real_out;             // STORE: out; SIM: noiseless; FINAL: estimate (in/out, one image)
    T* real_out2;            // SIM: noisy
    const T* aux;            // MID: measurement [nimg][Ny][Nx]; FINAL: normalisation [Ny][Nx]
    int clip;                // STORE: clip negatives to zero
    int accumulate;          // STORE: real_out += result (used by the exact-clip H_t)
    unsigned long long seed; // SIM
    unsigned int img0;       // SIM: global index of image 0 (RNG stream id)
    int prefetch_ahead;      // fast path: L2-prefetch the operands of CTA (blockIdx + this); 0 = off
    const void* tmap_in;     // fast path, fp32: tensor maps (CUtensorMap in global memory) of spec_in /
    const void* tmap_out;    // spec_out; when set the pair's spectrum chunks travel by TMA (fft_core.cuh)
};

// clip(v, 0): NaN clips to 0 on the device (FMNMX) and in the CPU replay alike
template <typename T> LSTED_HD T clip0(T v) { return v > (T)0 ? v : (T)0; }
#ifdef __CUDA_ARCH__
LSTED_HD float clip0(float v) { return fmaxf(v, 0.f); }   // one FMNMX instead of FSETP + FSEL
#endif
// RL ratio measurement / clip(expected) (line_sted_tools.py:527-528).  Deliberate deviation, pinned by
// tests: where the expected image is not positive -- FFT round-off clipped to zero in dark regions,
// which fp32 reaches on any zero-background object -- the reference divides by zero (inf, and NaN in
// every pixel after the next transform); here such a pixel contributes a ratio of 0 and the
// estimate stays finite.  `FAST`: MUFU.RCP based division (<= 2 ulp), fp32 fast path.
template <bool FAST, typename T> LSTED_HD T rl_ratio(T meas, T expected) {
    const T e = clip0(expected);
    const T q = FAST ? fast_div(meas, e) : meas / e;
    return e > (T)0 ? q : (T)0;
}

// Row-pair kernel body.  smem: 2 * PR * Lpx complex.
template <int MODE, typename T, class Ctx>
LSTED_HD void row_body(Ctx& cx, int block, const RowArgs<T>& a, cplx<T>* smem) {
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Nx = g.Nx, Lx = g.Lx, Lxh = g.Lxh, Lp = g.Lpx, C = g.C;
    const int Py = (Ny + 1) / 2;
    const int bpi = (Py + g.PR - 1) / g.PR;
    const int img = block / bpi;
    const int pair0 = (block - img * bpi) * g.PR;
    const int npair = (Py - pair0) < g.PR ? (Py - pair0) : g.PR;
    const int row0 = 2 * pair0;
    const int nrow = (Ny - row0) < 2 * npair ? (Ny - row0) : 2 * npair;
    cplx<T>* b0 = smem;
    cplx<T>* b1 = smem + (size_t)g.PR * Lp;
    const size_t real_off = (size_t)img * Ny * Nx;
    const size_t spec_off = (size_t)img * g.nxb * C * even_rows(Ny);
    cplx<T>* fwd_in;    // packed rows (a + i b) ready for the forward transform
    cplx<T>* fwd_free;  // scratch for it

    if (MODE == ROW_FWD) {
        const T* src = a.real_in + real_off;
        cx.parallel_for(npair * Lx, [&](int w) {
            const int f = w / Lx, i = w - f * Lx, y = row0 + 2 * f;
            T va = 0, vb = 0;
            if (i < Nx) {
                va = src[(size_t)y * Nx + i];
                if (y + 1 < Ny) vb = src[(size_t)(y + 1) * Nx + i];
            }
            b0[f * Lp + pad<T>(i)] = mk<T>(va, vb);
        });
        fwd_in = b0;
        fwd_free = b1;
    } else {
        // Hermitian unpack of the two half spectra A (row y) and B (row y+1)
        // into Z = A + iB over all Lx bins.
        const cplx<T>* src = a.spec_in + spec_off;
        const int per_xb = npair * C;
        cx.parallel_for(g.nxb * per_xb, [&](int w) {
            const int xb = w / per_xb, r = w - xb * per_xb;
            const int f = r / C, c = r - f * C;
            const int k = xb * C + c;
            if (k >= Lxh) return;
            const int y = row0 + 2 * f;
            const size_t ia = xb2_index(xb, y, c, Ny, C);
            cplx<T> A = src[ia];
            cplx<T> B = (y + 1 < Ny) ? src[ia + 1] : mk<T>(0, 0);
            const bool self = (k == 0) || (2 * k == Lx);
            if (self) { A.y = 0; B.y = 0; }
            b0[f * Lp + pad<T>(k)] = mk<T>(A.x - B.y, A.y + B.x);
            if (!self) b0[f * Lp + pad<T>(Lx - k)] = mk<T>(A.x + B.y, B.x - A.y);
        });
        SmemSrc<T> s0 = {b0, Lp};
        cplx<T>* z = fft_batch<+1, T>(cx, g.px, a.tw, s0, b1, b0, npair, Lp);
        cplx<T>* other = (z == b0) ? b1 : b0;
        if (MODE == ROW_INV_STORE || MODE == ROW_INV_SIM) {
            T* out = a.real_out + real_off;
            T* out2 = (MODE == ROW_INV_SIM) ? a.real_out2 + real_off : 0;
            cx.parallel_for(nrow * Nx, [&](int w) {
                const int r = w / Nx, i = w - r * Nx;
                const cplx<T> v = z[(r >> 1) * Lp + pad<T>(crop_pos(g.sx, i, Lx))];
                T val = (r & 1) ? v.y : v.x;
                const size_t o = (size_t)(row0 + r) * Nx + i;
                if (MODE == ROW_INV_SIM) {
                    val = clip0(val);
                    out[o] = val;
                    const unsigned long long pix = (unsigned long long)(row0 + r) * Nx + i;
                    out2[o] = (T)(poisson_sample((double)val, a.seed, pix, a.img0 + img) + 1e-9);
                } else {
                    if (a.clip) val = clip0(val);
                    out[o] = a.accumulate ? out[o] + val : val;
                }
            });
            return;
        }
        // ROW_MID / ROW_FINAL: pointwise step, then re-pack for the forward FFT
        const T* aux = a.aux + (MODE == ROW_MID ? real_off : 0);
        T* est = a.real_out;
        cx.parallel_for(npair * Lx, [&](int w) {
            const int f = w / Lx, i = w - f * Lx, y = row0 + 2 * f;
            T ra = 0, rb = 0;
            if (i < Nx) {
                const cplx<T> v = z[f * Lp + pad<T>(crop_pos(g.sx, i, Lx))];
                const size_t o = (size_t)y * Nx + i;
                if (MODE == ROW_MID) {
                    ra = rl_ratio<false>(aux[o], v.x);
                    if (y + 1 < Ny) rb = rl_ratio<false>(aux[o + Nx], v.y);
                } else {
                    ra = est[o] * (clip0(v.x) / aux[o]);
                    est[o] = ra;
                    if (y + 1 < Ny) {
                        rb = est[o + Nx] * (clip0(v.y) / aux[o + Nx]);
                        est[o + Nx] = rb;
                    }
                }
            }
            other[f * Lp + pad<T>(i)] = mk<T>(ra, rb);
        });
        fwd_in = other;
        fwd_free = z;
    }
    // Forward transform of the packed rows, split into the two half spectra.
    SmemSrc<T> s1 = {fwd_in, Lp};
    const cplx<T>* Z = fft_batch<-1, T>(cx, g.px, a.tw, s1, fwd_free, fwd_in, npair, Lp);
    cplx<T>* dst = a.spec_out + spec_off;
    const int per_xb = nrow * C;
    cx.parallel_for(g.nxb * per_xb, [&](int w) {
        const int xb = w / per_xb, r = w - xb * per_xb;
        const int rr = r / C, c = r - rr * C;
        const int k = xb * C + c;
        cplx<T> o = mk<T>(0, 0);
        if (k < Lxh) {
            const int f = rr >> 1;
            const cplx<T> z1 = Z[f * Lp + pad<T>(k)];
            const cplx<T> z2 = Z[f * Lp + pad<T>(k == 0 ? 0 : Lx - k)];
            if (rr & 1) o = mk<T>((T)0.5 * (z1.y + z2.y), (T)0.5 * (z2.x - z1.x));  // (z1 - conj z2)/(2i)
            else        o = mk<T>((T)0.5 * (z1.x + z2.x), (T)0.5 * (z1.y - z2.y));  // (z1 + conj z2)/2
        }
        dst[xb2_index(xb, row0 + rr, c, Ny, C)] = o;
    });
}

enum { kMaxPeers = 8 };
enum ColMode {
    COL_OTF = 0,  // zero-padded forward transform, scaled, all Ly rows stored
    COL_H = 1,    // one input spectrum -> K products -> K inverse transforms
    COL_HT = 2,   // K input spectra -> sum_k product -> one inverse transform
    COL_LOGMAG = 3  // forward transform -> log(1 + |.|), fftshift-ed, full plane (record_iteration :539-546)
};

template <typename T> struct ColArgs {
    ConvGeom g;
    const cplx<T>* tw;      // exp(-2 pi i m / Ly)
    const cplx<T>* src;     // OTF: [K] XB(rows_in); H: XB(Ny); HT: [K] XB(Ny)
    const cplx<T>* otf;     // [K] XB(rows = Ly)
    const T* otf_real;      // same layout, real part only: set when every PSF is point-symmetric
                            // and the OTFs are stored centred (then they are real; sy = sx = 0)
    cplx<T>* dst;           // OTF: [K] XB(Ly); H: [K] XB(Ny); HT: XB(Ny)
    int K;
    int k_split;            // H (generic kernel): > 1 = the K orientations of a column block are spread
                            // over k_split CTAs (grid nxb * k_split, block = part * nxb + xb), each
                            // repeating the forward transform: small frames are latency-, not work-bound
    int rows_in;            // valid input rows (zero padded up to Ly)
    int src_same;           // HT: every k reads the same input spectrum (H_t of all-ones images)
    T scale;                // OTF: 1/(Lx*Ly)
    double* logmag;         // LOGMAG: [Ny][Nx] out (the transform lengths equal the image size)
    // HT of the fast path with orientations sharded over GPUs: the cross-GPU sum is done inside
    // the kernel over NVLink peer memory (see col_fast_body).  world <= 1: off.
    int p2p_world, p2p_rank;
    unsigned p2p_epoch;               // number of this fused reduction (1, 2, ...)
    cplx<T>* p2p_recv[kMaxPeers];     // per rank: [world][nxb][Ly*C] partial Fourier-domain sums
    unsigned* p2p_flags[kMaxPeers];   // per rank: [world][nxb] "partial of (src, block) landed" epochs
    cplx<T>* p2p_spec[kMaxPeers];     // per rank: the output spectrum (XB2, Ny rows)
    unsigned* p2p_done[kMaxPeers];    // per rank: count of finished blocks
};

// Column-block kernel body.  smem: 3 * C * Lpy complex.
// Grid: nxb blocks (COL_OTF: nxb * K, block = k * nxb + xb).
template <int MODE, typename T, class Ctx>
LSTED_HD void col_body(Ctx& cx, int block, const ColArgs<T>& a, cplx<T>* smem) {
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Ly = g.Ly, Lp = g.Lpy, C = g.C;
    const bool split = MODE == COL_H && a.k_split > 1;
    const int xb = (MODE == COL_OTF || split) ? block % g.nxb : block;
    const int kk = (MODE == COL_OTF || split) ? block / g.nxb : 0;
    cplx<T>* b0 = smem;
    cplx<T>* b1 = smem + (size_t)C * Lp;
    cplx<T>* b2 = smem + (size_t)2 * C * Lp;
    const size_t slab_ly = (size_t)C * Ly;          // one OTF slab
    const size_t img_ly = (size_t)g.nxb * slab_ly;  // one OTF image
    const size_t slab_ny = (size_t)C * even_rows(Ny);   // row spectra: XB2 layout
    const size_t img_ny = (size_t)g.nxb * slab_ny;

    if (MODE == COL_OTF) {
        const int rows = a.rows_in;
        const cplx<T>* src = a.src + ((size_t)kk * g.nxb + xb) * C * even_rows(rows);
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            b0[c * Lp + pad<T>(y)] = (y < rows) ? src[slab2_index(y, c, C)] : mk<T>(0, 0);
        });
        SmemSrc<T> s0 = {b0, Lp};
        const cplx<T>* z = fft_batch<-1, T>(cx, g.py, a.tw, s0, b1, b0, C, Lp);
        cplx<T>* dst = a.dst + (size_t)kk * img_ly + (size_t)xb * slab_ly;
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            dst[w] = scale(z[c * Lp + pad<T>(y)], a.scale);
        });
        return;
    }
    if (MODE == COL_LOGMAG) {
        // log(1 + |fftshift(fft2(x))|) of a real image: the Lx/2+1 stored columns give the
        // rest by Hermitian symmetry, |F[ky][Nx-kx]| = |F[-ky][kx]|; fftshift moves k to
        // (k + N/2) mod N
        const int Nx = g.Nx;
        const cplx<T>* src = a.src + (size_t)xb * slab_ny;
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            b0[c * Lp + pad<T>(y)] = src[slab2_index(y, c, C)];
        });
        SmemSrc<T> s0 = {b0, Lp};
        const cplx<T>* z = fft_batch<-1, T>(cx, g.py, a.tw, s0, b1, b0, C, Lp);
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            const int x = xb * C + c;
            if (x >= g.Lxh) return;
            const cplx<T> v = z[c * Lp + pad<T>(y)];
            const double mag = log(1.0 + sqrt((double)v.x * (double)v.x + (double)v.y * (double)v.y));
            int ys = y + Ny / 2; if (ys >= Ny) ys -= Ny;
            int xs = x + Nx / 2; if (xs >= Nx) xs -= Nx;
            a.logmag[(size_t)ys * Nx + xs] = mag;
            if (x > 0 && 2 * x != Nx) {
                int ym = (y == 0 ? 0 : Ny - y) + Ny / 2; if (ym >= Ny) ym -= Ny;
                int xm = (Nx - x) + Nx / 2; if (xm >= Nx) xm -= Nx;
                a.logmag[(size_t)ym * Nx + xm] = mag;
            }
        });
        return;
    }
    if (MODE == COL_H) {
        const cplx<T>* src = a.src + (size_t)xb * slab_ny;
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            b0[c * Lp + pad<T>(y)] = (y < Ny) ? src[slab2_index(y, c, C)] : mk<T>(0, 0);
        });
        SmemSrc<T> s0 = {b0, Lp};
        cplx<T>* A = fft_batch<-1, T>(cx, g.py, a.tw, s0, b1, b0, C, Lp);
        cplx<T>* f1 = (A == b0) ? b1 : b0;
        const int k_lo = split ? (int)((long long)a.K * kk / a.k_split) : 0;
        const int k_hi = split ? (int)((long long)a.K * (kk + 1) / a.k_split) : a.K;
        for (int k = k_lo; k < k_hi; ++k) {
            const cplx<T>* otf = a.otf + (size_t)k * img_ly + (size_t)xb * slab_ly;
            cx.parallel_for(C * Ly, [&](int w) {
                const int y = w / C, c = w - y * C;
                const int s = c * Lp + pad<T>(y);
                f1[s] = A[s] * otf[w];
            });
            SmemSrc<T> s1 = {f1, Lp};
            const cplx<T>* z = fft_batch<+1, T>(cx, g.py, a.tw, s1, b2, f1, C, Lp);
            cplx<T>* dst = a.dst + (size_t)k * img_ny + (size_t)xb * slab_ny;
            cx.parallel_for(C * Ny, [&](int w) {
                const int y = w / C, c = w - y * C;
                dst[slab2_index(y, c, C)] = z[c * Lp + pad<T>(crop_pos(g.sy, y, Ly))];
            });
        }
        return;
    }
    // COL_HT: accumulate the products in the Fourier domain (b2).
    for (int k = 0; k < a.K; ++k) {
        const cplx<T>* src = a.src + (a.src_same ? 0 : (size_t)k * img_ny) + (size_t)xb * slab_ny;
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            b0[c * Lp + pad<T>(y)] = (y < Ny) ? src[slab2_index(y, c, C)] : mk<T>(0, 0);
        });
        SmemSrc<T> s0 = {b0, Lp};
        const cplx<T>* z = fft_batch<-1, T>(cx, g.py, a.tw, s0, b1, b0, C, Lp);
        const cplx<T>* otf = a.otf + (size_t)k * img_ly + (size_t)xb * slab_ly;
        cx.parallel_for(C * Ly, [&](int w) {
            const int y = w / C, c = w - y * C;
            const int s = c * Lp + pad<T>(y);
            const cplx<T> p = z[s] * otf[w];
            b2[s] = (k == 0) ? p : b2[s] + p;
        });
    }
    SmemSrc<T> s2 = {b2, Lp};
    const cplx<T>* z = fft_batch<+1, T>(cx, g.py, a.tw, s2, b0, b1, C, Lp);
    cplx<T>* dst = a.dst + (size_t)xb * slab_ny;
    cx.parallel_for(C * Ny, [&](int w) {
        const int y = w / C, c = w - y * C;
        dst[slab2_index(y, c, C)] = z[c * Lp + pad<T>(crop_pos(g.sy, y, Ly))];
    });
}

// Centring of the OTFs of point-symmetric PSFs.  psf[c + a] == psf[c - a] about the centre
// pixel c = ((ny-1)/2, (nx-1)/2) makes OTF(f) * exp(+2 pi i (fy cy / Ly + fx cx / Lx)) REAL: the
// phase ramp is exactly the 'same' crop offset (s = c), so with the centred OTF the convolution
// lands at offset 0 (geometry sy = sx = 0, no phase ramp in the row kernels) and only the real
// part has to be stored and streamed (half the OTF bytes of the column kernels).
template <typename T> struct OtfCenterArgs {
    cplx<T>* otf;         // in/out: [K][nxb][Ly][C]
    T* otf_real;          // out: real part, [K][nxb][C/CS][Ly][CS] (CS = C: same indexing)
    const cplx<T>* tw_y;  // exp(-2 pi i m / Ly)
    const cplx<T>* tw_x;  // exp(-2 pi i m / Lx)
    int nxb, Ly, Lx, C, cy, cx;
    int CS;               // columns per sub-block of the real array (set by the backend)
    size_t n;             // K * nxb * Ly * C
};
template <typename T> LSTED_HD void otf_center_apply(const OtfCenterArgs<T>& a, size_t i) {
    const int c = (int)(i % a.C);
    size_t r = i / a.C;
    const int y = (int)(r % a.Ly);
    r /= a.Ly;
    const int x = (int)(r % a.nxb) * a.C + c;
    const cplx<T> ph = conj(a.tw_y[((size_t)y * a.cy) % a.Ly] * a.tw_x[((size_t)x * a.cx) % a.Lx]);
    const cplx<T> v = a.otf[i] * ph;
    a.otf[i] = v;
    // real array: each sub-block of CS columns contiguous (one bulk copy per sub-block CTA)
    const size_t blk = i / ((size_t)a.Ly * a.C);
    a.otf_real[((blk * (a.C / a.CS) + c / a.CS) * a.Ly + y) * a.CS + c % a.CS] = v.x;
}

}  // namespace lsted
