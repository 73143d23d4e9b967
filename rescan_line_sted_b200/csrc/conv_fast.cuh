// Fast path of the convolution engine for transform lengths with a
// compile-time three-radix plan (2160 = 16*9*15 first: the 2048^2 / 107^2
// geometry of BASELINE config 4).  Same operators, same HBM layout and same
// argument structs as conv_bodies.cuh; what changes is the inside of a CTA:
//
//  * butterflies of all three passes live in registers (fft_static.cuh),
//    two shared-memory exchanges per transform, no staging pass;
//  * operands of the first pass come straight from global memory (coalesced:
//    consecutive threads own consecutive rows / pixels), results of the last
//    pass go straight back;
//  * the OTF product, the Fourier-domain sum over orientations, the 'same'
//    crop, clip, RL ratio and RL update all happen on registers between an
//    inverse and a forward transform (the crop offset becomes a phase ramp
//    applied while the half spectra are split).
//
// Bodies are written as per-thread phases separated by CTA barriers
// (`cx.phase(regs, f)`), so tests/host_emul can replay them thread by thread.
#pragma once
#include "conv_bodies.cuh"
#include "fft_static.cuh"

namespace lsted {

template <typename T_, int RA, int RB, int RC, int NT_, int C_, int PR_> struct FastPlan {
    typedef T_ T;
    typedef Fft3<T, -1, RA, RB, RC, NT_> Fwd;
    typedef Fft3<T, +1, RC, RB, RA, NT_> Inv;
    enum {
        L = RA * RB * RC, NT = NT_, C = C_, PR = PR_,
        SEQ = imax(Fwd::SEQ, Inv::SEQ),
        // column CTA: C interleaved sequences; the odd-ish offset spreads them over banks
        LSM_COL = (SEQ + 15) / 16 * 16 + 16 / C_,
        COL_THREADS = NT_ * C_,
        // row CTA: PR groups of NTG threads (whole warps), one row pair each
        NTG = (NT_ + 31) / 32 * 32,
        LSM_ROW = (SEQ + 15) / 16 * 16 + 8,
        ROW_THREADS = NTG * PR_,
        VREG = imax(Fwd::VREG, Inv::VREG),
        NKEEP = Fwd::MC * Fwd::RC
    };
    static_assert(Fwd::MC * Fwd::RC == Inv::MA * Inv::RA, "pass C / pass A operand sets must match");
    static_assert(Inv::MC * Inv::RC == Fwd::MA * Fwd::RA, "pass C / pass A operand sets must match");
};

template <class P> struct ColRegs {
    cplx<typename P::T> keep[P::NKEEP];  // forward spectrum (COL_H) or Fourier-domain sum (COL_HT)
    cplx<typename P::T> v[P::VREG];
};
template <class P> struct RowRegs {
    cplx<typename P::T> v[P::VREG];
};

template <class P> LSTED_HD size_t fast_col_smem_bytes() {
    return sizeof(cplx<typename P::T>) * (size_t)P::C * P::LSM_COL;
}
template <class P> LSTED_HD size_t fast_row_smem_bytes() {
    return sizeof(cplx<typename P::T>) * (size_t)P::PR * P::LSM_ROW;
}

// ---------------------------------------------------------------------------
// Column kernels
// ---------------------------------------------------------------------------
// Forward pass-A operands of column c from an XB slab with `rows` valid rows.
template <class P>
LSTED_HD void col_load_fwd_a(cplx<typename P::T>* v, int t, int c, const cplx<typename P::T>* slab,
                             int rows) {
    typedef typename P::Fwd F;
    typedef typename P::T T;
    LSTED_UNROLL
    for (int m = 0; m < F::MA; ++m) {
        const int j = t + m * P::NT;
        if (j < F::NA) {
            LSTED_UNROLL
            for (int q = 0; q < F::RA; ++q) {
                const int y = j + q * F::NA;
                v[m * F::RA + q] = (y < rows) ? slab[(size_t)y * P::C + c] : mk<T>(0, 0);
            }
        }
    }
}

// Inverse pass-C results (logical row y = j + q*NC) -> cropped rows of dst.
template <class P>
LSTED_HD void col_store_inv_c(const cplx<typename P::T>* v, int t, int c, cplx<typename P::T>* slab,
                              int sy, int Ny) {
    typedef typename P::Inv I;
    LSTED_UNROLL
    for (int m = 0; m < I::MC; ++m) {
        const int j = t + m * P::NT;
        if (j < I::NC) {
            LSTED_UNROLL
            for (int q = 0; q < I::RC; ++q) {
                const int y = j + q * I::NC - sy;
                if (y >= 0 && y < Ny) slab[(size_t)y * P::C + c] = v[m * I::RC + q];
            }
        }
    }
}

template <int MODE, class P, class Ctx>
LSTED_HD void col_fast_body(Ctx& cx, int block, const ColArgs<typename P::T>& a,
                            cplx<typename P::T>* smem, ColRegs<P>* regs) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Ly = g.Ly;
    const int xb = block;
    const size_t slab_ly = (size_t)P::C * Ly, img_ly = (size_t)g.nxb * slab_ly;
    const size_t slab_ny = (size_t)P::C * Ny, img_ny = (size_t)g.nxb * slab_ny;
    const cplx<T>* tw = a.tw;

    if (MODE == COL_H) {
        const cplx<T>* src = a.src + (size_t)xb * slab_ny;
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            col_load_fwd_a<P>(r.v, t, c, src, Ny);
            F::pass_a(r.v, t, smem + c * P::LSM_COL);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::load_b(r.v, t, smem + c * P::LSM_COL, tw);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::pass_b(r.v, t, smem + c * P::LSM_COL);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::pass_c(r.v, t, smem + c * P::LSM_COL, tw);
            LSTED_UNROLL
            for (int i = 0; i < P::NKEEP; ++i) r.keep[i] = r.v[i];
        });
        for (int k = 0; k < a.K; ++k) {
            const cplx<T>* otf = a.otf + (size_t)k * img_ly + (size_t)xb * slab_ly;
            cplx<T>* dst = a.dst + (size_t)k * img_ny + (size_t)xb * slab_ny;
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                const int t = tid / P::C, c = tid - t * P::C;
                LSTED_UNROLL
                for (int m = 0; m < F::MC; ++m) {
                    const int j = t + m * P::NT;
                    if (j < F::NC) {
                        LSTED_UNROLL
                        for (int q = 0; q < F::RC; ++q) {
                            const int y = j + q * F::NC;
                            r.v[m * F::RC + q] = r.keep[m * F::RC + q] * otf[(size_t)y * P::C + c];
                        }
                    }
                }
                I::pass_a(r.v, t, smem + c * P::LSM_COL);
            });
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                const int t = tid / P::C, c = tid - t * P::C;
                I::load_b(r.v, t, smem + c * P::LSM_COL, tw);
            });
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                const int t = tid / P::C, c = tid - t * P::C;
                I::pass_b(r.v, t, smem + c * P::LSM_COL);
            });
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                const int t = tid / P::C, c = tid - t * P::C;
                I::pass_c(r.v, t, smem + c * P::LSM_COL, tw);
                col_store_inv_c<P>(r.v, t, c, dst, g.sy, Ny);
            });
        }
        return;
    }
    // COL_HT
    for (int k = 0; k < a.K; ++k) {
        const cplx<T>* src = a.src + (size_t)k * img_ny + (size_t)xb * slab_ny;
        const cplx<T>* otf = a.otf + (size_t)k * img_ly + (size_t)xb * slab_ly;
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            col_load_fwd_a<P>(r.v, t, c, src, Ny);
            F::pass_a(r.v, t, smem + c * P::LSM_COL);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::load_b(r.v, t, smem + c * P::LSM_COL, tw);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::pass_b(r.v, t, smem + c * P::LSM_COL);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            const int t = tid / P::C, c = tid - t * P::C;
            F::pass_c(r.v, t, smem + c * P::LSM_COL, tw);
            LSTED_UNROLL
            for (int m = 0; m < F::MC; ++m) {
                const int j = t + m * P::NT;
                if (j < F::NC) {
                    LSTED_UNROLL
                    for (int q = 0; q < F::RC; ++q) {
                        const int y = j + q * F::NC;
                        const cplx<T> p = r.v[m * F::RC + q] * otf[(size_t)y * P::C + c];
                        r.keep[m * F::RC + q] = (k == 0) ? p : r.keep[m * F::RC + q] + p;
                    }
                }
            }
        });
    }
    cplx<T>* dst = a.dst + (size_t)xb * slab_ny;
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        const int t = tid / P::C, c = tid - t * P::C;
        LSTED_UNROLL
        for (int i = 0; i < P::NKEEP; ++i) r.v[i] = r.keep[i];
        I::pass_a(r.v, t, smem + c * P::LSM_COL);
    });
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        const int t = tid / P::C, c = tid - t * P::C;
        I::load_b(r.v, t, smem + c * P::LSM_COL, tw);
    });
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        const int t = tid / P::C, c = tid - t * P::C;
        I::pass_b(r.v, t, smem + c * P::LSM_COL);
    });
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        const int t = tid / P::C, c = tid - t * P::C;
        I::pass_c(r.v, t, smem + c * P::LSM_COL, tw);
        col_store_inv_c<P>(r.v, t, c, dst, g.sy, Ny);
    });
}

// ---------------------------------------------------------------------------
// Row kernels
// ---------------------------------------------------------------------------
template <int MODE, class P, class Ctx>
LSTED_HD void row_fast_body(Ctx& cx, int block, const RowArgs<typename P::T>& a,
                            cplx<typename P::T>* smem, RowRegs<P>* regs) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Nx = g.Nx;
    const int Lx = P::L, Lxh = P::L / 2 + 1, C = P::C;  // == g.Lx, g.Lxh, g.C (checked at launch)
    const int Py = (Ny + 1) / 2;
    const int bpi = (Py + P::PR - 1) / P::PR;
    const int img = block / bpi;
    const int pair0 = (block - img * bpi) * P::PR;
    const size_t real_off = (size_t)img * Ny * Nx;
    const size_t spec_off = (size_t)img * g.nxb * C * Ny;
    const cplx<T>* tw = a.tw;
    // the data of a pair sit at logical positions shift + pixel during the transforms
    const int shift = (MODE == ROW_FWD) ? 0 : g.sx;

#define LSTED_ROW_IDS                                      \
    const int f = tid / P::NTG, t = tid - f * P::NTG;      \
    const int pair = pair0 + f;                            \
    const bool live = pair < Py && t < P::NT;              \
    const int y = 2 * pair;                                \
    cplx<T>* sm = smem + f * P::LSM_ROW;                   \
    (void)sm; (void)y; (void)live;

    if (MODE == ROW_FWD) {
        const T* src = a.real_in + real_off;
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (!live) return;
            LSTED_UNROLL
            for (int m = 0; m < F::MA; ++m) {
                const int j = t + m * P::NT;
                if (j < F::NA) {
                    LSTED_UNROLL
                    for (int q = 0; q < F::RA; ++q) {
                        const int i = j + q * F::NA;
                        T va = 0, vb = 0;
                        if (i < Nx) {
                            va = src[(size_t)y * Nx + i];
                            if (y + 1 < Ny) vb = src[(size_t)(y + 1) * Nx + i];
                        }
                        r.v[m * F::RA + q] = mk<T>(va, vb);
                    }
                }
            }
            F::pass_a(r.v, t, sm);
        });
    } else {
        // Hermitian unpack straight into the inverse pass-A registers:
        // Z[i] = A[i] + i B[i] (i <= L/2), conj(A[L-i]) + i conj(B[L-i]) otherwise.
        const cplx<T>* src = a.spec_in + spec_off;
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (!live) return;
            LSTED_UNROLL
            for (int m = 0; m < I::MA; ++m) {
                const int j = t + m * P::NT;
                if (j < I::NA) {
                    LSTED_UNROLL
                    for (int q = 0; q < I::RA; ++q) {
                        const int i = j + q * I::NA;
                        const bool upper = 2 * i > Lx;
                        const int k = upper ? Lx - i : i;
                        const size_t ia = ((size_t)(k / C) * Ny + y) * C + (k % C);
                        cplx<T> A = src[ia];
                        cplx<T> B = (y + 1 < Ny) ? src[ia + C] : mk<T>(0, 0);
                        if (k == 0 || 2 * k == Lx) { A.y = 0; B.y = 0; }
                        r.v[m * I::RA + q] = upper ? mk<T>(A.x + B.y, B.x - A.y)
                                                   : mk<T>(A.x - B.y, A.y + B.x);
                    }
                }
            }
            I::pass_a(r.v, t, sm);
        });
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (live) I::load_b(r.v, t, sm, tw);
        });
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (live) I::pass_b(r.v, t, sm);
        });
        // inverse pass C, then the pointwise step on registers: logical
        // position idx = j + q*NC holds pixel idx - sx of rows y (re) and y+1 (im)
        T* out = (MODE == ROW_FINAL) ? a.real_out : a.real_out + real_off;
        T* out2 = (MODE == ROW_INV_SIM) ? a.real_out2 + real_off : 0;
        const T* aux = (MODE == ROW_MID) ? a.aux + real_off : a.aux;
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (!live) return;
            const bool two = y + 1 < Ny;
            // Issue every global load of the pointwise step up front (they overlap
            // pass C); divisions with their slow-path branches come afterwards.
            cplx<T> pa[(MODE == ROW_MID || MODE == ROW_FINAL) ? I::MC * I::RC : 1];
            cplx<T> pe[MODE == ROW_FINAL ? I::MC * I::RC : 1];
            if (MODE == ROW_MID || MODE == ROW_FINAL) {
                LSTED_UNROLL
                for (int m = 0; m < I::MC; ++m) {
                    const int j = t + m * P::NT;
                    LSTED_UNROLL
                    for (int q = 0; q < I::RC; ++q) {
                        const int i = j + q * I::NC - shift;
                        cplx<T> av = mk<T>(1, 1), ev = mk<T>(0, 0);
                        if (j < I::NC && i >= 0 && i < Nx) {
                            const size_t o = (size_t)y * Nx + i;
                            av.x = aux[o];
                            if (two) av.y = aux[o + Nx];
                            if (MODE == ROW_FINAL) {
                                ev.x = out[o];
                                if (two) ev.y = out[o + Nx];
                            }
                        }
                        pa[m * I::RC + q] = av;
                        if (MODE == ROW_FINAL) pe[m * I::RC + q] = ev;
                    }
                }
            }
            I::pass_c(r.v, t, sm, tw);
            LSTED_UNROLL
            for (int m = 0; m < I::MC; ++m) {
                const int j = t + m * P::NT;
                if (j < I::NC) {
                    LSTED_UNROLL
                    for (int q = 0; q < I::RC; ++q) {
                        const int i = j + q * I::NC - shift;
                        cplx<T> z = r.v[m * I::RC + q];
                        cplx<T> w = mk<T>(0, 0);
                        if (i >= 0 && i < Nx) {
                            const size_t o = (size_t)y * Nx + i;
                            if (MODE == ROW_INV_STORE) {
                                if (a.clip) { z.x = clip0(z.x); z.y = clip0(z.y); }
                                out[o] = a.accumulate ? out[o] + z.x : z.x;
                                if (two) out[o + Nx] = a.accumulate ? out[o + Nx] + z.y : z.y;
                            } else if (MODE == ROW_INV_SIM) {
                                out[o] = clip0(z.x);       // noisy image: second loop below
                                if (two) out[o + Nx] = clip0(z.y);
                            } else if (MODE == ROW_MID) {
                                const cplx<T> mv = pa[m * I::RC + q];
                                w.x = fast_div(mv.x, clip0(z.x));
                                if (two) w.y = fast_div(mv.y, clip0(z.y));
                            } else {  // ROW_FINAL
                                const cplx<T> nv = pa[m * I::RC + q], ev = pe[m * I::RC + q];
                                w.x = ev.x * fast_div(clip0(z.x), nv.x);
                                out[o] = w.x;
                                if (two) {
                                    w.y = ev.y * fast_div(clip0(z.y), nv.y);
                                    out[o + Nx] = w.y;
                                }
                            }
                        }
                        r.v[m * I::RC + q] = w;
                    }
                }
            }
        });
        if (MODE == ROW_INV_SIM) {
            // Shot noise.  A rolled loop over the pixels this thread just wrote (one
            // copy of the sampler in the instruction stream, values re-read from L1/L2).
            cx.phase(regs, [&](int tid, RowRegs<P>& r) {
                LSTED_ROW_IDS
                (void)r;
                if (!live) return;
                const int nr = (y + 1 < Ny) ? 2 : 1;
                LSTED_NOUNROLL
                for (int e = 0; e < I::MC * I::RC * 2; ++e) {
                    const int rr = e & 1, mq = e >> 1;
                    const int m = mq / I::RC, q = mq - m * I::RC;
                    const int j = t + m * P::NT;
                    const int i = j + q * I::NC - shift;
                    if (j < I::NC && rr < nr && i >= 0 && i < Nx) {
                        const size_t o = (size_t)(y + rr) * Nx + i;
                        out2[o] = (T)(poisson_sample((double)out[o], a.seed, o, a.img0 + img) + 1e-9);
                    }
                }
            });
        }
        if (MODE == ROW_INV_STORE || MODE == ROW_INV_SIM) return;
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (live) F::pass_a(r.v, t, sm);
        });
    }
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        if (live) F::load_b(r.v, t, sm, tw);
    });
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        if (live) F::pass_b(r.v, t, sm);
    });
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        if (live) F::pass_c(r.v, t, sm, tw);
    });
    // natural-order spectrum of (a + i b) to shared memory ...
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        if (!live) return;
        LSTED_UNROLL
        for (int m = 0; m < F::MC; ++m) {
            const int j = t + m * P::NT;
            if (j < F::NC) {
                LSTED_UNROLL
                for (int q = 0; q < F::RC; ++q) sm[j + q * F::NC] = r.v[m * F::RC + q];
            }
        }
    });
    // ... then Hermitian split, crop-offset phase ramp, XB store.
    cplx<T>* dst = a.spec_out + spec_off;
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        const int f = tid / P::NTG, tl = tid - f * P::NTG;
        const int pair = pair0 + f;
        if (pair >= Py) return;
        const int y = 2 * pair;
        const int nr = (y + 1 < Ny) ? 2 : 1;
        const cplx<T>* sm = smem + f * P::LSM_ROW;
        const int per_xb = 2 * C;
        for (int w = tl; w < g.nxb * per_xb; w += P::NTG) {
            const int xb = w / per_xb, rem = w - xb * per_xb;
            const int rr = rem / C, c = rem - rr * C;
            if (rr >= nr) continue;
            const int k = xb * C + c;
            cplx<T> o = mk<T>(0, 0);
            if (k < Lxh) {
                const cplx<T> z1 = sm[k];
                const cplx<T> z2 = sm[k == 0 ? 0 : Lx - k];
                if (rr) o = mk<T>((T)0.5 * (z1.y + z2.y), (T)0.5 * (z2.x - z1.x));
                else    o = mk<T>((T)0.5 * (z1.x + z2.x), (T)0.5 * (z1.y - z2.y));
                if (shift) o = o * conj(tw[(k * shift) % Lx]);  // k*shift < L*L/2 fits an int
            }
            dst[((size_t)xb * Ny + y + rr) * C + c] = o;
        }
        (void)r;
    });
#undef LSTED_ROW_IDS
}

}  // namespace lsted
