// Fast path of the convolution engine for transform lengths with a
// compile-time three-radix plan (2160 = 16*9*15 first: the 2048^2 / 107^2
// geometry of BASELINE config 4).  Same operators, same HBM layout and same
// argument structs as conv_bodies.cuh; what changes is the inside of a CTA:
//
//  * butterflies of all three passes live in registers (fft_static.cuh); a
//    transform costs two shared-memory exchanges through two ping-pong
//    buffers, i.e. two CTA barriers, and no staging pass;
//  * operands of the first pass come straight from global memory (coalesced:
//    consecutive threads own consecutive rows / pixels), results of the last
//    pass go straight back;
//  * the OTF product, the Fourier-domain sum over orientations, the 'same'
//    crop, clip, RL ratio and RL update all happen on registers between an
//    inverse and a forward transform, inside one barrier-to-barrier phase
//    (the crop offset becomes a phase ramp applied with the Hermitian split);
//  * the Hermitian split of a packed row pair exchanges only the upper half
//    of the spectrum with the "mirror" thread (NC - t).
//
// Bodies are written as per-thread phases separated by CTA barriers
// (`cx.phase(regs, f)`), so tests/host_emul can replay them thread by thread.
#pragma once
#include <type_traits>
#include "conv_bodies.cuh"
#include "fft_static.cuh"

namespace lsted {

// CS_: columns per column CTA when the OTFs are real ("sub-block": C_/CS_ CTAs share one
// C_-column block of the HBM layout; fp32: 2 of 4, two CTAs per SM) -- also the layout of the
// real OTF array, [k][xb][sub][y][CS] (each sub-block slab contiguous for its bulk copy).
template <typename T_, int RA, int RB, int RC, int NT_, int C_, int PR_, int CS_ = C_> struct FastPlan {
    typedef T_ T;
    typedef Fft3<T, -1, RA, RB, RC, NT_> Fwd;
    typedef Fft3<T, +1, RC, RB, RA, NT_> Inv;
    enum {
        L = RA * RB * RC, NT = NT_, C = C_, PR = PR_,
        SEQ = imax(Fwd::SEQ, Inv::SEQ),
        // column CTA: C interleaved sequences; the small offset spreads them over banks
        LSM_COL = (SEQ + 15) / 16 * 16 + 16 / C_,
        COL_THREADS = NT_ * C_,
        COL_SMEM_ELEMS = 2 * C_ * LSM_COL,          // two ping-pong buffers
        COL_OTF_ELEMS = C_ * RA * RB * RC,          // + one staged OTF slab [L][C] (bulk copy)
        COL_TW = 256,                               // + base twiddles ...
        COL_TW_PAD = 256 + 32,                      // ... + the pass-B ones gathered (ColTwTable, fft_static.cuh)
        CS = CS_, NSUB = C_ / CS_,
        LSM_SUB = (SEQ + 15) / 16 * 16 + 16 / CS_,
        SUB_THREADS = NT_ * CS_,
        SUB_SMEM_ELEMS = 2 * CS_ * LSM_SUB,
        // row CTA: PR groups of NTG threads (whole warps), one row pair each
        NTG = (NT_ + 31) / 32 * 32,
        LSM_ROW = (SEQ + 15) / 16 * 16 + 16,    // multiple of 128 bytes: tensor-map copies land here
        ROW_THREADS = NTG * PR_,
        // two ping-pong buffers + staging per pair: 2 measurement rows of <= L pixels (ROW_MID),
        // 2 normalisation rows + 2 estimate rows (ROW_FINAL)
        ROW_BUFS = 3, ROW_FINAL_BUFS = sizeof(T_) == 4 ? 4 : 3,
        ROW_SMEM_ELEMS = ROW_BUFS * PR_ * LSM_ROW,
        VREG = imax(Fwd::VREG, Inv::VREG),
        NKEEP = Fwd::MC * Fwd::RC,
        // Hermitian split: thread t owns bins t + q*NC; its mirror bins live in thread NC - t
        NC = Fwd::NC, RCF = RC, QH = (RC - 1) / 2, PX = Fwd::NC + 1
    };
    static_assert(Fwd::MC * Fwd::RC == Inv::MA * Inv::RA, "pass C / pass A operand sets must match");
    static_assert(Inv::MC * Inv::RC == Fwd::MA * Fwd::RA, "pass C / pass A operand sets must match");
    static_assert(Fwd::MC == 1 && Fwd::NC == NT_, "row split assumes one pass-C butterfly per thread");
    static_assert(Inv::MA == 1 && Inv::NA == NT_, "row unpack assumes one pass-A butterfly per thread");
    static_assert(RC % 2 == 1, "row split assumes an odd last radix");
    static_assert(C_ % CS_ == 0, "sub-blocks must tile a column block");
    static_assert(RA * RC < 256 && RC * RA < 256 && NT_ <= 256 && RA * RB * RC >= 256,
                  "base twiddles of both transforms must sit in the first COL_TW table entries");
    static_assert((RC - QH) * PX <= LSM_ROW, "mirror exchange must fit one row buffer");
    static_assert(2 * L * sizeof(T_) + sizeof(mbar_t) <= LSM_ROW * 2 * sizeof(T_),
                  "two staged rows + their mbarrier must fit one row buffer");
};

template <class P> struct ColRegs {
    cplx<typename P::T> keep[P::NKEEP];  // forward spectrum (COL_H) or Fourier-domain sum (COL_HT)
    cplx<typename P::T> v[P::VREG];
};
template <class P> struct RowRegs {
    cplx<typename P::T> v[P::VREG];
    typename P::Fwd::Tw twf;
    typename P::Inv::Tw twi;
    cplx<typename P::T> ramp0;           // crop-offset phase ramp at this thread's first bin
};

#ifndef LSTED_ROW_BULK_STAGE
#define LSTED_ROW_BULK_STAGE 1   // measurement / normalisation / estimate rows by cp.async.bulk
#endif
#ifndef LSTED_COL_HT_PRELOAD
#define LSTED_COL_HT_PRELOAD 1   // COL_HT: next orientation's pass-A operands requested before the barrier
#endif
#ifndef LSTED_COL_STAGE_OTF
#define LSTED_COL_STAGE_OTF 1   // OTF slabs travel global -> shared by cp.async.bulk, one k ahead
#endif
// + the first COL_TW twiddles in shared memory: every base twiddle of the plan
// (tw[RC*(j%RA)], tw[j], j < NT) has an index below that, and with >200 KB of shared memory
// per CTA the L1 is too small to keep the global table resident (the LDGs at the head of
// each pass went to L2)
template <class P> LSTED_HD size_t fast_col_smem_bytes() {
    return sizeof(cplx<typename P::T>) * (size_t)(P::COL_SMEM_ELEMS + P::COL_TW_PAD +
                                                  (LSTED_COL_STAGE_OTF ? P::COL_OTF_ELEMS : 0)) +
           (LSTED_COL_STAGE_OTF ? 16 : 0);
}
// column CTA on a sub-block of CS columns with real OTFs: exchange buffers + twiddles + one
// real OTF slab [L][CS] + its mbarrier
template <class P> LSTED_HD size_t fast_col_sub_smem_bytes() {
    return sizeof(cplx<typename P::T>) * (size_t)(P::SUB_SMEM_ELEMS + P::COL_TW_PAD) +
           sizeof(typename P::T) * (size_t)P::CS * P::L + 16;
}
template <class P> LSTED_HD size_t fast_row_smem_bytes(int mode, bool lean = false) {
    const int bufs = lean ? (mode == ROW_FINAL ? 3 : 2) : mode == ROW_FINAL ? (int)P::ROW_FINAL_BUFS : (int)P::ROW_BUFS;
    return sizeof(cplx<typename P::T>) * (size_t)bufs * P::PR * P::LSM_ROW;
}

// ---------------------------------------------------------------------------
// Column kernels
// ---------------------------------------------------------------------------
// XB2 slab addressing for the rows y = y0 + q*STEP a column thread touches (q is a
// compile-time index): slab2_index(y, c) = y*C + 2c - (y & 1)*(C - 1), and the parity of y
// is the parity of y0, flipped for odd q when STEP is odd -- two base offsets per thread
// and immediate offsets per q instead of integer work per access.
template <class P, int STEP> struct SlabRows {
    long long base[2];   // element offset at q even / q odd, without the q*STEP*C term
    LSTED_HD SlabRows(int y0, int c) {
        const int par = y0 & 1;
        const long long lin = (long long)y0 * P::C + 2 * c;
        base[0] = lin - par * (P::C - 1);
        base[1] = lin - ((STEP % 2) ? 1 - par : par) * (P::C - 1);
    }
    LSTED_HD long long at(int q) const { return base[q & 1] + (long long)q * STEP * P::C; }
};

// Forward pass-A operands of column c from an XB2 slab with `rows` valid rows.
template <class P>
LSTED_HD void col_load_fwd_a(cplx<typename P::T>* v, int t, int c, const cplx<typename P::T>* slab,
                             int rows) {
    typedef typename P::Fwd F;
    typedef typename P::T T;
    LSTED_UNROLL
    for (int m = 0; m < F::MA; ++m) {
        const int j = t + m * P::NT;
        if (j < F::NA) {
            const SlabRows<P, F::NA> sr(j, c);
            LSTED_UNROLL
            for (int q = 0; q < F::RA; ++q) {
                const int y = j + q * F::NA;
                v[m * F::RA + q] = (y < rows) ? slab[sr.at(q)] : mk<T>(0, 0);
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
            const SlabRows<P, I::NC> sr(j - sy, c);
            LSTED_UNROLL
            for (int q = 0; q < I::RC; ++q) {
                const int y = j + q * I::NC - sy;
                if (y >= 0 && y < Ny) slab[sr.at(q)] = v[m * I::RC + q];
            }
        }
    }
}

// v = keep (.) otf on the pass-C / pass-A operand set (rows j + q*NC of the OTF slab)
template <class P, bool ACCUMULATE>
LSTED_HD void col_otf_product(ColRegs<P>& r, int t, int c, const cplx<typename P::T>* otf, bool first) {
    // (complex OTFs: always whole C-column blocks, [y][C])
    typedef typename P::Fwd F;
    typedef typename P::T T;
    LSTED_UNROLL
    for (int m = 0; m < F::MC; ++m) {
        const int j = t + m * P::NT;
        if (j < F::NC) {
            LSTED_UNROLL
            for (int q = 0; q < F::RC; ++q) {
                const cplx<T> o = otf[(size_t)(j + q * F::NC) * P::C + c];
                if (ACCUMULATE) {   // keep starts at zero: one complex multiply-accumulate
                    (void)first;
                    r.keep[m * F::RC + q] = cmac(r.v[m * F::RC + q], o, r.keep[m * F::RC + q]);
                } else {
                    r.v[m * F::RC + q] = r.keep[m * F::RC + q] * o;
                }
            }
        }
    }
}

// the same with a REAL (centred, point-symmetric PSF) OTF: a real scaling / real FMA per element.
// Real OTF slabs are laid out in sub-blocks of CS columns, [sub][y][CS]; `otf` points at the
// thread's sub-block and column: &slab[(c / CS) * CS * L + c % CS] (sub-block CTAs: &slab[c]).
template <class P> LSTED_HD const typename P::T* real_otf_at(const typename P::T* slab, int c) {
    return slab + (size_t)(c / P::CS) * P::CS * P::L + c % P::CS;
}
template <class P, bool ACCUMULATE>
LSTED_HD void col_otf_product_real(ColRegs<P>& r, int t, const typename P::T* otf) {
    typedef typename P::Fwd F;
    typedef typename P::T T;
    LSTED_UNROLL
    for (int m = 0; m < F::MC; ++m) {
        const int j = t + m * P::NT;
        if (j < F::NC) {
            LSTED_UNROLL
            for (int q = 0; q < F::RC; ++q) {
                const T o = otf[(size_t)(j + q * F::NC) * P::CS];
                if (ACCUMULATE) r.keep[m * F::RC + q] = fma_real(r.v[m * F::RC + q], o, r.keep[m * F::RC + q]);
                else r.v[m * F::RC + q] = scale(r.keep[m * F::RC + q], o);
            }
        }
    }
}

// Compile-time image geometry of the column kernels (rows and crop offset): as for the row
// kernels, the `0 <= y < Ny` tests then survive only on the first and last butterfly leg.
struct ColGeomRuntime { enum { NY = 0, SY = 0 }; };
template <int NY_, int SY_> struct ColGeomFixed { enum { NY = NY_, SY = SY_ }; };

// RO: the OTFs are real (a.otf_real, see OtfCenterArgs) -- half the bytes to stage and stream.
// SUB (needs RO): the CTA works on a sub-block of CS of the C columns of a block
// (`block` = xb * NSUB + sub): NT*CS threads and about 40 % of the shared memory, so TWO CTAs
// share an SM and one runs butterflies while the other waits at its barrier; the HBM layout
// (C-column blocks, 64-byte row-pair chunks) does not change: a sub-block reads / writes
// whole 32-byte sectors of them.
template <int MODE, class P, class Ctx, class G = ColGeomRuntime, bool RO = false, bool SUB = false>
LSTED_HD void col_fast_body(Ctx& cx, int block, const ColArgs<typename P::T>& a,
                            cplx<typename P::T>* smem, ColRegs<P>* regs, G = G()) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    const ConvGeom& g = a.g;
    const int Ny = G::NY ? (int)G::NY : g.Ny, Ly = P::L;  // == g.Ly (checked at launch)
    const int sy = G::NY ? (int)G::SY : g.sy;
    static_assert(!SUB || RO, "sub-block column CTAs read the real OTF layout");
    enum { CC = SUB ? (int)P::CS : (int)P::C,                 // columns of this CTA
           LSM = SUB ? (int)P::LSM_SUB : (int)P::LSM_COL,     // shared elements per sequence
           NTHR = P::NT * CC, XELEMS = 2 * CC * LSM };
    const int xb = SUB ? block / (int)P::NSUB : block;
    const int sub = SUB ? block - xb * (int)P::NSUB : 0;
    const size_t slab_ly = (size_t)P::C * Ly, img_ly = (size_t)g.nxb * slab_ly;
    const size_t slab_ny = (size_t)P::C * even_rows(Ny), img_ny = (size_t)g.nxb * slab_ny;   // XB2
    cplx<T>* const tw_s = smem + (size_t)XELEMS;   // base twiddles (filled below)
    typedef ColTwTable<cplx<T>, P::COL_TW, F::RC, I::RC> TwTable;
    static_assert(F::RA <= 16 && I::RA <= 16 && (int)TwTable::ELEMS == (int)P::COL_TW_PAD, "gathered pass-B twiddles: 16 slots per transform");
    const TwTable tw = {tw_s};
    cplx<T>* const buf0 = smem;
    cplx<T>* const buf1 = smem + (size_t)CC * LSM;
    const int K = a.K;
    // debug build: the column block and sub-block of this CTA exist, the OTF slab is a whole
    // number of 16-byte units inside [nxb][C][Ly] of orientation 0 .. K-1
    LSTED_DCHECK(xb >= 0 && xb < g.nxb && sub >= 0 && sub < (SUB ? (int)P::NSUB : 1) && K >= 1);

#define LSTED_COL_IDS                                          \
    const int t = tid / CC, cl = tid - t * CC, c = sub * CC + cl;  \
    LSTED_ASSUME(tid >= 0 && t < P::NT);                       \
    cplx<T>* const s0 = buf0 + cl * LSM;                       \
    cplx<T>* const s1 = buf1 + cl * LSM;                       \
    (void)cl;

    // OTF slabs are staged in shared memory by the bulk-copy engine: the slab of
    // orientation k+1 is requested as soon as every thread is done with slab k and
    // lands while the transform of k runs (no registers, no per-thread copy work).
    const bool stage = LSTED_COL_STAGE_OTF != 0;
    cplx<T>* const otf_s = tw_s + P::COL_TW_PAD;
    mbar_t* const mbar = SUB ? (mbar_t*)((T*)otf_s + (size_t)P::CS * Ly)
                             : (mbar_t*)(otf_s + (size_t)P::COL_OTF_ELEMS);
    // one OTF slab: complex, or real when RO (`otf0` then counts T elements; a sub-block's
    // slab is the contiguous [Ly][CS] piece of the block's [NSUB][Ly][CS])
    typedef typename std::conditional<RO, T, cplx<T> >::type OtfT;
    const unsigned slab_bytes = (unsigned)((SUB ? (size_t)P::CS * Ly : slab_ly) * sizeof(OtfT));
    const OtfT* otf0 = (RO ? (const OtfT*)a.otf_real : (const OtfT*)a.otf) + (size_t)xb * slab_ly +
                       (SUB ? (size_t)sub * P::CS * Ly : 0);
    // this thread's column inside a staged / streamed real slab
#define LSTED_OTF_AT(base) (SUB ? (const T*)(base) + cl : real_otf_at<P>((const T*)(base), c))
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        // (hoisting base twiddles into registers costs spills at 96 registers / 576 threads)
        for (int i = tid; i < (int)TwTable::ELEMS; i += NTHR) tw_s[i] = a.tw[TwTable::source_index(i)];
        if (MODE == COL_HT) {
            LSTED_UNROLL
            for (int i = 0; i < P::NKEEP; ++i) r.keep[i] = mk<T>(0, 0);
        }
        if (stage && tid == 0) mbar_init(mbar);
    });
    if (MODE == COL_H) {
        const cplx<T>* src = a.src + (size_t)xb * slab_ny;
        cplx<T>* dst0 = a.dst + (size_t)xb * slab_ny;
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            if (stage && tid == 0) bulk_load(otf_s, otf0, slab_bytes, mbar);
            col_load_fwd_a<P>(r.v, t, c, src, Ny);
            F::pass_a(r.v, t, s0);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            F::load_b(r.v, t, s0, tw);
            F::pass_b(r.v, t, s1);
        });
        // k-loop: [finish transform k-1 and store it | product k, inverse pass A] then pass B
        for (int k = 0; k <= K; ++k) {
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                LSTED_COL_IDS
                // HBM -> L2 ahead of the bulk copy (staged) or of the loads (direct)
                const int ahead = stage ? 2 : 1;
                if (k + ahead < K)
                    prefetch_l2_range(otf0 + (size_t)(k + ahead) * img_ly, slab_bytes, tid, NTHR);
                if (k == 0) {
                    F::pass_c(r.v, t, s1, tw);
                    LSTED_UNROLL
                    for (int i = 0; i < P::NKEEP; ++i) r.keep[i] = r.v[i];
                } else {
                    I::pass_c(r.v, t, s1, tw);
                    col_store_inv_c<P>(r.v, t, c, dst0 + (size_t)(k - 1) * img_ny, sy, Ny);
                }
                if (k < K) {
                    if (stage) {
                        mbar_wait(mbar, (unsigned)(k & 1));
                        if (RO) col_otf_product_real<P, false>(r, t, LSTED_OTF_AT(otf_s));
                        else col_otf_product<P, false>(r, t, c, otf_s, false);
                    } else {
                        if (RO) col_otf_product_real<P, false>(r, t, LSTED_OTF_AT(otf0 + (size_t)k * img_ly));
                        else col_otf_product<P, false>(r, t, c, (const cplx<T>*)(otf0 + (size_t)k * img_ly), false);
                    }
                    I::pass_a(r.v, t, s0);
                }
            });
            if (k == K) break;
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                LSTED_COL_IDS
                if (stage && tid == 0 && k + 1 < K)
                    bulk_load(otf_s, otf0 + (size_t)(k + 1) * img_ly, slab_bytes, mbar);
                I::load_b(r.v, t, s0, tw);
                I::pass_b(r.v, t, s1);
            });
        }
        return;
    }
    // COL_HT: per k [forward pass A of k | accumulate product k-1] then [pass B | operands of k+1].
    // The pass-A operands of orientation k+1 are requested from global memory at the END of the
    // pass-B phase of k, into the registers pass B has just emptied, and are consumed right
    // after the barrier: their latency hides behind the barrier wait instead of sitting between
    // the product and pass A (ncu: long_sb 16 % of col_ht's stall samples before).
    const cplx<T>* src0 = a.src + (size_t)xb * slab_ny;
    const bool preload = LSTED_COL_HT_PRELOAD != 0;
    if (preload && K > 0)
        cx.phase_nosync(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            col_load_fwd_a<P>(r.v, t, c, src0, Ny);
        });
    for (int k = 0; k <= K; ++k) {
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            if (k + 1 < K && !a.src_same)
                prefetch_l2_range(src0 + (size_t)(k + 1) * img_ny, slab_ny * sizeof(cplx<T>), tid,
                                  NTHR);
            if (stage) {
                // slab m is consumed in iteration m+1; slab k is requested once slab k-1 is done with
                if (k + 1 < K) prefetch_l2_range(otf0 + (size_t)(k + 1) * img_ly, slab_bytes, tid, NTHR);
                if (tid == 0 && k == 0) bulk_load(otf_s, otf0, slab_bytes, mbar);
            } else if (k < K) {   // product k happens at the start of the next phase
                prefetch_l2_range(otf0 + (size_t)k * img_ly, slab_bytes, tid, NTHR);
            }
            // (pass A writes the first exchange buffer, pass C below reads the second one)
            if (preload && k < K) F::pass_a(r.v, t, s0);
            if (k > 0) {
                F::pass_c(r.v, t, s1, tw);
                if (stage) {
                    mbar_wait(mbar, (unsigned)((k - 1) & 1));
                    if (RO) col_otf_product_real<P, true>(r, t, LSTED_OTF_AT(otf_s));
                    else col_otf_product<P, true>(r, t, c, otf_s, k == 1);
                } else {
                    if (RO) col_otf_product_real<P, true>(r, t, LSTED_OTF_AT(otf0 + (size_t)(k - 1) * img_ly));
                    else col_otf_product<P, true>(r, t, c, (const cplx<T>*)(otf0 + (size_t)(k - 1) * img_ly), k == 1);
                }
            }
            if (k < K) {
                if (!preload) {
                    col_load_fwd_a<P>(r.v, t, c, src0 + (a.src_same ? 0 : (size_t)k * img_ny), Ny);
                    F::pass_a(r.v, t, s0);
                }
            } else {
                LSTED_UNROLL
                for (int i = 0; i < P::NKEEP; ++i) r.v[i] = r.keep[i];
                I::pass_a(r.v, t, s0);
            }
        });
        if (k == K) break;
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            if (stage && tid == 0 && k > 0 && k < K)
                bulk_load(otf_s, otf0 + (size_t)k * img_ly, slab_bytes, mbar);
            F::load_b(r.v, t, s0, tw);
            F::pass_b(r.v, t, s1);
            if (preload && k + 1 < K)
                col_load_fwd_a<P>(r.v, t, c, src0 + (a.src_same ? 0 : (size_t)(k + 1) * img_ny), Ny);
        });
    }
    cplx<T>* dst = a.dst + (size_t)xb * slab_ny;
    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        LSTED_COL_IDS
        I::load_b(r.v, t, s0, tw);
        I::pass_b(r.v, t, s1);
    });
    cx.phase_nosync(regs, [&](int tid, ColRegs<P>& r) {
        LSTED_COL_IDS
        I::pass_c(r.v, t, s1, tw);
        col_store_inv_c<P>(r.v, t, c, dst, sy, Ny);
    });
#undef LSTED_COL_IDS
#undef LSTED_OTF_AT
}

// ---------------------------------------------------------------------------
// COL_HT with the orientations sharded over GPUs: the sum of the ranks' partial spectra
// runs INSIDE the kernel over NVLink peer memory (no all-reduce after it).
//
//   * block xb is finished by its owner rank xb % world: the other ranks push their
//     partial Fourier-domain sum (the `keep` registers, 69 KB) into the owner's receive
//     slab and raise a flag; the owner waits for world-1 flags, adds, runs the ONE inverse
//     transform and writes the cropped column block into the spectrum of EVERY rank;
//   * the kernel is persistent (one CTA per SM walks several blocks) and every rank walks
//     the blocks it does NOT own first: a partial is pushed with plain posted stores and
//     its flag is released only two orientations into the NEXT block, when those stores
//     have long drained, so neither the NVLink burst nor the fence sits on the critical
//     path; by the time a rank reaches the blocks it owns, the peers' partials are there;
//   * `p2p_done` counts finished blocks on every rank; the next kernel on the stream is
//     held back by a one-thread wait until nxb blocks of this reduction are counted.
// ---------------------------------------------------------------------------
LSTED_HD int p2p_block_at(int pos, int me, int world, int nxb) {
    // positions 0 .. : the blocks with xb % world != me in ascending order, then the owned ones
    const int owned = (nxb - me + world - 1) / world;          // #{xb < nxb : xb % world == me}
    const int others = nxb - owned;
    if (pos >= others) return me + (pos - others) * world;
    // pos-th block that is not congruent to me: every run of `world` blocks holds world-1 of them
    const int run = pos / (world - 1), off = pos - run * (world - 1);
    return run * world + (off < me ? off : off + 1);
}

template <class P, class Ctx, class G = ColGeomRuntime, bool RO = false>
LSTED_HD void col_ht_p2p_body(Ctx& cx, int cta, int ncta, const ColArgs<typename P::T>& a,
                              cplx<typename P::T>* smem, ColRegs<P>* regs, G = G()) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    const ConvGeom& g = a.g;
    const int Ny = G::NY ? (int)G::NY : g.Ny, Ly = P::L;
    const int sy = G::NY ? (int)G::SY : g.sy;
    const size_t slab_ly = (size_t)P::C * Ly, img_ly = (size_t)g.nxb * slab_ly;
    const size_t slab_ny = (size_t)P::C * even_rows(Ny), img_ny = (size_t)g.nxb * slab_ny;
    cplx<T>* const tw_s = smem + (size_t)P::COL_SMEM_ELEMS;
    const cplx<T>* tw = tw_s;
    cplx<T>* const buf0 = smem;
    cplx<T>* const buf1 = smem + (size_t)P::C * P::LSM_COL;
    cplx<T>* const otf_s = tw_s + P::COL_TW_PAD;
    mbar_t* const mbar = (mbar_t*)(otf_s + (size_t)P::COL_OTF_ELEMS);
    typedef typename std::conditional<RO, T, cplx<T> >::type OtfT;   // real OTFs: see col_fast_body
    const unsigned slab_bytes = (unsigned)(slab_ly * sizeof(OtfT));
    const int K = a.K, world = a.p2p_world, me = a.p2p_rank;
    // one partial sum in the receive slab: pairs of register values, so that a thread moves
    // 16 bytes per store / load (512 contiguous bytes per warp on the NVLink side)
    enum { NPAIR = (P::NKEEP + 1) / 2 };
    const size_t part = (size_t)2 * NPAIR * P::COL_THREADS;

#define LSTED_COL_IDS                                          \
    const int t = tid / P::C, c = tid - t * P::C;              \
    cplx<T>* const s0 = buf0 + c * P::LSM_COL;                 \
    cplx<T>* const s1 = buf1 + c * P::LSM_COL;

    cx.phase(regs, [&](int tid, ColRegs<P>& r) {
        (void)r;
        for (int i = tid; i < (int)P::COL_TW; i += (int)P::COL_THREADS) tw_s[i] = a.tw[i];
        if (tid == 0) mbar_init(mbar);
    });
    unsigned seq = 0;                  // OTF slabs consumed so far (mbarrier phase parity)
    unsigned* pending = 0;             // flag of the last pushed partial, not yet released
    unsigned finished = 0;             // owned blocks this CTA has written to every rank
    for (int pos = cta; pos < g.nxb; pos += ncta) {
        const int xb = p2p_block_at(pos, me, world, g.nxb);
        const int owner = xb % world;
        const cplx<T>* src0 = a.src + (size_t)xb * slab_ny;
        const OtfT* otf0 = (RO ? (const OtfT*)a.otf_real : (const OtfT*)a.otf) + (size_t)xb * slab_ly;
        for (int k = 0; k <= K; ++k) {
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                LSTED_COL_IDS
                if (k + 1 < K)
                    prefetch_l2_range(src0 + (size_t)(k + 1) * img_ny, slab_ny * sizeof(cplx<T>), tid,
                                      P::COL_THREADS);
                if (k + 1 < K) prefetch_l2_range(otf0 + (size_t)(k + 1) * img_ly, slab_bytes, tid, P::COL_THREADS);
                if (k == 0) {
                    if (tid == 0) bulk_load(otf_s, otf0, slab_bytes, mbar);
                    LSTED_UNROLL
                    for (int i = 0; i < P::NKEEP; ++i) r.keep[i] = mk<T>(0, 0);
                }
                if (k == 2 && pending) sys_fence();     // the previous block's pushes have drained
                if (k > 0) {
                    F::pass_c(r.v, t, s1, tw);
                    mbar_wait(mbar, (seq + (unsigned)(k - 1)) & 1u);
                    if (RO) col_otf_product_real<P, true>(r, t, real_otf_at<P>((const T*)otf_s, c));
                    else col_otf_product<P, true>(r, t, c, otf_s, k == 1);
                }
                if (k < K) {
                    col_load_fwd_a<P>(r.v, t, c, src0 + (size_t)k * img_ny, Ny);
                    F::pass_a(r.v, t, s0);
                }
            });
            if (k == K) break;
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                LSTED_COL_IDS
                if (tid == 0 && k > 0 && k < K)
                    bulk_load(otf_s, otf0 + (size_t)k * img_ly, slab_bytes, mbar);
                if (tid == 0 && k == 2 && pending) flag_release(pending, a.p2p_epoch);
                F::load_b(r.v, t, s0, tw);
                F::pass_b(r.v, t, s1);
            });
            if (k == 2) pending = 0;
        }
        seq += (unsigned)K;
        if (owner != me) {
            // push the partial sum: posted stores, released later (k == 2 of the next block / exit)
            if (pending) {   // (K < 3: the previous flag is still held back)
                cx.phase(regs, [&](int tid, ColRegs<P>& r) { (void)r; (void)tid; sys_fence(); });
                cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                    (void)r;
                    if (tid == 0) flag_release(pending, a.p2p_epoch);
                });
            }
            cplx<T>* out = a.p2p_recv[owner] + ((size_t)me * g.nxb + xb) * part;
            cx.phase_nosync(regs, [&](int tid, ColRegs<P>& r) {
                LSTED_UNROLL
                for (int i = 0; i < NPAIR; ++i)
                    store_pair(out + ((size_t)i * P::COL_THREADS + tid) * 2, r.keep[2 * i],
                               2 * i + 1 < P::NKEEP ? r.keep[2 * i + 1 < P::NKEEP ? 2 * i + 1 : 0] : mk<T>(0, 0));
            });
            pending = a.p2p_flags[owner] + (size_t)me * g.nxb + xb;
            continue;
        }
        // owned block: wait for the peers' partials, add, ONE inverse transform, write everywhere
        if (pending) {   // never wait while holding back a flag a peer may be waiting for
            cx.phase(regs, [&](int tid, ColRegs<P>& r) { (void)r; (void)tid; sys_fence(); });
            cx.phase(regs, [&](int tid, ColRegs<P>& r) {
                (void)r;
                if (tid == 0) flag_release(pending, a.p2p_epoch);
            });
            pending = 0;
        }
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            (void)r;
            if (tid < world && tid != me) flag_wait(a.p2p_flags[me] + (size_t)tid * g.nxb + xb, a.p2p_epoch);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            for (int src = 0; src < world; ++src) {
                if (src == me) continue;
                const cplx<T>* in = a.p2p_recv[me] + ((size_t)src * g.nxb + xb) * part;
                LSTED_UNROLL
                for (int i = 0; i < NPAIR; ++i) {
                    cplx<T> u, w;
                    load_pair_l2(in + ((size_t)i * P::COL_THREADS + tid) * 2, u, w);
                    r.keep[2 * i] = r.keep[2 * i] + u;
                    if (2 * i + 1 < P::NKEEP) r.keep[2 * i + 1 < P::NKEEP ? 2 * i + 1 : 0] = r.keep[2 * i + 1 < P::NKEEP ? 2 * i + 1 : 0] + w;
                }
            }
            LSTED_UNROLL
            for (int i = 0; i < P::NKEEP; ++i) r.v[i] = r.keep[i];
            I::pass_a(r.v, t, s0);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            I::load_b(r.v, t, s0, tw);
            I::pass_b(r.v, t, s1);
        });
        cx.phase(regs, [&](int tid, ColRegs<P>& r) {
            LSTED_COL_IDS
            I::pass_c(r.v, t, s1, tw);
            for (int peer = 0; peer < world; ++peer)
                col_store_inv_c<P>(r.v, t, c, a.p2p_spec[peer] + (size_t)xb * slab_ny, sy, Ny);
        });
        ++finished;
    }
    // everything this CTA sent has to be visible before its last flag and its counts
    cx.phase(regs, [&](int tid, ColRegs<P>& r) { (void)r; (void)tid; sys_fence(); });
    cx.phase_nosync(regs, [&](int tid, ColRegs<P>& r) {
        (void)r;
        if (tid == 0 && pending) flag_release(pending, a.p2p_epoch);
        if (tid < world && finished) sys_counter_add(a.p2p_done[tid], finished);
    });
#undef LSTED_COL_IDS
}

// ---------------------------------------------------------------------------
// Row kernels
// ---------------------------------------------------------------------------
// Compile-time image geometry of the row kernels: with Nx and the crop offset known, the
// per-pixel bounds tests `0 <= i < Nx` fold away for all but the first and last butterfly leg
// (and the row strides become constants).  RowGeomRuntime keeps everything in registers.
// A fixed geometry also promises an even number of rows: every pair is whole (`two`).
struct RowGeomRuntime { enum { NX = 0, SX = 0 }; };
template <int NX_, int SX_> struct RowGeomFixed { enum { NX = NX_, SX = SX_ }; };

// TMA: the spectrum chunks of the pair come and go through tensor-map bulk copies staged in
// the second exchange buffer (free while they are needed) instead of per-thread LDG / STG.
// TMA == 2 ("lean"): additionally no buffer of their own for the measurement (ROW_MID) /
// normalisation (ROW_FINAL) rows: they are staged in the first exchange buffer between the two
// transforms -- two buffers instead of three (ROW_MID, 6 CTAs/SM), three instead of four
// (ROW_FINAL, 4 CTAs/SM).  Needs 16-byte aligned rows (checked at launch).
template <int MODE, class P, class Ctx, class G = RowGeomRuntime, int TMA = 0>
LSTED_HD void row_fast_body(Ctx& cx, int block, const RowArgs<typename P::T>& a,
                            cplx<typename P::T>* smem, RowRegs<P>* regs, G = G()) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Nx = G::NX ? (int)G::NX : g.Nx;   // (checked at launch)
    const int Lx = P::L, C = P::C;  // == g.Lx, g.C (checked at launch)
    const int Py = (Ny + 1) / 2;
    const int bpi = (Py + P::PR - 1) / P::PR;
    const int img = block / bpi;
    const int pair0 = (block - img * bpi) * P::PR;
    const size_t real_off = (size_t)img * Ny * Nx;
    const int Nye = even_rows(Ny);
    const size_t spec_off = (size_t)img * g.nxb * C * Nye;
    const cplx<T>* tw = a.tw;
    // the data of a pair sit at logical positions shift + pixel during the transforms
    const int shift = (MODE == ROW_FWD) ? 0 : (G::NX ? (int)G::SX : g.sx);
    const size_t xb_stride = (size_t)Nye * C;  // elements between consecutive column blocks
    // XB2: the pair (row y, row y+1) of column c sits at ((xb*Nye + y)*C + 2c) + {0, 1}
    const bool LEAN = TMA == 2 && (MODE == ROW_MID || MODE == ROW_FINAL);
    // ROW_INV_SIM: dense rounds of exact PTRS attempts before the per-lane sampler loop takes
    // the rest (queue sizes fall 24 % -> 10 % -> 1 % -> 0.1 % -> 0.01 % of the pixels)
    enum { NOISE_ROUNDS = 4 };
    static_assert(2 * P::L + NOISE_ROUNDS + 1 <= 2 * P::LSM_ROW, "noise queue + its counters must fit one row buffer");
    const int NBUF = LEAN ? (MODE == ROW_MID ? 2 : 3) : MODE == ROW_FINAL ? (int)P::ROW_FINAL_BUFS : (int)P::ROW_BUFS;
    const bool stage_est = MODE == ROW_FINAL && P::ROW_FINAL_BUFS == 4 && !LEAN;
    enum { CHUNK = 2 * P::PR * P::C,                                  // complex numbers per chunk
           TMA_BYTES = 2 * kTmaBoxBlocks * CHUNK * (int)sizeof(cplx<T>) };  // two boxes cover nxb blocks
    enum { TMA_MB_OFF = P::LSM_ROW * (int)sizeof(cplx<T>) - 16 };   // its mbarrier: the buffer's last 16 bytes
    static_assert(!TMA || (P::PR == 1 && (P::L / 2 / P::C + 1) * CHUNK * (int)sizeof(cplx<T>) <= TMA_MB_OFF),
                  "staged spectrum + its mbarrier must fit the second exchange buffer");
    const unsigned row_bytes = (unsigned)(Nx * sizeof(T));
    const bool bulk_rows = LSTED_ROW_BULK_STAGE && row_bytes % 16 == 0 &&
                           (((size_t)a.aux | (size_t)a.real_out) & 15) == 0;

#define LSTED_ROW_IDS                                      \
    const int f = tid / P::NTG, t = tid - f * P::NTG;      \
    const int pair = pair0 + f;                            \
    const bool live = pair < Py && t < P::NT;              \
    const int y = 2 * pair;                                \
    const bool two = G::NX ? true : y + 1 < Ny;  /* fixed geometry: Ny even (checked at launch) */ \
    cplx<T>* const s0 = smem + (size_t)(NBUF * f) * P::LSM_ROW;      \
    cplx<T>* const s1 = smem + (size_t)(NBUF * f + 1) * P::LSM_ROW;  \
    T* const stage = (T*)(smem + (size_t)(NBUF * f + 2) * P::LSM_ROW); \
    T* const stage2 = (T*)(smem + (size_t)(NBUF * f + 3) * P::LSM_ROW); /* only if NBUF == 4 */ \
    (void)s0; (void)s1; (void)stage; (void)stage2; (void)y; (void)live; (void)two;

    // Software prefetch across CTAs: the operands of the row pairs that will be
    // scheduled roughly one wave later are pulled into L2 now (spectra: one
    // 2-row chunk per column block; measurement / normalisation rows: contiguous).
    if (MODE == ROW_MID || MODE == ROW_FINAL) {
        const int ahead = block + a.prefetch_ahead;
        const int img2 = ahead / bpi;
        const int pair2 = (ahead - img2 * bpi) * P::PR;
        if (a.prefetch_ahead > 0 && img2 < a.nimg && pair2 < Py) {
            const int y2 = 2 * pair2;
            const int rows2 = (Ny - y2) < 2 * P::PR ? (Ny - y2) : 2 * P::PR;
            const cplx<T>* sp = a.spec_in + (size_t)img2 * g.nxb * C * Nye + (size_t)y2 * C;
            const T* ax = (MODE == ROW_MID ? a.aux + (size_t)img2 * Ny * Nx : a.aux) + (size_t)y2 * Nx;
            cx.phase_nosync(regs, [&](int tid, RowRegs<P>& r) {   // prefetches only: no barrier
                (void)r;
                if (TMA) {
                    if (tid < 2) tma_prefetch_chunks(a.tmap_in, img2, y2, tid * (g.nxb - kTmaBoxBlocks), C);
                } else {
                    for (int xb = tid; xb < g.nxb; xb += P::ROW_THREADS)
                        prefetch_chunk(sp + (size_t)xb * xb_stride, (unsigned)(2 * P::PR * C * sizeof(cplx<T>)));
                }
                prefetch_l2_range(ax, (size_t)rows2 * Nx * sizeof(T), tid, P::ROW_THREADS);
                if (MODE == ROW_FINAL)
                    prefetch_l2_range(a.real_out + (size_t)y2 * Nx, (size_t)rows2 * Nx * sizeof(T), tid,
                                      P::ROW_THREADS);
            });
        }
    }
    if (MODE == ROW_FWD) {
        const T* src = a.real_in + real_off;
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (!live) return;
            LSTED_ASSUME(t >= 0 && t < P::NT);
            F::load_tw(r.twf, t, tw);
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
                            if (two) vb = src[(size_t)(y + 1) * Nx + i];
                        }
                        r.v[m * F::RA + q] = mk<T>(va, vb);
                    }
                }
            }
            F::pass_a(r.v, t, s0);
        });
    } else {
        // Hermitian unpack straight into the inverse pass-A registers (bins t + q*NC):
        // Z[i] = A[i] + i B[i] (i <= L/2), conj(A[L-i]) + i conj(B[L-i]) otherwise.
        // q < QH is always the lower half, q > QH always the upper half.
        const cplx<T>* src = a.spec_in + spec_off;
        if (TMA) {
            // request the pair's spectrum chunks; the barrier of this phase publishes the
            // initialised mbarrier to the threads that wait on it in the next one
            cx.phase(regs, [&](int tid, RowRegs<P>& r) {
                LSTED_ROW_IDS
                (void)r;
                if (t == 0 && pair < Py) {
                    mbar_t* const mb = (mbar_t*)((char*)s1 + TMA_MB_OFF);
                    mbar_init(mb);
                    bulk_expect(mb, (unsigned)TMA_BYTES);
                    // two boxes of kTmaBoxBlocks column blocks; the second one ends at the last block
                    // (it overlaps the first by a block or so: no out-of-bounds box rows)
                    for (int h = 0; h < 2; ++h) {
                        const int xb0 = h * (g.nxb - kTmaBoxBlocks);
                        tma_load_chunks<T>((char*)s1 + (size_t)xb0 * CHUNK * sizeof(cplx<T>), a.tmap_in, src, img, y,
                                           xb0, g.nxb, Nye, C, CHUNK, mb);
                    }
                }
            });
        }
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            mbar_t* const mb_spec = (mbar_t*)((char*)s1 + TMA_MB_OFF);
            if (MODE == ROW_FINAL && LEAN && pair < Py) {
                // lean ROW_FINAL: only the estimate rows go to the third buffer now; the
                // normalisation rows follow into the first exchange buffer (below)
                if (t == 0) {
                    mbar_t* const mb = (mbar_t*)(stage + 2 * P::L);
                    const T* e0 = a.real_out + (size_t)y * Nx;
                    const int nrow = two ? 2 : 1;
                    mbar_init(mb);
                    bulk_expect(mb, row_bytes * nrow);
                    for (int rr = 0; rr < nrow; ++rr) bulk_copy(stage + rr * P::L, e0 + (size_t)rr * Nx, row_bytes, mb);
                }
            }
            if ((MODE == ROW_MID || MODE == ROW_FINAL) && !LEAN && pair < Py) {
                // measurement (MID) or normalisation + estimate (FINAL) rows y, y+1 -> shared
                // memory, asynchronously (used after the inverse transform, two barriers from here)
                const T* m0 = (MODE == ROW_MID ? a.aux + real_off : a.aux) + (size_t)y * Nx;
                if (bulk_rows) {
                    // whole rows by the bulk-copy engine: one thread, one instruction per row,
                    // nothing through the LSU (the rows are 16-byte multiples, 16-byte aligned)
                    if (t == 0) {
                        mbar_t* const mb = (mbar_t*)(stage + 2 * P::L);
                        const int nrow = two ? 2 : 1;
                        mbar_init(mb);
                        bulk_expect(mb, row_bytes * nrow * (stage_est ? 2 : 1));
                        for (int rr = 0; rr < nrow; ++rr) bulk_copy(stage + rr * P::L, m0 + (size_t)rr * Nx, row_bytes, mb);
                        if (stage_est) {
                            const T* e0 = a.real_out + (size_t)y * Nx;
                            for (int rr = 0; rr < nrow; ++rr)
                                bulk_copy(stage2 + rr * P::L, e0 + (size_t)rr * Nx, row_bytes, mb);
                        }
                    }
                } else {
                    async_copy_row(stage, m0, Nx, t, P::NTG);
                    if (two) async_copy_row(stage + P::L, m0 + Nx, Nx, t, P::NTG);
                    if (stage_est) {
                        const T* e0 = a.real_out + (size_t)y * Nx;
                        async_copy_row(stage2, e0, Nx, t, P::NTG);
                        if (two) async_copy_row(stage2 + P::L, e0 + Nx, Nx, t, P::NTG);
                    }
                }
            }
            if (!live) return;
            LSTED_ASSUME(t >= 0 && t < P::NT);
            I::load_tw(r.twi, t, tw);
            if (MODE == ROW_MID || MODE == ROW_FINAL) {
                F::load_tw(r.twf, t, tw);
                r.ramp0 = tw[(t * shift) % Lx];
            }
            const cplx<T>* lo = src + ((size_t)(t / C) * Nye + y) * C + 2 * (t % C);
            const int tm = Lx - t;  // mirror of bin t; (tm - q*NC) is the mirror of bin t + q*NC
            if (TMA) mbar_wait(mb_spec, 0u);   // staged chunks: bin k sits at s1[2k], s1[2k + 1]
            LSTED_UNROLL
            for (int q = 0; q < I::RA; ++q) {
                const int i = t + q * P::NC;
                bool upper = q > P::QH;
                if (q == P::QH) upper = 2 * i > Lx;
                const cplx<T>* p;
                if (q < P::QH || (q == P::QH && !upper)) {
                    p = TMA ? s1 + 2 * i : lo + (size_t)(q * (P::NC / C)) * xb_stride;
                } else {
                    const int k = tm - q * P::NC;
                    p = TMA ? s1 + 2 * k : src + ((size_t)(k / C) * Nye + y) * C + 2 * (k % C);
                }
                cplx<T> A, B;
                load_pair(p, A, B);   // one 16-byte (fp32) access; the slot of a missing last row is never used
                if (!two) B = mk<T>(0, 0);
                // bin 0 and the Nyquist bin L/2 = t + q*NC (one q, one t: a compile-time q, so
                // the other RA - 2 iterations carry no test) hold the spectra of real rows
                if ((q == 0 && t == 0) ||
                    (P::L % 2 == 0 && q == (P::L / 2) / P::NC && t == (P::L / 2) % P::NC)) { A.y = 0; B.y = 0; }
                r.v[q] = upper ? mk<T>(A.x + B.y, B.x - A.y) : mk<T>(A.x - B.y, A.y + B.x);
            }
            I::pass_a(r.v, t, s0);
        });
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (live) {
                I::load_b(r.v, t, s0, r.twi);
                I::pass_b(r.v, t, s1);
            }
            if ((MODE == ROW_MID || MODE == ROW_FINAL) && !LEAN && !bulk_rows) async_copy_wait_all();   // visible after the barrier
        });
        // inverse pass C, the pointwise step on registers (logical position
        // idx = j + q*NC holds pixel idx - sx of rows y (re) and y+1 (im)) and the
        // forward pass A of the result, all in one phase.
        T* out = (MODE == ROW_FINAL) ? a.real_out : a.real_out + real_off;
        T* out2 = (MODE == ROW_INV_SIM) ? a.real_out2 + real_off : 0;
        if (LEAN) {
            // two-buffer variant: the first exchange buffer is idle from here until the forward
            // transform, so the measurement rows are staged in IT (bulk copy, own mbarrier in its
            // last 16 bytes) while inverse pass C runs; the forward transform then starts in the
            // second buffer (one more barrier, one shared-memory buffer less)
            cx.phase(regs, [&](int tid, RowRegs<P>& r) {
                LSTED_ROW_IDS
                if (t == 0 && pair < Py) {
                    mbar_t* const mb = (mbar_t*)((char*)s0 + TMA_MB_OFF);
                    const T* m0 = (MODE == ROW_MID ? a.aux + real_off : a.aux) + (size_t)y * Nx;
                    const int nrow = two ? 2 : 1;
                    mbar_init(mb);
                    bulk_expect(mb, row_bytes * nrow);
                    for (int rr = 0; rr < nrow; ++rr) bulk_copy((T*)s0 + rr * P::L, m0 + (size_t)rr * Nx, row_bytes, mb);
                }
                if (!live) return;
                LSTED_ASSUME(t >= 0 && t < P::NT);
                I::pass_c(r.v, t, s1, r.twi);
            });
        }
        cx.phase(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            if (MODE == ROW_INV_SIM && t < NOISE_ROUNDS + 1) ((int*)stage)[t] = 0;   // noise queue counters (below)
            if (!live) return;
            LSTED_ASSUME(t >= 0 && t < P::NT);
            if (!LEAN) I::pass_c(r.v, t, s1, r.twi);
            const T* const rows_s = LEAN ? (const T*)s0 : stage;   // staged measurement / normalisation rows
            (void)rows_s;
            if (LEAN) mbar_wait((mbar_t*)((char*)s0 + TMA_MB_OFF), 0u);
            if (LEAN && MODE == ROW_FINAL) mbar_wait((mbar_t*)(stage + 2 * P::L), 0u);   // estimate rows
            if ((MODE == ROW_MID || MODE == ROW_FINAL) && !LEAN && bulk_rows)
                mbar_wait((mbar_t*)(stage + 2 * P::L), 0u);    // the staged rows have landed
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
                                // noiseless image; the noise steps below read lambda from the
                                // idle first exchange buffer, not back from global memory
                                const T l0 = clip0(z.x), l1 = clip0(z.y);
                                out[o] = l0;
                                ((T*)s0)[i] = l0;
                                if (two) { out[o + Nx] = l1; ((T*)s0)[P::L + i] = l1; }
                            } else if (MODE == ROW_MID) {
                                w.x = rl_ratio<true>(rows_s[i], z.x);
                                if (two) w.y = rl_ratio<true>(rows_s[P::L + i], z.y);
                            } else {  // ROW_FINAL
                                w.x = (LEAN ? stage[i] : stage_est ? stage2[i] : out[o]) * fast_div(clip0(z.x), rows_s[i]);
                                out[o] = w.x;
                                if (two) {
                                    w.y = (LEAN ? stage[P::L + i] : stage_est ? stage2[P::L + i] : out[o + Nx]) *
                                          fast_div(clip0(z.y), rows_s[P::L + i]);
                                    out[o + Nx] = w.y;
                                }
                            }
                        }
                        r.v[m * I::RC + q] = w;
                    }
                }
            }
            if (MODE == ROW_MID || MODE == ROW_FINAL) F::pass_a(r.v, t, LEAN ? s1 : s0);
        });
        if (MODE == ROW_INV_SIM) {
            // Shot noise in dense steps.  (1) Every pixel this thread just wrote gets the squeeze
            // test of PTRS attempt 0 (or the whole small-lambda sampler); the ~24 % it does not
            // accept go to a queue in shared memory.  (2) Round r = 0..NOISE_ROUNDS-1: all
            // threads share the pixels of queue r, one exact attempt r each (logarithms on full
            // warps only); a pixel whose attempt is rejected moves to queue r + 1, so no lane
            // waits for another lane's retries (a per-lane retry loop runs every warp for its
            // unluckiest lane: 3.1 instead of 1.5 attempts per queued pixel).  (3) The handful
            // left after the last round finish with the plain sampler loop.  The counters are
            // the same as poisson_sample's, so the field is the one it would give everywhere.
            // queue counters: stage[0..NOISE_ROUNDS]; entries of even rounds behind them, of
            // odd rounds in the second exchange buffer; lambda rows in the first one.
            cx.phase(regs, [&](int tid, RowRegs<P>& r) {
                LSTED_ROW_IDS
                (void)r;
                if (!live) return;
                LSTED_ASSUME(t >= 0 && t < P::NT);
                int* const cnt = (int*)stage;
                int* const queue = cnt + NOISE_ROUNDS + 1;
                const T* const lam_s = (const T*)s0;
                const int nr = two ? 2 : 1;
                LSTED_NOUNROLL
                for (int mq = 0; mq < I::MC * I::RC; ++mq) {
                    const int m = mq / I::RC, q = mq - m * I::RC;
                    const int j = t + m * P::NT;
                    const int i = j + q * I::NC - shift;
                    if (!(j < I::NC && i >= 0 && i < Nx)) continue;
                    // both rows of the pair: two independent straight-line attempts the
                    // compiler interleaves (the fp64 sqrt / divisions are latency chains)
                    size_t o[2];
                    double lam[2], k[2];
                    bool ok[2];
                    LSTED_UNROLL
                    for (int rr = 0; rr < 2; ++rr) {
                        o[rr] = (size_t)(y + (rr < nr ? rr : 0)) * Nx + i;
                        lam[rr] = (double)lam_s[(rr < nr ? rr : 0) * P::L + i];
                        ok[rr] = poisson_fast_ptrs(lam[rr] >= 10.0 ? lam[rr] : 10.0, a.seed, o[rr],
                                                   a.img0 + img, k[rr]);
                    }
                    LSTED_UNROLL
                    for (int rr = 0; rr < 2; ++rr) {
                        if (rr >= nr) continue;
                        if (!(lam[rr] >= 10.0)) {
                            out2[o[rr]] = (T)(poisson_sample(lam[rr], a.seed, o[rr], a.img0 + img) + 1e-9);
                        } else if (ok[rr]) {
                            out2[o[rr]] = (T)(k[rr] + 1e-9);
                        } else {
                            queue[smem_counter_next(cnt)] = rr * Nx + i;
                        }
                    }
                }
            });
            LSTED_NOUNROLL
            for (int round = 0; round < NOISE_ROUNDS; ++round) {
                cx.phase(regs, [&](int tid, RowRegs<P>& r) {
                    LSTED_ROW_IDS
                    (void)r;
                    if (pair >= Py) return;
                    int* const cnt = (int*)stage;
                    const int* const qin = (round & 1) ? (const int*)s1 : cnt + NOISE_ROUNDS + 1;
                    int* const qout = (round & 1) ? cnt + NOISE_ROUNDS + 1 : (int*)s1;
                    const T* const lam_s = (const T*)s0;
                    const int n = cnt[round];
                    LSTED_NOUNROLL
                    for (int e = t; e < n; e += P::NTG) {
                        const int off = qin[e];                        // rr * Nx + i
                        const size_t o = (size_t)y * Nx + off;
                        const double lam = (double)lam_s[off < Nx ? off : P::L + off - Nx];
                        double k;
                        if (ptrs_attempt<true>(lam, a.seed, o, a.img0 + img, (uint32_t)round, k))
                            out2[o] = (T)(k + 1e-9);
                        else
                            qout[smem_counter_next(cnt + round + 1)] = off;
                    }
                });
            }
            cx.phase_nosync(regs, [&](int tid, RowRegs<P>& r) {
                LSTED_ROW_IDS
                (void)r;
                if (pair >= Py) return;
                const int* const cnt = (const int*)stage;
                const int* const qin = (NOISE_ROUNDS & 1) ? (const int*)s1 : cnt + NOISE_ROUNDS + 1;
                const T* const lam_s = (const T*)s0;
                const int n = cnt[NOISE_ROUNDS];
                LSTED_NOUNROLL
                for (int e = t; e < n; e += P::NTG) {
                    const int off = qin[e];
                    const size_t o = (size_t)y * Nx + off;
                    const double lam = (double)lam_s[off < Nx ? off : P::L + off - Nx];
                    out2[o] = (T)(poisson_sample(lam, a.seed, o, a.img0 + img, (uint32_t)NOISE_ROUNDS) + 1e-9);
                }
            });
        }
        if (MODE == ROW_INV_STORE || MODE == ROW_INV_SIM) return;
    }
    // (two-buffer ROW_MID: the forward transform runs with the two buffers swapped)
#define LSTED_ROW_FWD_BUFS cplx<T>* const fa = LEAN ? s1 : s0; cplx<T>* const fb = LEAN ? s0 : s1; (void)fa; (void)fb;
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        LSTED_ROW_FWD_BUFS
        if (!live) return;
        LSTED_ASSUME(t >= 0 && t < P::NT);
        F::load_b(r.v, t, fa, r.twf);
        F::pass_b(r.v, t, fb);
    });
    // Forward pass C; the upper half of the spectrum goes to the mirror thread.
    cx.phase(regs, [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        LSTED_ROW_FWD_BUFS
        if (!live) return;
        LSTED_ASSUME(t >= 0 && t < P::NT);
        F::pass_c(r.v, t, fb, r.twf);
        LSTED_UNROLL
        for (int q = P::QH; q < P::RCF; ++q) fa[(q - P::QH) * P::PX + t] = r.v[q];
    });
    // Hermitian split of the lower half, crop-offset phase ramp, XB store:
    // bins k = t + q*NC with mirror L - k = (NC - t) + (RC - 1 - q)*NC.
    cplx<T>* dst = a.spec_out + spec_off;
    auto split_and_store = [&](int tid, RowRegs<P>& r) {
        LSTED_ROW_IDS
        LSTED_ROW_FWD_BUFS
        if (!live) return;
        LSTED_ASSUME(t >= 0 && t < P::NT);
        const int tp = t == 0 ? 0 : P::NC - t;         // mirror thread
        const int qoff = t == 0 ? 1 : 0;               // thread 0 mirrors onto itself, one q up
        // TMA: the pair's chunks are assembled in s1 (bin k at s1[2k], s1[2k+1]) and leave in bulk
        cplx<T>* pk = TMA ? fb + 2 * t : dst + ((size_t)(t / C) * Nye + y) * C + 2 * (t % C);
        const size_t xb_stride = TMA ? (size_t)2 * C : (size_t)Nye * C;   // elements per column block
        // phase ramp exp(+2 pi i k shift / L) at bins k = t + q*NC: ramp0 * d^q with the
        // thread-independent step d (powers by binary splitting, no table gathers)
        cplx<T> dq[P::QH + 2];
        if (shift) F::template twiddle_powers<P::QH>(tw[(P::NC * shift) % Lx], dq);
        LSTED_UNROLL
        for (int q = 0; q <= P::QH; ++q) {
            const int k = t + q * P::NC;
            if (q == P::QH && 2 * k > Lx) {
                // past the Nyquist bin: only the zero padding of the last column block
                if (k < g.nxb * C)
                    store_pair(pk + (size_t)(q * (P::NC / C)) * xb_stride, mk<T>(0, 0), mk<T>(0, 0));
                continue;
            }
            const cplx<T> z1 = r.v[q];
            cplx<T> z2;
            const int qm = P::RCF - 1 - q + qoff;      // q of the mirror bin in thread tp
            if (t == 0 && q == 0) z2 = z1;
            else z2 = fa[(qm - P::QH) * P::PX + tp];
            cplx<T> oa = mk<T>((T)0.5 * (z1.x + z2.x), (T)0.5 * (z1.y - z2.y));  // (z1 + conj z2)/2
            cplx<T> ob = mk<T>((T)0.5 * (z1.y + z2.y), (T)0.5 * (z2.x - z1.x));  // (z1 - conj z2)/(2i)
            if (shift) {
                const cplx<T> ph = conj(q == 0 ? r.ramp0 : r.ramp0 * dq[q]);
                oa = oa * ph;
                ob = ob * ph;
            }
            // (a missing last row has its own never-read slot: the pair is always stored whole)
            store_pair(pk + (size_t)(q * (P::NC / C)) * xb_stride, oa, two ? ob : mk<T>(0, 0));
        }
        if (TMA) fence_async_smem();
    };
    if (!TMA) {
        cx.phase_nosync(regs, split_and_store);   // last phase: nothing to wait for
    } else {
        cx.phase(regs, split_and_store);
        cx.phase_nosync(regs, [&](int tid, RowRegs<P>& r) {
            LSTED_ROW_IDS
            LSTED_ROW_FWD_BUFS
            (void)r;
            if (t != 0 || pair >= Py) return;
            for (int h = 0; h < 2; ++h) {
                const int xb0 = h * (g.nxb - kTmaBoxBlocks);
                tma_store_chunks<T>((const char*)fb + (size_t)xb0 * CHUNK * sizeof(cplx<T>), a.tmap_out, dst, img, y,
                                    xb0, g.nxb, Nye, C, CHUNK);
            }
            tma_store_finish();
        });
    }
#undef LSTED_ROW_IDS
#undef LSTED_ROW_FWD_BUFS
}


// ---------------------------------------------------------------------------
// ROW_MID, two row pairs per thread group ("dual", fp32).  Same operator, same HBM
// layout and same arithmetic as row_fast_body<ROW_MID>; the elements are c2 (two
// independent packed-complex sequences side by side), so every twiddle, address,
// predicate and barrier is shared by four image rows, every shared-memory access
// is 128 bits wide, and each thread carries two independent dependency chains.
// Needs Ny % 4 == 0 (the caller falls back to the single-pair kernel otherwise).
// ---------------------------------------------------------------------------
template <class P> struct RowDual {
    typedef typename P::T T;
    typedef Fft3E<c2, -1, P::Fwd::RA, P::Fwd::RB, P::Fwd::RC, P::NT> F;
    typedef Fft3E<c2, +1, P::Inv::RA, P::Inv::RB, P::Inv::RC, P::NT> I;
    enum {
        SEQ = imax(F::SEQ, I::SEQ),
        LSM = (SEQ + 7) / 8 * 8 + 4,            // c2 elements per exchange buffer
        THREADS = P::NTG,
        STAGE_FLOATS = 4 * P::L,                // four measurement rows
        VREG = imax(F::VREG, I::VREG)
    };
    struct Regs {
        c2 v[VREG];
        typename F::Tw twf;
        typename I::Tw twi;
        cplx<T> ramp0;
    };
    static LSTED_HD size_t smem_bytes() { return sizeof(c2) * 2 * (size_t)LSM + sizeof(T) * (size_t)STAGE_FLOATS; }
    static_assert(sizeof(typename P::T) == 4, "dual row kernel is fp32 only");
    static_assert((P::RCF - P::QH) * P::PX <= LSM, "mirror exchange must fit one buffer");
};

template <class P, class Ctx>
LSTED_HD void row_mid_dual_body(Ctx& cx, int block, const RowArgs<typename P::T>& a,
                                unsigned char* smem_raw, typename RowDual<P>::Regs* regs) {
    typedef typename P::T T;
    typedef RowDual<P> D;
    typedef typename D::F F;
    typedef typename D::I I;
    typedef typename D::Regs R;
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Nx = g.Nx;
    const int Lx = P::L, C = P::C;
    const int qpi = Ny / 4;                        // quads per image
    const int img = block / qpi;
    const int y = 4 * (block - img * qpi);         // first of the four rows
    const size_t real_off = (size_t)img * Ny * Nx;
    const size_t spec_off = (size_t)img * g.nxb * C * Ny;   // Ny % 4 == 0: already even
    const cplx<T>* tw = a.tw;
    const int shift = g.sx;
    const size_t xb_stride = (size_t)Ny * C;
    c2* const s0 = (c2*)smem_raw;
    c2* const s1 = s0 + D::LSM;
    T* const stage = (T*)(s1 + D::LSM);
    const cplx<T>* src = a.spec_in + spec_off;
    const T* meas = a.aux + real_off + (size_t)y * Nx;

    // L2 prefetch for the CTA one wave ahead (as in row_fast_body)
    {
        const int ahead = block + a.prefetch_ahead;
        const int img2 = ahead / qpi;
        if (a.prefetch_ahead > 0 && img2 < a.nimg) {
            const int y2 = 4 * (ahead - img2 * qpi);
            const cplx<T>* sp = a.spec_in + (size_t)img2 * g.nxb * C * Ny + (size_t)y2 * C;
            const T* ax = a.aux + (size_t)img2 * Ny * Nx + (size_t)y2 * Nx;
            cx.phase_nosync(regs, [&](int tid, R& r) {
                (void)r;
                for (int xb = tid; xb < g.nxb; xb += D::THREADS)
                    prefetch_chunk(sp + (size_t)xb * xb_stride, (unsigned)(4 * C * sizeof(cplx<T>)));
                prefetch_l2_range(ax, (size_t)4 * Nx * sizeof(T), tid, D::THREADS);
            });
        }
    }
    // Hermitian unpack of the four half spectra straight into inverse pass-A registers
    cx.phase(regs, [&](int tid, R& r) {
        const int t = tid;
        LSTED_UNROLL
        for (int rr = 0; rr < 4; ++rr) async_copy_row(stage + rr * P::L, meas + (size_t)rr * Nx, Nx, t, D::THREADS);
        if (t >= P::NT) return;
        I::load_tw(r.twi, t, tw);
        F::load_tw(r.twf, t, tw);
        r.ramp0 = tw[(t * shift) % Lx];
        const cplx<T>* lo = src + ((size_t)(t / C) * Ny + y) * C + 2 * (t % C);
        const int u = Lx - t;   // mirror of bin t; (u - q*NC) is the mirror of bin t + q*NC
        LSTED_UNROLL
        for (int q = 0; q < I::RA; ++q) {
            const int i = t + q * P::NC;
            bool upper = q > P::QH;
            if (q == P::QH) upper = 2 * i > Lx;
            const cplx<T>* p;
            if (!upper) {
                p = lo + (size_t)(q * (P::NC / C)) * xb_stride;
            } else {
                const int k = u - q * P::NC;
                p = src + ((size_t)(k / C) * Ny + y) * C + 2 * (k % C);
            }
            cplx<T> A0, B0, A1, B1;   // rows y, y+1 and (one row pair = 2C elements further) y+2, y+3
            load_pair(p, A0, B0);
            load_pair(p + 2 * C, A1, B1);
            if ((q == 0 && t == 0) || 2 * i == Lx) { A0.y = 0; B0.y = 0; A1.y = 0; B1.y = 0; }
            // Z = A + iB below the Nyquist bin, conj(A) + i conj(B) of the mirror bin above it
            r.v[q] = upper ? mk2(mk<T>(A0.x + B0.y, B0.x - A0.y), mk<T>(A1.x + B1.y, B1.x - A1.y))
                           : mk2(mk<T>(A0.x - B0.y, A0.y + B0.x), mk<T>(A1.x - B1.y, A1.y + B1.x));
        }
        I::pass_a(r.v, t, s0);
    });
    cx.phase(regs, [&](int tid, R& r) {
        const int t = tid;
        if (t < P::NT) {
            I::load_b(r.v, t, s0, r.twi);
            I::pass_b(r.v, t, s1);
        }
        async_copy_wait_all();
    });
    // inverse pass C, ratio = measurement / clip(expected) on registers, forward pass A
    cx.phase(regs, [&](int tid, R& r) {
        const int t = tid;
        if (t >= P::NT) return;
        I::pass_c(r.v, t, s1, r.twi);
        LSTED_UNROLL
        for (int m = 0; m < I::MC; ++m) {
            const int j = t + m * P::NT;
            if (j < I::NC) {
                LSTED_UNROLL
                for (int q = 0; q < I::RC; ++q) {
                    const int i = j + q * I::NC - shift;
                    const c2 z = r.v[m * I::RC + q];
                    c2 w = mk2(mk<T>(0, 0), mk<T>(0, 0));
                    if (i >= 0 && i < Nx) {
                        w.a.x = rl_ratio<true>(stage[i], z.a.x);
                        w.a.y = rl_ratio<true>(stage[P::L + i], z.a.y);
                        w.b.x = rl_ratio<true>(stage[2 * P::L + i], z.b.x);
                        w.b.y = rl_ratio<true>(stage[3 * P::L + i], z.b.y);
                    }
                    r.v[m * I::RC + q] = w;
                }
            }
        }
        F::pass_a(r.v, t, s0);
    });
    cx.phase(regs, [&](int tid, R& r) {
        const int t = tid;
        if (t >= P::NT) return;
        F::load_b(r.v, t, s0, r.twf);
        F::pass_b(r.v, t, s1);
    });
    cx.phase(regs, [&](int tid, R& r) {
        const int t = tid;
        if (t >= P::NT) return;
        F::pass_c(r.v, t, s1, r.twf);
        LSTED_UNROLL
        for (int q = P::QH; q < P::RCF; ++q) s0[(q - P::QH) * P::PX + t] = r.v[q];
    });
    cplx<T>* dst = a.spec_out + spec_off;
    cx.phase_nosync(regs, [&](int tid, R& r) {
        const int t = tid;
        if (t >= P::NT) return;
        const int tp = t == 0 ? 0 : P::NC - t;         // mirror thread
        const int qoff = t == 0 ? 1 : 0;               // thread 0 mirrors onto itself, one q up
        cplx<T>* pk = dst + ((size_t)(t / C) * Ny + y) * C + 2 * (t % C);
        cplx<T> dq[P::QH + 2];
        if (shift) Fft3E<cplx<T>, -1, P::Fwd::RA, P::Fwd::RB, P::Fwd::RC, P::NT>::template twiddle_powers<P::QH>(
            tw[(P::NC * shift) % Lx], dq);
        LSTED_UNROLL
        for (int q = 0; q <= P::QH; ++q) {
            const int k = t + q * P::NC;
            cplx<T>* p = pk + (size_t)(q * (P::NC / C)) * xb_stride;
            if (q == P::QH && 2 * k > Lx) {
                // past the Nyquist bin: only the zero padding of the last column block
                if (k < g.nxb * C) {
                    store_pair(p, mk<T>(0, 0), mk<T>(0, 0));
                    store_pair(p + 2 * C, mk<T>(0, 0), mk<T>(0, 0));
                }
                continue;
            }
            const c2 z1 = r.v[q];
            c2 z2;
            const int qm = P::RCF - 1 - q + qoff;      // q of the mirror bin in thread tp
            if (t == 0 && q == 0) z2 = z1;
            else z2 = s0[(qm - P::QH) * P::PX + tp];
            // (z1 + conj z2)/2 and (z1 - conj z2)/(2i) of both sequences
            cplx<T> oa0 = mk<T>((T)0.5 * (z1.a.x + z2.a.x), (T)0.5 * (z1.a.y - z2.a.y));
            cplx<T> ob0 = mk<T>((T)0.5 * (z1.a.y + z2.a.y), (T)0.5 * (z2.a.x - z1.a.x));
            cplx<T> oa1 = mk<T>((T)0.5 * (z1.b.x + z2.b.x), (T)0.5 * (z1.b.y - z2.b.y));
            cplx<T> ob1 = mk<T>((T)0.5 * (z1.b.y + z2.b.y), (T)0.5 * (z2.b.x - z1.b.x));
            if (shift) {
                const cplx<T> ph = conj(q == 0 ? r.ramp0 : r.ramp0 * dq[q]);
                oa0 = oa0 * ph; ob0 = ob0 * ph; oa1 = oa1 * ph; ob1 = ob1 * ph;
            }
            store_pair(p, oa0, ob0);
            store_pair(p + 2 * C, oa1, ob1);
        }
    });
}

}  // namespace lsted

namespace lsted {

// ---------------------------------------------------------------------------
// Row kernels on the two-pass plan (Fft2E): L = RA*RC = 48*45 for 2160.
// 48 threads per row pair hold a whole butterfly leg (45 / 48 complex values) in registers;
// a transform costs ONE shared-memory exchange and ONE twiddle stage, the pair needs one
// exchange buffer + one staging buffer (33.8 KB: six pairs per SM), and a CTA of PR = 2
// pairs is three full warps.  Same operators, HBM layouts and pointwise arithmetic as
// row_fast_body; per thread the bins are t + q*48 (q < 45) and the pixels t + q*45 (q < 48).
// ---------------------------------------------------------------------------
template <typename T_, int RA_, int RC_, int C_, int PR_> struct FastPlan2 {
    typedef T_ T;
    enum { NT = imax(RA_, RC_) };
    typedef Fft2E<cplx<T>, -1, RA_, RC_, NT> Fwd;
    typedef Fft2E<cplx<T>, +1, RC_, RA_, NT> Inv;
    enum {
        L = RA_ * RC_, C = C_, PR = PR_,
        SEQ = imax(Fwd::SEQ, Inv::SEQ),
        LSM_ROW = (SEQ + 15) / 16 * 16 + 8,
        ROW_THREADS = NT * PR_,
        VREG = imax(Fwd::VREG, Inv::VREG),
        NC = Fwd::NC, RCF = Fwd::RC, QH = (Fwd::RC - 1) / 2, PX = Fwd::NC + 1
    };
    static_assert((int)Fwd::NC == (int)NT && (int)Inv::NA == (int)NT, "bins: one leg per thread, stride NT");
    static_assert((int)Fwd::NA == (int)Inv::NC, "pixels: forward pass A takes what inverse pass C leaves");
    static_assert(Fwd::RC % 2 == 1, "row split assumes an odd last radix");
    static_assert((Fwd::RC - QH) * PX <= LSM_ROW, "mirror exchange must fit one buffer");
    static_assert(NT % C_ == 0, "column blocks must not straddle thread legs");
};
template <class P> struct Row2Regs {
    cplx<typename P::T> v[P::VREG];
    typename P::Fwd::Tw twf;
    typename P::Inv::Tw twi;
    cplx<typename P::T> ramp0;
};
// exchange buffer + staging buffer (+ estimate rows for ROW_FINAL) per pair
template <class P> LSTED_HD size_t fast_row2_smem_bytes(int mode) {
    return sizeof(cplx<typename P::T>) * (size_t)(mode == ROW_FINAL ? 3 : 2) * P::PR * P::LSM_ROW;
}

template <int MODE, class P, class Ctx, class G = RowGeomRuntime>
LSTED_HD void row2_fast_body(Ctx& cx, int block, const RowArgs<typename P::T>& a,
                             cplx<typename P::T>* smem, Row2Regs<P>* regs, G = G()) {
    typedef typename P::T T;
    typedef typename P::Fwd F;
    typedef typename P::Inv I;
    typedef Row2Regs<P> R;
    const ConvGeom& g = a.g;
    const int Ny = g.Ny, Nx = G::NX ? (int)G::NX : g.Nx;
    const int Lx = P::L, C = P::C;
    const int Py = (Ny + 1) / 2;
    const int bpi = (Py + P::PR - 1) / P::PR;
    const int img = block / bpi;
    const int pair0 = (block - img * bpi) * P::PR;
    const size_t real_off = (size_t)img * Ny * Nx;
    const int Nye = even_rows(Ny);
    const size_t spec_off = (size_t)img * g.nxb * C * Nye;
    const cplx<T>* tw = a.tw;
    const int shift = (MODE == ROW_FWD) ? 0 : (G::NX ? (int)G::SX : g.sx);
    const size_t xb_stride = (size_t)Nye * C;
    const int NBUF = MODE == ROW_FINAL ? 3 : 2;

#define LSTED_ROW2_IDS                                     \
    const int f = tid / P::NT, t = tid - f * P::NT;        \
    const int pair = pair0 + f;                            \
    const bool live = pair < Py;                           \
    const int y = 2 * pair;                                \
    const bool two = y + 1 < Ny;                           \
    cplx<T>* const s0 = smem + (size_t)(NBUF * f) * P::LSM_ROW;      \
    cplx<T>* const sx = smem + (size_t)(NBUF * f + 1) * P::LSM_ROW;  /* staging / mirror */ \
    T* const stage = (T*)sx;                               \
    T* const stage2 = (T*)(smem + (size_t)(NBUF * f + 2) * P::LSM_ROW); /* ROW_FINAL only */ \
    (void)s0; (void)sx; (void)stage; (void)stage2; (void)y; (void)live; (void)two;

    if (MODE == ROW_MID || MODE == ROW_FINAL) {
        const int ahead = block + a.prefetch_ahead;
        const int img2 = ahead / bpi;
        const int pair2 = (ahead - img2 * bpi) * P::PR;
        if (a.prefetch_ahead > 0 && img2 < a.nimg && pair2 < Py) {
            const int y2 = 2 * pair2;
            const int rows2 = (Ny - y2) < 2 * P::PR ? (Ny - y2) : 2 * P::PR;
            const cplx<T>* sp = a.spec_in + (size_t)img2 * g.nxb * C * Nye + (size_t)y2 * C;
            const T* ax = (MODE == ROW_MID ? a.aux + (size_t)img2 * Ny * Nx : a.aux) + (size_t)y2 * Nx;
            cx.phase_nosync(regs, [&](int tid, R& r) {
                (void)r;
                for (int xb = tid; xb < g.nxb; xb += P::ROW_THREADS)
                    prefetch_chunk(sp + (size_t)xb * xb_stride, (unsigned)(2 * P::PR * C * sizeof(cplx<T>)));
                prefetch_l2_range(ax, (size_t)rows2 * Nx * sizeof(T), tid, P::ROW_THREADS);
                if (MODE == ROW_FINAL)
                    prefetch_l2_range(a.real_out + (size_t)y2 * Nx, (size_t)rows2 * Nx * sizeof(T), tid,
                                      P::ROW_THREADS);
            });
        }
    }
    if (MODE == ROW_FWD) {
        const T* src = a.real_in + real_off;
        cx.phase(regs, [&](int tid, R& r) {
            LSTED_ROW2_IDS
            if (!live) return;
            F::load_tw(r.twf, t, tw);
            if (t < F::NA) {
                LSTED_UNROLL
                for (int q = 0; q < F::RA; ++q) {
                    const int i = t + q * F::NA;
                    T va = 0, vb = 0;
                    if (i < Nx) {
                        va = src[(size_t)y * Nx + i];
                        if (two) vb = src[(size_t)(y + 1) * Nx + i];
                    }
                    r.v[q] = mk<T>(va, vb);
                }
            }
            F::pass_a(r.v, t, s0);
        });
    } else {
        const cplx<T>* src = a.spec_in + spec_off;
        cx.phase(regs, [&](int tid, R& r) {
            LSTED_ROW2_IDS
            if ((MODE == ROW_MID || MODE == ROW_FINAL) && live) {
                const T* m0 = (MODE == ROW_MID ? a.aux + real_off : a.aux) + (size_t)y * Nx;
                async_copy_row(stage, m0, Nx, t, P::NT);
                if (two) async_copy_row(stage + P::L, m0 + Nx, Nx, t, P::NT);
                if (MODE == ROW_FINAL) {
                    const T* e0 = a.real_out + (size_t)y * Nx;
                    async_copy_row(stage2, e0, Nx, t, P::NT);
                    if (two) async_copy_row(stage2 + P::L, e0 + Nx, Nx, t, P::NT);
                }
            }
            if (MODE == ROW_INV_SIM && t == 0) *(int*)stage = 0;   // noise queue (below)
            if (!live) return;
            I::load_tw(r.twi, t, tw);
            if (MODE == ROW_MID || MODE == ROW_FINAL) {
                F::load_tw(r.twf, t, tw);
                r.ramp0 = tw[(t * shift) % Lx];
            }
            // Hermitian unpack into the inverse pass-A leg (bins t + q*NC)
            const cplx<T>* lo = src + ((size_t)(t / C) * Nye + y) * C + 2 * (t % C);
            const int tm = Lx - t;
            LSTED_UNROLL
            for (int q = 0; q < I::RA; ++q) {
                const int i = t + q * P::NC;
                bool upper = q > P::QH;
                if (q == P::QH) upper = 2 * i > Lx;
                const cplx<T>* p;
                if (q < P::QH || (q == P::QH && !upper)) {
                    p = lo + (size_t)(q * (P::NC / C)) * xb_stride;
                } else {
                    const int k = tm - q * P::NC;
                    p = src + ((size_t)(k / C) * Nye + y) * C + 2 * (k % C);
                }
                cplx<T> A, B;
                load_pair(p, A, B);
                if (!two) B = mk<T>(0, 0);
                if ((q == 0 && t == 0) || 2 * i == Lx) { A.y = 0; B.y = 0; }
                r.v[q] = upper ? mk<T>(A.x + B.y, B.x - A.y) : mk<T>(A.x - B.y, A.y + B.x);
            }
            I::pass_a(r.v, t, s0);
            if (MODE == ROW_MID || MODE == ROW_FINAL) async_copy_wait_all();   // visible after the barrier
        });
        // inverse pass C and the pointwise step on registers (logical position idx = t + q*NC'
        // holds pixel idx - shift of rows y (re) and y+1 (im))
        T* out = (MODE == ROW_FINAL) ? a.real_out : a.real_out + real_off;
        T* out2 = (MODE == ROW_INV_SIM) ? a.real_out2 + real_off : 0;
        cx.phase(regs, [&](int tid, R& r) {
            LSTED_ROW2_IDS
            if (!live) return;
            I::pass_c(r.v, t, s0, r.twi);
            if (t < I::NC) {
                LSTED_UNROLL
                for (int q = 0; q < I::RC; ++q) {
                    const int i = t + q * I::NC - shift;
                    cplx<T> z = r.v[q];
                    cplx<T> w = mk<T>(0, 0);
                    if (i >= 0 && i < Nx) {
                        const size_t o = (size_t)y * Nx + i;
                        if (MODE == ROW_INV_STORE) {
                            if (a.clip) { z.x = clip0(z.x); z.y = clip0(z.y); }
                            out[o] = a.accumulate ? out[o] + z.x : z.x;
                            if (two) out[o + Nx] = a.accumulate ? out[o + Nx] + z.y : z.y;
                        } else if (MODE == ROW_INV_SIM) {
                            out[o] = clip0(z.x);
                            if (two) out[o + Nx] = clip0(z.y);
                        } else if (MODE == ROW_MID) {
                            w.x = rl_ratio<true>(stage[i], z.x);
                            if (two) w.y = rl_ratio<true>(stage[P::L + i], z.y);
                        } else {  // ROW_FINAL
                            w.x = stage2[i] * fast_div(clip0(z.x), stage[i]);
                            out[o] = w.x;
                            if (two) {
                                w.y = stage2[P::L + i] * fast_div(clip0(z.y), stage[P::L + i]);
                                out[o + Nx] = w.y;
                            }
                        }
                    }
                    r.v[q] = w;
                }
            }
        });
        if (MODE == ROW_INV_SIM) {
            cx.phase(regs, [&](int tid, R& r) {
                LSTED_ROW2_IDS
                (void)r;
                if (!live || t >= I::NC) return;
                int* const queue = (int*)stage;
                const int nr = two ? 2 : 1;
                LSTED_NOUNROLL
                for (int e = 0; e < I::RC * 2; ++e) {
                    const int rr = e & 1, q = e >> 1;
                    const int i = t + q * I::NC - shift;
                    if (rr < nr && i >= 0 && i < Nx) {
                        const size_t o = (size_t)(y + rr) * Nx + i;
                        const double lam = (double)out[o];
                        double k;
                        if (!(lam >= 10.0)) {
                            out2[o] = (T)(poisson_sample(lam, a.seed, o, a.img0 + img) + 1e-9);
                        } else if (poisson_fast_ptrs(lam, a.seed, o, a.img0 + img, k)) {
                            out2[o] = (T)(k + 1e-9);
                        } else {
                            queue[1 + smem_counter_next(queue)] = rr * Nx + i;
                        }
                    }
                }
            });
            cx.phase_nosync(regs, [&](int tid, R& r) {
                LSTED_ROW2_IDS
                (void)r;
                if (!live) return;
                const int* const queue = (const int*)stage;
                const int n = queue[0];
                LSTED_NOUNROLL
                for (int e = t; e < n; e += P::NT) {
                    const size_t o = (size_t)y * Nx + queue[1 + e];
                    out2[o] = (T)(poisson_sample((double)out[o], a.seed, o, a.img0 + img) + 1e-9);
                }
            });
        }
        if (MODE == ROW_INV_STORE || MODE == ROW_INV_SIM) return;
        // every thread is done reading the exchange buffer: forward pass A may overwrite it
        cx.phase(regs, [&](int tid, R& r) {
            LSTED_ROW2_IDS
            if (!live) return;
            F::pass_a(r.v, t, s0);
        });
    }
    // Forward pass C; the upper half of the spectrum goes to the mirror thread through the
    // staging buffer (its rows were consumed by the pointwise step)
    cx.phase(regs, [&](int tid, R& r) {
        LSTED_ROW2_IDS
        if (!live) return;
        F::pass_c(r.v, t, s0, r.twf);
        LSTED_UNROLL
        for (int q = P::QH; q < P::RCF; ++q) sx[(q - P::QH) * P::PX + t] = r.v[q];
    });
    cplx<T>* dst = a.spec_out + spec_off;
    cx.phase_nosync(regs, [&](int tid, R& r) {
        LSTED_ROW2_IDS
        if (!live) return;
        const int tp = t == 0 ? 0 : P::NC - t;
        const int qoff = t == 0 ? 1 : 0;
        cplx<T>* pk = dst + ((size_t)(t / C) * Nye + y) * C + 2 * (t % C);
        cplx<T> dq[P::QH + 2];
        if (shift) F::template twiddle_powers<P::QH>(tw[(P::NC * shift) % Lx], dq);
        LSTED_UNROLL
        for (int q = 0; q <= P::QH; ++q) {
            const int k = t + q * P::NC;
            if (q == P::QH && 2 * k > Lx) {
                if (k < g.nxb * C)
                    store_pair(pk + (size_t)(q * (P::NC / C)) * xb_stride, mk<T>(0, 0), mk<T>(0, 0));
                continue;
            }
            const cplx<T> z1 = r.v[q];
            cplx<T> z2;
            const int qm = P::RCF - 1 - q + qoff;
            if (t == 0 && q == 0) z2 = z1;
            else z2 = sx[(qm - P::QH) * P::PX + tp];
            cplx<T> oa = mk<T>((T)0.5 * (z1.x + z2.x), (T)0.5 * (z1.y - z2.y));
            cplx<T> ob = mk<T>((T)0.5 * (z1.y + z2.y), (T)0.5 * (z2.x - z1.x));
            if (shift) {
                const cplx<T> ph = conj(q == 0 ? r.ramp0 : r.ramp0 * dq[q]);
                oa = oa * ph;
                ob = ob * ph;
            }
            store_pair(pk + (size_t)(q * (P::NC / C)) * xb_stride, oa, two ? ob : mk<T>(0, 0));
        }
    });
#undef LSTED_ROW2_IDS
}

}  // namespace lsted
