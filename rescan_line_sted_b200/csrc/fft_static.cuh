// Compile-time three-pass Stockham FFT with register-resident butterflies.
//
// L = RA*RB*RC.  NT threads cooperate on one sequence; thread t owns the
// butterflies j = t + m*NT of every pass and keeps their operands in
// registers, so a transform costs two shared-memory exchanges (after pass A
// and after pass B) instead of one round trip per pass plus staging:
//
//   pass A: in  logical  j + q*NA          (caller fills registers: straight
//           out position q*PA + j           from global memory, coalesced)
//   pass B: in  position (j%RA)*PA + j/RA + q*RC,  twiddle w^(RC*(j%RA)*q)
//           out position q*PB + j
//   pass C: in  position (j/RA)*PB + q*RA + j%RA,  twiddle w^(q*j)
//           out logical  j + q*NC           (stays in registers, natural order)
//
// The "[q][j]" exchange layouts make every shared-memory store unit-stride
// and, with PA odd and PB = 0 (mod 16), the loads conflict-free for RA = 16
// (odd pitches and at most 2-way conflicts otherwise).
// The inverse transform uses the reversed radix order, so the registers a
// forward pass C leaves (logical j + q*NC) are exactly the operands inverse
// pass A wants: pointwise products with an OTF happen in registers.
#pragma once
#include "fft_core.cuh"

namespace lsted {

constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
constexpr int imax(int a, int b) { return a > b ? a : b; }
// smallest p >= n with p % 16 == r % 16
constexpr int pitch_congruent(int n, int r) {
    int p = n;
    while ((p - r) % 16 != 0) ++p;
    return p;
}

// The base twiddles of pass B are tw[RC * k], k = j % RA.  In a dense table in shared memory
// the RC = 16 ones sit 128 bytes apart -- all in one bank group: a 15-way conflict, 16
// wavefronts per load (ncu: 14 % of all shared-memory wavefronts of col_h).  ColTwTable keeps
// them a second time, gathered: b<RC>(k) at p[N + 16 * (RC == RC1) + k], consecutive k in
// consecutive banks; tw[j] (16 consecutive j per warp) stays in the dense part p[0 .. N).
template <class W, int N_, int RC0_, int RC1_> struct ColTwTable {
    enum { N = N_, RC0 = RC0_, RC1 = RC1_, ELEMS = N_ + 32 };
    const W* p;
    LSTED_HD const W& operator[](int i) const { return p[i]; }
    template <int RC> LSTED_HD const W& b(int k) const {
        static_assert(RC == RC0_ || RC == RC1_, "no gathered table for this stride");
        return p[N_ + (RC == RC0_ ? 0 : 16) + k];
    }
    // fill from the global table (entry i of ELEMS)
    static LSTED_HD int source_index(int i) {
        return i < N_ ? i : i < N_ + 16 ? RC0_ * (i - N_) : RC1_ * (i - N_ - 16);
    }
};
template <int RC, class W> LSTED_HD W tw_base_b(const W* tw, int k) { return tw[RC * k]; }
template <int RC, class W, int N, int RC0, int RC1>
LSTED_HD W tw_base_b(const ColTwTable<W, N, RC0, RC1>& tw, int k) { return tw.template b<RC>(k); }

template <class V, int DIR, int RA_, int RB_, int RC_, int NT_> struct Fft3E {
    typedef typename ScalarOf<V>::type T;
    typedef cplx<T> W;   // twiddles are always single complex numbers
    enum {
        RA = RA_, RB = RB_, RC = RC_, NT = NT_,
        L = RA * RB * RC,
        NA = RB * RC, NB = RA * RC, NC = RA * RB,   // butterflies per pass
        MA = ceil_div(NA, NT), MB = ceil_div(NB, NT), MC = ceil_div(NC, NT),
        // exchange pitches (brute-forced over the access patterns of load_b / pass_c,
        // scripts/smem_bank_conflicts.py): RA = 16 is conflict-free with PA odd, PB = 0 (mod 16);
        // RA = 15 wants PB = 15 (mod 16) (and PA = 7 (mod 16) for 8-byte elements: also conflict-free
        // for the 4 or 2 interleaved sequences of a column CTA; 3 (mod 16) conflicts with 2)
        PA = (RA % 16 == 15 && NA % 16 == 0 && sizeof(V) == 8) ? NA + 7 : ((NA % 2) ? NA : NA + 1),
        PB = (RA % 16 == 0) ? pitch_congruent(NB, 0)
           : (RA % 16 == 15) ? pitch_congruent(NB, 15) : ((NB % 2) ? NB : NB + 1),
        SEQ = imax(imax(RA * PA, RB * PB), L),      // shared elements per sequence
        VREG = imax(imax(MA * RA, MB * RB), MC * RC)
    };

    // Pass A on registers v[m*RA + q] = x[j + q*NA], j = t + m*NT.
    static LSTED_HD void pass_a(V* v, int t, V* sm) {
        LSTED_UNROLL
        for (int m = 0; m < MA; ++m) {
            const int j = t + m * NT;
            if (j < NA) {
                DftE<RA, DIR, V>::run(v + m * RA);
                LSTED_UNROLL
                for (int q = 0; q < RA; ++q) sm[q * PA + j] = v[m * RA + q];
            }
        }
    }
    // w[q] = w1^q for q = 1..N by binary splitting (product depth <= log2 N, so the
    // rounding error stays at a few ulp); replaces N-1 scattered table look-ups,
    // which cost up to 32 L1 wavefronts each, by one coalesced load + N-1 products.
    template <int N> static LSTED_HD void twiddle_powers(W w1, W* w) {
        w[1] = w1;
        LSTED_UNROLL
        for (int q = 2; q <= N; ++q) w[q] = w[q / 2] * w[q - q / 2];
    }
    // Base twiddles of one thread (they depend on t only): loaded once per kernel, not
    // once per transform -- no global load sits at the head of a pass any more.
    struct Tw { W b[(NT_ % RA_ == 0) ? 1 : MB]; W c[MC]; };
    // TWP: the twiddle table -- a pointer, or a ColTwTable (above)
    template <class TWP> static LSTED_HD void load_tw(Tw& w, int t, TWP tw) {
        LSTED_UNROLL
        for (int m = 0; m < ((NT % RA == 0) ? 1 : (int)MB); ++m) {
            const int j = t + m * NT;
            w.b[m] = tw_base_b<RC>(tw, (j < NB) ? j % RA : 0);
        }
        LSTED_UNROLL
        for (int m = 0; m < MC; ++m) {
            const int j = t + m * NT;
            w.c[m] = tw[(j < NC) ? j : 0];
        }
    }
    static LSTED_HD void load_b(V* v, int t, const V* sm, const Tw& bw) {
        // every butterfly of this thread has the same k = j % RA when NT % RA == 0
        W w[RB];
        if (NT % RA == 0) twiddle_powers<RB - 1>(bw.b[0], w);
        LSTED_UNROLL
        for (int m = 0; m < MB; ++m) {
            const int j = t + m * NT;
            if (j < NB) {
                const int k = j % RA;
                const int pos = k * PA + j / RA;
                if (NT % RA != 0) twiddle_powers<RB - 1>(bw.b[(NT % RA != 0) ? m : 0], w);
                LSTED_UNROLL
                for (int q = 0; q < RB; ++q) {
                    V x = sm[pos + q * RC];
                    if (q > 0) x = mul_tw<DIR>(x, w[q]);
                    v[m * RB + q] = x;
                }
            }
        }
    }
    template <class TWP> static LSTED_HD void load_b(V* v, int t, const V* sm, TWP tw) {
        Tw bw;
        load_tw(bw, t, tw);
        load_b(v, t, sm, bw);
    }
    static LSTED_HD void pass_b(V* v, int t, V* sm) {
        LSTED_UNROLL
        for (int m = 0; m < MB; ++m) {
            const int j = t + m * NT;
            if (j < NB) {
                DftE<RB, DIR, V>::run(v + m * RB);
                LSTED_UNROLL
                for (int q = 0; q < RB; ++q) sm[q * PB + j] = v[m * RB + q];
            }
        }
    }
    // Loads pass-C operands, applies twiddles and the butterflies; on return
    // v[m*RC + q] = X[j + q*NC].
    static LSTED_HD void pass_c(V* v, int t, const V* sm, const Tw& bw) {
        LSTED_UNROLL
        for (int m = 0; m < MC; ++m) {
            const int j = t + m * NT;
            if (j < NC) {
                const int pos = (j / RA) * PB + (j % RA);
                W w[RC];
                twiddle_powers<RC - 1>(bw.c[m], w);
                LSTED_UNROLL
                for (int q = 0; q < RC; ++q) {
                    V x = sm[pos + q * RA];
                    if (q > 0) x = mul_tw<DIR>(x, w[q]);
                    v[m * RC + q] = x;
                }
                DftE<RC, DIR, V>::run(v + m * RC);
            }
        }
    }
    template <class TWP> static LSTED_HD void pass_c(V* v, int t, const V* sm, TWP tw) {
        Tw bw;
        load_tw(bw, t, tw);
        pass_c(v, t, sm, bw);
    }
};
// Two-pass variant: L = RA*RC, ONE shared-memory exchange and ONE twiddle stage per
// transform.  max(RA, RC) threads cooperate on a sequence; thread t owns butterfly t of
// each pass and RA (RC) operands in registers:
//   pass A: in  logical  t + q*NA  (NA = RC butterflies of radix RA), out position q*PA + t
//   pass C: in  position t*PA + q  (NC = RA butterflies of radix RC), twiddle w^(q*t),
//           out logical  t + q*NC  (natural order, stays in registers)
// As for Fft3E the inverse uses the reversed radix order, so forward pass-C results are the
// inverse pass-A operands.  PA odd: the strided pass-C loads are conflict-free.
template <class V, int DIR, int RA_, int RC_, int NT_> struct Fft2E {
    typedef typename ScalarOf<V>::type T;
    typedef cplx<T> W;
    enum {
        RA = RA_, RC = RC_, NT = NT_,
        L = RA * RC,
        NA = RC, NC = RA,
        PA = (NA % 2) ? NA : NA + 1,
        SEQ = imax(RA * PA, L),
        VREG = imax(RA, RC)
    };
    static_assert(NA <= NT_ && NC <= NT_, "one butterfly per thread and pass");
    struct Tw { W c[1]; };
    static LSTED_HD void load_tw(Tw& w, int t, const W* tw) { w.c[0] = tw[(t < NC) ? t : 0]; }
    template <int N> static LSTED_HD void twiddle_powers(W w1, W* w) {
        w[1] = w1;
        LSTED_UNROLL
        for (int q = 2; q <= N; ++q) w[q] = w[q / 2] * w[q - q / 2];
    }
    static LSTED_HD void pass_a(V* v, int t, V* sm) {
        if (t < NA) {
            DftE<RA, DIR, V>::run(v);
            LSTED_UNROLL
            for (int q = 0; q < RA; ++q) sm[q * PA + t] = v[q];
        }
    }
    static LSTED_HD void pass_c(V* v, int t, const V* sm, const Tw& bw) {
        if (t < NC) {
            // operand q of butterfly t is output t of pass-A butterfly q
            W w[RC];
            twiddle_powers<RC - 1>(bw.c[0], w);
            LSTED_UNROLL
            for (int q = 0; q < RC; ++q) {
                V x = sm[t * PA + q];
                if (q > 0) x = mul_tw<DIR>(x, w[q]);
                v[q] = x;
            }
            DftE<RC, DIR, V>::run(v);
        }
    }
};

template <typename T, int DIR, int RA, int RB, int RC, int NT> struct Fft3 : Fft3E<cplx<T>, DIR, RA, RB, RC, NT> {};

}  // namespace lsted
