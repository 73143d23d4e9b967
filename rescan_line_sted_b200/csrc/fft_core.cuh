// Shared-memory Stockham FFT building blocks (mixed radix 2/3/4/5/8/9/16).
//
// Everything in this header is plain C++ that compiles both as sm_100a
// device code (nvcc) and as host code (g++): the kernels are written as
// sequences of "CTA-wide parallel phases" over a work-item index, so the
// very same bodies can be replayed serially on the CPU by the test-only
// emulator in tests/host_emul/ (no product path uses the host build).
//
// Replaces, for the hot path of SURVEY.md section 8a, the pocketfft/ducc
// r2c/c2r transforms that scipy.signal.fftconvolve runs inside
// figure_generation/line_sted_tools.py:574 and :586.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LSTED_HD __host__ __device__ __forceinline__
#define LSTED_UNROLL _Pragma("unroll")
#define LSTED_NOUNROLL _Pragma("unroll 1")
#else
#define LSTED_HD inline
#define LSTED_UNROLL
#define LSTED_NOUNROLL
#endif

namespace lsted {

template <typename T> struct alignas(2 * sizeof(T)) cplx { T x, y; };

template <typename T> LSTED_HD cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename T> LSTED_HD cplx<T> operator+(cplx<T> a, cplx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> LSTED_HD cplx<T> operator-(cplx<T> a, cplx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> LSTED_HD cplx<T> operator*(cplx<T> a, cplx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> LSTED_HD cplx<T> scale(cplx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// a * s + b with a real factor s
template <typename T> LSTED_HD cplx<T> fma_real(cplx<T> a, T s, cplx<T> b) {
    return mk<T>(a.x * s + b.x, a.y * s + b.y);
}
#ifndef LSTED_NO_PACKED_F32X2
#define LSTED_PACKED_F32X2 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && defined(LSTED_PACKED_F32X2)
// Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2): a complex add, subtract or
// real scaling is one instruction on a 64-bit register pair.
LSTED_HD cplx<float> operator+(cplx<float> a, cplx<float> b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return mk<float>(r.x, r.y);
}
LSTED_HD cplx<float> operator-(cplx<float> a, cplx<float> b) {
    const float2 r = __ffma2_rn(make_float2(b.x, b.y), make_float2(-1.f, -1.f), make_float2(a.x, a.y));
    return mk<float>(r.x, r.y);
}
LSTED_HD cplx<float> scale(cplx<float> a, float s) {
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(s, s));
    return mk<float>(r.x, r.y);
}
LSTED_HD cplx<float> fma_real(cplx<float> a, float s, cplx<float> b) {
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(s, s), make_float2(b.x, b.y));
    return mk<float>(r.x, r.y);
}
#endif
template <typename T> LSTED_HD cplx<T> conj(cplx<T> a) { return mk<T>(a.x, -a.y); }
// Branch-free fp32 division (MUFU.RCP based, <= 2 ulp) so the compiler can keep
// many independent loads/divisions in flight; fp64 divides exactly.
LSTED_HD float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
LSTED_HD double fast_div(double a, double b) { return a / b; }
// Pull one cache line towards L2 ahead of use (no register, no dependency).
LSTED_HD void prefetch_l2(const void* p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
// Asynchronous global -> shared copies (LDGSTS / cp.async): no registers held while
// the data is in flight.  `async_copy_row` spreads n elements over the threads of a
// group, 16 bytes per request when source and destination allow it.
template <typename T>
LSTED_HD void async_copy_row(T* dst_smem, const T* src, int n, int tid, int nthreads) {
#ifdef __CUDA_ARCH__
    const int per16 = 16 / (int)sizeof(T);
    const bool wide = ((((size_t)src) | ((size_t)dst_smem)) & 15) == 0 && (n % per16) == 0;
    if (wide) {
        for (int i = tid * per16; i < n; i += nthreads * per16) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem + i);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i));
        }
    } else {
        for (int i = tid; i < n; i += nthreads) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem + i);
            if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src + i));
            else                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + i));
        }
    }
    asm volatile("cp.async.commit_group;");
#else
    for (int i = tid; i < n; i += nthreads) dst_smem[i] = src[i];
#endif
}
LSTED_HD void async_copy_wait_all() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// `nthreads` threads sweep [p, p + bytes) in 128-byte lines
LSTED_HD void prefetch_l2_range(const void* p, size_t bytes, int tid, int nthreads) {
    const char* c = (const char*)p;
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthreads * 128) prefetch_l2(c + off);
}
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR, typename T> LSTED_HD cplx<T> mul_dir_i(cplx<T> a) {
    return DIR < 0 ? mk<T>(a.y, -a.x) : mk<T>(-a.y, a.x);
}
// a * (c + DIR*i*s) where (c, s) = (cos t, sin t), t >= 0: forward uses exp(-it)
template <int DIR, typename T> LSTED_HD cplx<T> mul_cs(cplx<T> a, T c, T s) {
    return DIR < 0 ? mk<T>(a.x * c + a.y * s, a.y * c - a.x * s)
                   : mk<T>(a.x * c - a.y * s, a.y * c + a.x * s);
}
// twiddle from a forward table entry w = exp(-i t)
template <int DIR, typename T> LSTED_HD cplx<T> mul_tw(cplx<T> a, cplx<T> w) {
    return DIR < 0 ? a * w : a * conj(w);
}

// ---------------------------------------------------------------------------
// In-register DFTs, natural order in, natural order out.
// ---------------------------------------------------------------------------
template <int R, int DIR, typename T> struct Dft;

template <int DIR, typename T> struct Dft<2, DIR, T> {
    static LSTED_HD void run(cplx<T>* v) {
        cplx<T> a = v[0];
        v[0] = a + v[1];
        v[1] = a - v[1];
    }
};

template <int DIR, typename T> struct Dft<3, DIR, T> {
    static LSTED_HD void run(cplx<T>* v) {
        const T s3 = (T)0.86602540378443864676;
        cplx<T> t = v[1] + v[2];
        cplx<T> d = mul_dir_i<DIR>(scale(v[1] - v[2], s3));
        cplx<T> m = fma_real(t, (T)-0.5, v[0]);
        v[0] = v[0] + t;
        v[1] = m + d;
        v[2] = m - d;
    }
};

template <int DIR, typename T> struct Dft<4, DIR, T> {
    static LSTED_HD void run(cplx<T>* v) {
        cplx<T> a = v[0] + v[2], b = v[0] - v[2];
        cplx<T> c = v[1] + v[3], d = mul_dir_i<DIR>(v[1] - v[3]);
        v[0] = a + c;
        v[1] = b + d;
        v[2] = a - c;
        v[3] = b - d;
    }
};

template <int DIR, typename T> struct Dft<5, DIR, T> {
    static LSTED_HD void run(cplx<T>* v) {
        const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
        cplx<T> t1 = v[1] + v[4], t2 = v[2] + v[3];
        cplx<T> t3 = v[1] - v[4], t4 = v[2] - v[3];
        cplx<T> a1 = fma_real(t2, c2, fma_real(t1, c1, v[0]));
        cplx<T> a2 = fma_real(t2, c1, fma_real(t1, c2, v[0]));
        cplx<T> b1 = mul_dir_i<DIR>(fma_real(t4, s2, scale(t3, s1)));
        cplx<T> b2 = mul_dir_i<DIR>(fma_real(t4, -s1, scale(t3, s2)));
        v[0] = v[0] + t1 + t2;
        v[1] = a1 + b1;
        v[4] = a1 - b1;
        v[2] = a2 + b2;
        v[3] = a2 - b2;
    }
};

// cos/sin of 2*pi*m/R for the composite radices (compile-time folded after
// unrolling: m is always a constant expression at the call sites).
template <int R, typename T> LSTED_HD void unit_root(int m, T& c, T& s) {
    if (R == 8) {
        const T h = (T)0.70710678118654752440;
        const T C[8] = {1, h, 0, -h, -1, -h, 0, h};
        const T S[8] = {0, h, 1, h, 0, -h, -1, -h};
        c = C[m & 7]; s = S[m & 7];
    } else if (R == 9) {
        const T C[9] = {(T)1, (T)0.76604444311897803520, (T)0.17364817766693034885, (T)-0.5,
                        (T)-0.93969262078590838405, (T)-0.93969262078590838405, (T)-0.5,
                        (T)0.17364817766693034885, (T)0.76604444311897803520};
        const T S[9] = {(T)0, (T)0.64278760968653932632, (T)0.98480775301220805937,
                        (T)0.86602540378443864676, (T)0.34202014332566873304,
                        (T)-0.34202014332566873304, (T)-0.86602540378443864676,
                        (T)-0.98480775301220805937, (T)-0.64278760968653932632};
        c = C[m % 9]; s = S[m % 9];
    } else {  // 16
        const T a = (T)0.92387953251128675613, b = (T)0.38268343236508977173;
        const T h = (T)0.70710678118654752440;
        const T C[16] = {1, a, h, b, 0, -b, -h, -a, -1, -a, -h, -b, 0, b, h, a};
        const T S[16] = {0, b, h, a, 1, a, h, b, 0, -b, -h, -a, -1, -a, -h, -b};
        c = C[m & 15]; s = S[m & 15];
    }
}

// R = R1*R2 by one in-register Cooley-Tukey step.
template <int R, int R1, int R2, int DIR, typename T> struct DftComposite {
    static LSTED_HD void run(cplx<T>* v) {
        cplx<T> y[R2][R1];
        LSTED_UNROLL
        for (int n2 = 0; n2 < R2; ++n2) {
            LSTED_UNROLL
            for (int n1 = 0; n1 < R1; ++n1) y[n2][n1] = v[n2 + R2 * n1];
            Dft<R1, DIR, T>::run(y[n2]);
            LSTED_UNROLL
            for (int k1 = 1; k1 < R1; ++k1) {
                if (n2 * k1 != 0) {
                    T c, s;
                    unit_root<R, T>(n2 * k1, c, s);
                    y[n2][k1] = mul_cs<DIR>(y[n2][k1], c, s);
                }
            }
        }
        LSTED_UNROLL
        for (int k1 = 0; k1 < R1; ++k1) {
            cplx<T> z[R2];
            LSTED_UNROLL
            for (int n2 = 0; n2 < R2; ++n2) z[n2] = y[n2][k1];
            Dft<R2, DIR, T>::run(z);
            LSTED_UNROLL
            for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = z[k2];
        }
    }
};
template <int DIR, typename T> struct Dft<8, DIR, T> : DftComposite<8, 4, 2, DIR, T> {};
template <int DIR, typename T> struct Dft<9, DIR, T> : DftComposite<9, 3, 3, DIR, T> {};
template <int DIR, typename T> struct Dft<16, DIR, T> : DftComposite<16, 4, 4, DIR, T> {};

// ---------------------------------------------------------------------------
// Plans
// ---------------------------------------------------------------------------
enum { kMaxPasses = 12 };
struct FftPlan {
    int L;                   // transform length
    int npass;
    int radix[kMaxPasses];   // product == L
};

// Shared-memory index padding: one spare slot every 16 (8-byte) or 8
// (16-byte) elements, so power-of-two strides spread over all banks.
template <typename T> struct PadShift { enum { value = sizeof(T) == 4 ? 4 : 3 }; };
template <typename T> LSTED_HD int pad(int i) { return i + (i >> PadShift<T>::value); }

// Source functors for the first pass of a transform: where element i of
// sequence f comes from.  SmemSrc reads the padded shared-memory layout.
template <typename T> struct SmemSrc {
    const cplx<T>* base;
    int Lp;
    LSTED_HD cplx<T> operator()(int f, int i) const { return base[f * Lp + pad<T>(i)]; }
};

// One Stockham butterfly of radix R: reads the source at stride L/R, writes
// dst in auto-sorted order.  `tw` is the forward table exp(-2*pi*i*m/L).
template <int R, int DIR, typename T, class Src>
LSTED_HD void stockham_butterfly(const Src& src, cplx<T>* dst, int f, int j, int LR, int Ns,
                                 int tw_stride, const cplx<T>* tw) {
    cplx<T> v[R];
    const int k = j % Ns;
    LSTED_UNROLL
    for (int r = 0; r < R; ++r) v[r] = src(f, j + r * LR);
    if (Ns > 1) {
        const int step = k * tw_stride;
        LSTED_UNROLL
        for (int r = 1; r < R; ++r) v[r] = mul_tw<DIR>(v[r], tw[r * step]);
    }
    Dft<R, DIR, T>::run(v);
    const int base = (j - k) * R + k;
    LSTED_UNROLL
    for (int r = 0; r < R; ++r) dst[pad<T>(base + r * Ns)] = v[r];
}

template <int R, int DIR, typename T, class Ctx, class Src>
LSTED_HD void stockham_pass(Ctx& cx, const Src& src, cplx<T>* dst, int nbatch, int Lp, int L,
                            int Ns, const cplx<T>* tw) {
    const int LR = L / R;
    const int tw_stride = L / (Ns * R);
    cx.parallel_for(nbatch * LR, [&](int w) {
        const int f = w / LR;
        const int j = w - f * LR;
        stockham_butterfly<R, DIR, T>(src, dst + f * Lp, f, j, LR, Ns, tw_stride, tw);
    });
}

template <int DIR, typename T, class Ctx, class Src>
LSTED_HD void stockham_dispatch(Ctx& cx, int R, const Src& src, cplx<T>* dst, int nbatch, int Lp,
                                int L, int Ns, const cplx<T>* tw) {
    switch (R) {
        case 2: stockham_pass<2, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        case 3: stockham_pass<3, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        case 4: stockham_pass<4, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        case 5: stockham_pass<5, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        case 8: stockham_pass<8, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        case 9: stockham_pass<9, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
        default: stockham_pass<16, DIR, T>(cx, src, dst, nbatch, Lp, L, Ns, tw); break;
    }
}

// Batched FFT of `nbatch` sequences.  The first pass pulls its input through
// `first` (which may live in a third buffer, or fuse a pointwise product);
// later passes ping-pong between the shared-memory buffers w1 and w2 laid
// out as w[f*Lp + pad(i)].  Returns the buffer that holds the result.
// `first` may alias w2 but not w1.
template <int DIR, typename T, class Ctx, class Src>
LSTED_HD cplx<T>* fft_batch(Ctx& cx, const FftPlan& plan, const cplx<T>* tw, const Src& first,
                            cplx<T>* w1, cplx<T>* w2, int nbatch, int Lp) {
    stockham_dispatch<DIR, T>(cx, plan.radix[0], first, w1, nbatch, Lp, plan.L, 1, tw);
    int Ns = plan.radix[0];
    for (int p = 1; p < plan.npass; ++p) {
        SmemSrc<T> s = {w1, Lp};
        stockham_dispatch<DIR, T>(cx, plan.radix[p], s, w2, nbatch, Lp, plan.L, Ns, tw);
        Ns *= plan.radix[p];
        cplx<T>* t = w1; w1 = w2; w2 = t;
    }
    return w1;
}

}  // namespace lsted
