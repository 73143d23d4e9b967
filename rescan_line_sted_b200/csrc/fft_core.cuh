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

// Range facts the compiler cannot derive (a thread index that survived its `live` guard):
// on the device they fold the per-bin tests of the unrolled loops away.
#ifdef __CUDA_ARCH__
#define LSTED_ASSUME(c) __builtin_assume(c)
#else
#define LSTED_ASSUME(c) ((void)0)
#endif

// Debug build (make debug -> liblsted_debug.so, -DLSTED_DEBUG): in-kernel checks of every
// tensor-map box, bulk-copy range / alignment and crop index.  compute-sanitizer is not
// available on the GPU pool, and an out-of-range TMA box once produced silent garbage
// (DESIGN.md section 4); a failed check prints its location and traps the kernel, which the
// next API call reports as a CUDA error.
#ifdef LSTED_DEBUG
#include <stdio.h>
#ifdef __CUDA_ARCH__
#define LSTED_DCHECK(cond)                                                                      \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("LSTED_DCHECK failed: %s  at %s:%d (block %d, thread %d)\n", #cond, __FILE__, \
                   __LINE__, (int)blockIdx.x, (int)threadIdx.x);                                \
            __trap();                                                                           \
        }                                                                                       \
    } while (0)
#else
#include <stdlib.h>
#define LSTED_DCHECK(cond)                                                                      \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            fprintf(stderr, "LSTED_DCHECK failed: %s  at %s:%d\n", #cond, __FILE__, __LINE__);  \
            abort();                                                                            \
        }                                                                                       \
    } while (0)
#endif
#else
#define LSTED_DCHECK(cond) do { } while (0)
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
#ifndef LSTED_NO_BULK_PREFETCH
#define LSTED_BULK_PREFETCH 1
#endif
#ifndef LSTED_NO_PACKED_F32X2
#define LSTED_PACKED_F32X2 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && defined(LSTED_PACKED_F32X2)
// Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2) on the 64-bit register pair that
// holds one complex number.  The operands of these instructions take a lane swap
// (.LO_HI), a per-lane sign (.NP / .PN) and a scalar broadcast (R.F32) for free, so
//   complex add / subtract / real scaling       = 1 instruction,
//   complex x complex, complex x (cos, sin)     = 2 instructions (FMUL2 + FFMA2),
//   multiplication by +-i                       = 0 (folded into the consumer).
LSTED_HD cplx<float> operator+(cplx<float> a, cplx<float> b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return mk<float>(r.x, r.y);
}
LSTED_HD cplx<float> operator-(cplx<float> a, cplx<float> b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return mk<float>(r.x, r.y);
}
LSTED_HD cplx<float> operator*(cplx<float> a, cplx<float> b) {
    const float2 t = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.x));
    const float2 r = __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), t);
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
// a * b + c
template <typename T> LSTED_HD cplx<T> cmac(cplx<T> a, cplx<T> b, cplx<T> c) {
    return mk<T>(a.x * b.x - a.y * b.y + c.x, a.x * b.y + a.y * b.x + c.y);
}
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && defined(LSTED_PACKED_F32X2)
LSTED_HD cplx<float> cmac(cplx<float> a, cplx<float> b, cplx<float> c) {   // 2 x FFMA2
    const float2 t = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.x), make_float2(c.x, c.y));
    const float2 r = __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), t);
    return mk<float>(r.x, r.y);
}
#endif
// Branch-free fp32 division (MUFU.RCP based, <= 2 ulp) so the compiler can keep
// many independent loads/divisions in flight; fp64 divides exactly.
LSTED_HD float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
    // MUFU.RCP + FMUL; `.ftz` spares the denormal-input scaling the compiler otherwise wraps
    // around MUFU.RCP (4 extra instructions per division; b < 2^-126 gives inf either way)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
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
// Two adjacent complex numbers (the two rows of a row pair in the XB2 spectrum layout)
// in one access: 16 bytes for fp32 (LDG.128 / STG.128), two 16-byte ones for fp64.
template <typename T> LSTED_HD void load_pair(const cplx<T>* p, cplx<T>& a, cplx<T>& b) { a = p[0]; b = p[1]; }
template <typename T> LSTED_HD void store_pair(cplx<T>* p, cplx<T> a, cplx<T> b) { p[0] = a; p[1] = b; }
#ifdef __CUDA_ARCH__
LSTED_HD void load_pair(const cplx<float>* p, cplx<float>& a, cplx<float>& b) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = mk<float>(v.x, v.y); b = mk<float>(v.z, v.w);
}
LSTED_HD void store_pair(cplx<float>* p, cplx<float> a, cplx<float> b) {
    *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
#endif
// System-scope (cross-GPU, NVLink peer memory) synchronisation primitives of the fused
// H_t reduction (conv_fast.cuh, COL_HT with P2P): release/acquire flags and counters that
// live in one GPU's memory and are written by kernels running on its peers.
LSTED_HD void sys_fence() {
#ifdef __CUDA_ARCH__
    __threadfence_system();
#endif
}
LSTED_HD void flag_release(unsigned* p, unsigned v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    *p = v;
#endif
}
LSTED_HD unsigned flag_acquire(const unsigned* p) {
#ifdef __CUDA_ARCH__
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
// spin until the counter has reached `target` (wrap-around safe)
LSTED_HD void flag_wait(const unsigned* p, unsigned target) {
#ifdef __CUDA_ARCH__
    while ((int)(flag_acquire(p) - target) < 0) __nanosleep(64);
#else
    (void)p; (void)target;
#endif
}
LSTED_HD void sys_counter_add(unsigned* p, unsigned v) {
#ifdef __CUDA_ARCH__
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    *p += v;
#endif
}
// load that bypasses L1 (data written by a peer GPU during this kernel)
template <typename T> LSTED_HD cplx<T> load_l2(const cplx<T>* p) {
#ifdef __CUDA_ARCH__
    cplx<T> r;
    if (sizeof(T) == 4) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
        r.x = (T)v.x; r.y = (T)v.y;
    } else {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
        r.x = (T)v.x; r.y = (T)v.y;
    }
    return r;
#else
    return *p;
#endif
}
// load_pair that bypasses L1 (data written by a peer GPU during this kernel)
template <typename T> LSTED_HD void load_pair_l2(const cplx<T>* p, cplx<T>& a, cplx<T>& b) {
    a = load_l2(p); b = load_l2(p + 1);
}
#ifdef __CUDA_ARCH__
LSTED_HD void load_pair_l2(const cplx<float>* p, cplx<float>& a, cplx<float>& b) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    a = mk<float>(v.x, v.y); b = mk<float>(v.z, v.w);
}
#endif
// CTA-wide counter in shared memory (host replay: threads run one after another).  A plain
// atomic per lane: taking a warp's slots with one aggregated atomic (__activemask / popc /
// shuffle) was measured slower in the noise queues of ROW_INV_SIM (1.37 -> 1.47 ms).
LSTED_HD int smem_counter_next(int* counter) {
#ifdef __CUDA_ARCH__
    return atomicAdd(counter, 1);
#else
    return (*counter)++;
#endif
}
// Bulk asynchronous copy global -> shared (TMA engine, `cp.async.bulk`): ONE thread
// issues one instruction for a whole contiguous slab, completion is tracked by an
// mbarrier (transaction bytes); no registers and no per-thread copy instructions.
// Host build (CPU replay): a plain memcpy at issue time.
typedef unsigned long long mbar_t;
LSTED_HD void mbar_init(mbar_t* bar) {
#ifdef __CUDA_ARCH__
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    *bar = 0;
#endif
}
// bytes: multiple of 16; dst / src 16-byte aligned
LSTED_HD void bulk_load(void* dst_smem, const void* src, unsigned bytes, mbar_t* bar) {
    LSTED_DCHECK(bytes % 16 == 0 && bytes > 0);
    LSTED_DCHECK(((size_t)src & 15) == 0 && ((size_t)dst_smem & 15) == 0);
#ifdef __CUDA_ARCH__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
#else
    (void)bar;
    const char* s = (const char*)src;
    char* d = (char*)dst_smem;
    for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
#endif
}
// several bulk copies on one mbarrier phase: announce the total once, then issue the copies
LSTED_HD void bulk_expect(mbar_t* bar, unsigned total_bytes) {
#ifdef __CUDA_ARCH__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(total_bytes) : "memory");
#else
    (void)bar; (void)total_bytes;
#endif
}
LSTED_HD void bulk_copy(void* dst_smem, const void* src, unsigned bytes, mbar_t* bar) {
    LSTED_DCHECK(bytes % 16 == 0 && bytes > 0);
    LSTED_DCHECK(((size_t)src & 15) == 0 && ((size_t)dst_smem & 15) == 0);
#ifdef __CUDA_ARCH__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
#else
    (void)bar;
    const char* s = (const char*)src;
    char* d = (char*)dst_smem;
    for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
#endif
}
// Tensor-map (TMA) copies of one row pair's spectrum: the pair owns one 64-byte chunk
// (2 rows x C columns) in every column block, 64 KB apart -- a 3-D tensor
// [image][block][floats of a block slab] with a box of {floats of the chunk, kTmaBoxBlocks, 1}
// moves the nxb chunks with two instructions, none of them through the LSU.  Device: the
// descriptor `tmap` (CUtensorMap in global memory) does the addressing; host replay: the
// same box is copied with plain loops from `base` (the image's spectrum, XB2 layout).
enum { kTmaBoxBlocks = 137 };
template <typename T>
LSTED_HD void tma_load_chunks(void* dst_smem, const void* tmap, const cplx<T>* base, int img, int y, int xb0,
                              int nxb, int rows_e, int C, int chunk_cplx, mbar_t* bar) {
    // the box must lie inside the tensor: blocks [xb0, xb0 + kTmaBoxBlocks) of nxb, an even row
    // of the slab, and a 128-byte aligned shared-memory destination
    LSTED_DCHECK(xb0 >= 0 && xb0 + kTmaBoxBlocks <= nxb);
    LSTED_DCHECK(y >= 0 && (y & 1) == 0 && y < rows_e && img >= 0);
    LSTED_DCHECK(((size_t)dst_smem & 127) == 0);
#ifdef __CUDA_ARCH__
    (void)base; (void)nxb; (void)rows_e;
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int c0 = y * C * 2;   // floats from the start of the block slab (y even)
    (void)chunk_cplx;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(d), "l"(tmap), "r"(b), "r"(c0), "r"(xb0), "r"(img) : "memory");
#else
    (void)tmap; (void)bar; (void)img;
    cplx<T>* d = (cplx<T>*)dst_smem;
    for (int i = 0; i < kTmaBoxBlocks; ++i)
        for (int e = 0; e < chunk_cplx; ++e) {
            const int xb = xb0 + i;
            d[(size_t)i * chunk_cplx + e] =
                xb < nxb ? base[((size_t)xb * rows_e + y) * C + e] : mk<T>(0, 0);
        }
#endif
}
template <typename T>
LSTED_HD void tma_store_chunks(const void* src_smem, const void* tmap, cplx<T>* base, int img, int y, int xb0,
                               int nxb, int rows_e, int C, int chunk_cplx) {
    LSTED_DCHECK(xb0 >= 0 && xb0 + kTmaBoxBlocks <= nxb);
    LSTED_DCHECK(y >= 0 && (y & 1) == 0 && y < rows_e && img >= 0);
    LSTED_DCHECK(((size_t)src_smem & 127) == 0);
#ifdef __CUDA_ARCH__
    (void)base; (void)nxb; (void)rows_e; (void)chunk_cplx;
    const unsigned d = (unsigned)__cvta_generic_to_shared(src_smem);
    const int c0 = y * C * 2;
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(d), "r"(c0), "r"(xb0), "r"(img) : "memory");
#else
    (void)tmap; (void)img;
    const cplx<T>* s = (const cplx<T>*)src_smem;
    for (int i = 0; i < kTmaBoxBlocks; ++i)
        for (int e = 0; e < chunk_cplx; ++e) {
            const int xb = xb0 + i;
            if (xb < nxb) base[((size_t)xb * rows_e + y) * C + e] = s[(size_t)i * chunk_cplx + e];
        }
#endif
}
LSTED_HD void tma_store_finish() {   // by the issuing thread: the bulk stores have read shared memory
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
LSTED_HD void tma_prefetch_chunks(const void* tmap, int img, int y, int xb0, int C) {
#ifdef __CUDA_ARCH__
    const int c0 = y * C * 2;
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(tmap), "r"(c0), "r"(xb0), "r"(img) : "memory");
#else
    (void)tmap; (void)img; (void)y; (void)xb0; (void)C;
#endif
}
LSTED_HD void fence_async_smem() {   // generic-proxy writes to shared memory -> visible to bulk copies
#ifdef __CUDA_ARCH__
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
LSTED_HD void mbar_wait(mbar_t* bar, unsigned parity) {
#ifdef __CUDA_ARCH__
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
    } while (!done);
#else
    (void)bar; (void)parity;
#endif
}
// Bulk L2 prefetch (`cp.async.bulk.prefetch.L2`, one instruction for a whole contiguous
// range, handled by the bulk-copy engine instead of one LSU request per 128-byte line).
// p 16-byte aligned, bytes a multiple of 16.
LSTED_HD void prefetch_l2_bulk(const void* p, unsigned bytes) {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
    (void)p; (void)bytes;
#endif
}
// one small chunk with a per-thread address: the plain line prefetch (a bulk prefetch is a
// warp-uniform instruction; with 32 different addresses the compiler serialises the lanes,
// ~6 issue slots per chunk, which costs more than the LSU request it saves)
LSTED_HD void prefetch_chunk(const void* p, unsigned bytes) {
    (void)bytes;
    prefetch_l2(p);
}
// [p, p + bytes) towards L2: split into `pieces` bulk prefetches issued by threads
// 0..pieces-1 when alignment allows, else `nthreads` threads sweep it in 128-byte lines
LSTED_HD void prefetch_l2_range(const void* p, size_t bytes, int tid, int nthreads, int pieces = 4) {
    const char* c = (const char*)p;
#ifdef LSTED_BULK_PREFETCH
    if ((((size_t)p) & 15) == 0 && bytes % ((size_t)16 * pieces) == 0 && bytes / pieces < (1u << 30)) {
        const size_t piece = bytes / pieces;
        if (tid < pieces) prefetch_l2_bulk(c + (size_t)tid * piece, (unsigned)piece);
        return;
    }
#endif
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nthreads * 128) prefetch_l2(c + off);
}
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR, typename T> LSTED_HD cplx<T> mul_dir_i(cplx<T> a) {
    return DIR < 0 ? mk<T>(a.y, -a.x) : mk<T>(-a.y, a.x);
}
// a * (c + DIR*i*s) where (c, s) = (cos t, sin t), t >= 0: forward uses exp(-it)
template <int DIR, typename T> LSTED_HD cplx<T> mul_cs(cplx<T> a, T c, T s) {
    return a * mk<T>(c, DIR < 0 ? -s : s);
}
// twiddle from a forward table entry w = exp(-i t)
template <int DIR, typename T> LSTED_HD cplx<T> mul_tw(cplx<T> a, cplx<T> w) {
    return DIR < 0 ? a * w : a * conj(w);
}

// ---------------------------------------------------------------------------
// Two complex sequences side by side ("dual" element, fp32 fast path v2): one
// thread carries the same butterfly of two independent transforms, so twiddles,
// addresses, predicates and shared-memory instructions (one 128-bit access per
// element) are shared; each half is an ordinary packed complex number.
// ---------------------------------------------------------------------------
struct alignas(16) c2 { cplx<float> a, b; };
LSTED_HD c2 mk2(cplx<float> a, cplx<float> b) { c2 r; r.a = a; r.b = b; return r; }
LSTED_HD c2 operator+(c2 u, c2 v) { return mk2(u.a + v.a, u.b + v.b); }
LSTED_HD c2 operator-(c2 u, c2 v) { return mk2(u.a - v.a, u.b - v.b); }
LSTED_HD c2 scale(c2 u, float s) { return mk2(scale(u.a, s), scale(u.b, s)); }
LSTED_HD c2 fma_real(c2 u, float s, c2 v) { return mk2(fma_real(u.a, s, v.a), fma_real(u.b, s, v.b)); }
template <int DIR> LSTED_HD c2 mul_dir_i(c2 u) { return mk2(mul_dir_i<DIR>(u.a), mul_dir_i<DIR>(u.b)); }
template <int DIR> LSTED_HD c2 mul_cs(c2 u, float c, float s) {
    return mk2(mul_cs<DIR>(u.a, c, s), mul_cs<DIR>(u.b, c, s));
}
template <int DIR> LSTED_HD c2 mul_tw(c2 u, cplx<float> w) {
    return mk2(mul_tw<DIR>(u.a, w), mul_tw<DIR>(u.b, w));
}
LSTED_HD c2 operator*(c2 u, c2 v) { return mk2(u.a * v.a, u.b * v.b); }

// scalar type behind an element type
template <class V> struct ScalarOf;
template <typename T> struct ScalarOf<cplx<T> > { typedef T type; };
template <> struct ScalarOf<c2> { typedef float type; };

// ---------------------------------------------------------------------------
// In-register DFTs, natural order in, natural order out, on any element type V
// (cplx<float>, cplx<double>, c2).
// ---------------------------------------------------------------------------
template <int R, int DIR, class V> struct DftE;

template <int DIR, class V> struct DftE<2, DIR, V> {
    static LSTED_HD void run(V* v) {
        V a = v[0];
        v[0] = a + v[1];
        v[1] = a - v[1];
    }
};

template <int DIR, class V> struct DftE<3, DIR, V> {
    typedef typename ScalarOf<V>::type T;
    static LSTED_HD void run(V* v) {
        const T s3 = (T)0.86602540378443864676;
        V t = v[1] + v[2];
        V d = mul_dir_i<DIR>(scale(v[1] - v[2], s3));
        V m = fma_real(t, (T)-0.5, v[0]);
        v[0] = v[0] + t;
        v[1] = m + d;
        v[2] = m - d;
    }
};

template <int DIR, class V> struct DftE<4, DIR, V> {
    static LSTED_HD void run(V* v) {
        V a = v[0] + v[2], b = v[0] - v[2];
        V c = v[1] + v[3], d = mul_dir_i<DIR>(v[1] - v[3]);
        v[0] = a + c;
        v[1] = b + d;
        v[2] = a - c;
        v[3] = b - d;
    }
};

template <int DIR, class V> struct DftE<5, DIR, V> {
    typedef typename ScalarOf<V>::type T;
    static LSTED_HD void run(V* v) {
        const T c1 = (T)0.30901699437494742410, c2_ = (T)-0.80901699437494742410;
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
        V t1 = v[1] + v[4], t2 = v[2] + v[3];
        V t3 = v[1] - v[4], t4 = v[2] - v[3];
        V a1 = fma_real(t2, c2_, fma_real(t1, c1, v[0]));
        V a2 = fma_real(t2, c1, fma_real(t1, c2_, v[0]));
        V b1 = mul_dir_i<DIR>(fma_real(t4, s2, scale(t3, s1)));
        V b2 = mul_dir_i<DIR>(fma_real(t4, -s1, scale(t3, s2)));
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
template <int R, int R1, int R2, int DIR, class V> struct DftComposite {
    typedef typename ScalarOf<V>::type T;
    static LSTED_HD void run(V* v) {
        V y[R2][R1];
        LSTED_UNROLL
        for (int n2 = 0; n2 < R2; ++n2) {
            LSTED_UNROLL
            for (int n1 = 0; n1 < R1; ++n1) y[n2][n1] = v[n2 + R2 * n1];
            DftE<R1, DIR, V>::run(y[n2]);
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
            V z[R2];
            LSTED_UNROLL
            for (int n2 = 0; n2 < R2; ++n2) z[n2] = y[n2][k1];
            DftE<R2, DIR, V>::run(z);
            LSTED_UNROLL
            for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = z[k2];
        }
    }
};
template <int DIR, class V> struct DftE<8, DIR, V> : DftComposite<8, 4, 2, DIR, V> {};
template <int DIR, class V> struct DftE<9, DIR, V> : DftComposite<9, 3, 3, DIR, V> {};
template <int DIR, class V> struct DftE<16, DIR, V> : DftComposite<16, 4, 4, DIR, V> {};

// Good-Thomas 15-point DFT (3 x 5, no twiddles): n = 5*n1 + 3*n2,
// k = 10*k1 + 6*k2 (mod 15).
template <int DIR, class V> struct DftE<15, DIR, V> {
    static LSTED_HD void run(V* v) {
        V y[5][3];
        LSTED_UNROLL
        for (int n2 = 0; n2 < 5; ++n2) {
            LSTED_UNROLL
            for (int n1 = 0; n1 < 3; ++n1) y[n2][n1] = v[(5 * n1 + 3 * n2) % 15];
            DftE<3, DIR, V>::run(y[n2]);
        }
        LSTED_UNROLL
        for (int k1 = 0; k1 < 3; ++k1) {
            V z[5];
            LSTED_UNROLL
            for (int n2 = 0; n2 < 5; ++n2) z[n2] = y[n2][k1];
            DftE<5, DIR, V>::run(z);
            LSTED_UNROLL
            for (int k2 = 0; k2 < 5; ++k2) v[(10 * k1 + 6 * k2) % 15] = z[k2];
        }
    }
};

// R = R1*R2 with coprime factors by the Good-Thomas map (no internal twiddles):
// n = (R2*n1 + R1*n2) mod R,  k = (R2*(R2^-1 mod R1)*k1 + R1*(R1^-1 mod R2)*k2) mod R.
constexpr int mod_inverse(int a, int m) {
    int r = 1;
    while ((a * r) % m != 1) ++r;
    return r;
}
template <int R, int R1, int R2, int DIR, class V> struct DftPFA {
    static LSTED_HD void run(V* v) {
        constexpr int E1 = R2 * mod_inverse(R2 % R1, R1), E2 = R1 * mod_inverse(R1 % R2, R2);
        V y[R2][R1];
        LSTED_UNROLL
        for (int n2 = 0; n2 < R2; ++n2) {
            LSTED_UNROLL
            for (int n1 = 0; n1 < R1; ++n1) y[n2][n1] = v[(R2 * n1 + R1 * n2) % R];
            DftE<R1, DIR, V>::run(y[n2]);
        }
        LSTED_UNROLL
        for (int k1 = 0; k1 < R1; ++k1) {
            V z[R2];
            LSTED_UNROLL
            for (int n2 = 0; n2 < R2; ++n2) z[n2] = y[n2][k1];
            DftE<R2, DIR, V>::run(z);
            LSTED_UNROLL
            for (int k2 = 0; k2 < R2; ++k2) v[(E1 * k1 + E2 * k2) % R] = z[k2];
        }
    }
};
template <int DIR, class V> struct DftE<45, DIR, V> : DftPFA<45, 9, 5, DIR, V> {};
template <int DIR, class V> struct DftE<48, DIR, V> : DftPFA<48, 3, 16, DIR, V> {};

// historical spelling: DFT on cplx<T>
template <int R, int DIR, typename T> struct Dft : DftE<R, DIR, cplx<T> > {};

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
