"""The in-kernel Poisson sampler (csrc/poisson.cuh, CPU replay) against a numpy restatement of
the textbook algorithm on the SAME Philox4x32-10 counters: Hoermann's PTRS exactly as numpy's
legacy generator spells it (squeeze with v_r, exact test with four logarithms and lgamma; the
reference draws with np.random.poisson, ref:510).  The kernel evaluates the two tests in
cheaper forms (no division for v_r; Stirling's series, two logarithms); the variates have to
be the same ones."""
import ctypes

import numpy as np
import pytest
from scipy.special import gammaln

import emul_support

dp = ctypes.POINTER(ctypes.c_double)
M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M32, p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


def u01(hi, lo):
    return (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def textbook_ptrs(lam, seed, pixels):
    """numpy's legacy `random_poisson_ptrs`, attempt `att` drawn from counter
    (pixel lo, pixel hi, image 0, att) with key = seed."""
    slam, loglam = np.sqrt(lam), np.log(lam)
    b = 0.931 + 2.53 * slam
    a = -0.059 + 0.02483 * b
    invalpha = 1.1239 + 1.1328 / (b - 3.4)
    vr = 0.9277 - 3.6224 / (b - 2.0)
    out = np.full(pixels.shape, -1.0)
    todo = np.arange(pixels.size)
    for att in range(64):
        if todo.size == 0:
            break
        p = pixels[todo].astype(np.uint64)
        zero = np.zeros_like(p)
        v = philox4x32_10(p & M32, p >> np.uint64(32), zero, zero + np.uint64(att),
                          seed & 0xFFFFFFFF, seed >> 32)
        U = u01(v[0], v[1]) - 0.5
        V = u01(v[2], v[3])
        us = 0.5 - np.abs(U)
        with np.errstate(divide='ignore', invalid='ignore'):
            k = np.floor((2.0 * a / us + b) * U + lam + 0.43)
            squeeze = (us >= 0.07) & (V <= vr)
            skip = (k < 0.0) | ((us < 0.013) & (V > us))
            exact = (np.log(V) + np.log(invalpha) - np.log(a / (us * us) + b)
                     <= -lam + k * loglam - gammaln(k + 1.0))
        accept = squeeze | (~skip & exact)
        out[todo[accept]] = k[accept]
        todo = todo[~accept]
    assert todo.size == 0
    return out


@pytest.mark.parametrize('lam', [10.0, 12.5, 17.0, 30.0, 100.0, 1e4, 1.8e7])
def test_sampler_is_textbook_ptrs_on_the_same_counters(lam):
    lib = emul_support.emulator_library()
    lib.cdll.emul_poisson.argtypes = [ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_int, dp]
    n, seed, first = 200000, 0x1234567ABCDEF, 3_000_000_000     # pixel index past 2^31
    got = np.empty(n)
    lib.cdll.emul_poisson(lam, seed, first, n, got.ctypes.data_as(dp))
    want = textbook_ptrs(lam, seed, first + np.arange(n, dtype=np.int64))
    # the two spellings of the tests can only disagree when a uniform falls within rounding of
    # a decision boundary (~1e-13 per sample)
    assert np.array_equal(got, want)


def textbook_multiplication(lam, seed, pixels):
    """Knuth's multiplication method (numpy's legacy `random_poisson_mult`): count uniforms
    until their product drops to exp(-lam); each Philox call supplies two of them."""
    enlam = np.exp(-lam)
    out = np.zeros(pixels.size)
    prod = np.ones(pixels.size)
    running = np.ones(pixels.size, dtype=bool)
    for att in range(200):
        idx = np.flatnonzero(running)
        if idx.size == 0:
            break
        p = pixels[idx].astype(np.uint64)
        zero = np.zeros_like(p)
        v = philox4x32_10(p & M32, p >> np.uint64(32), zero, zero + np.uint64(att),
                          seed & 0xFFFFFFFF, seed >> 32)
        ua, ub = u01(v[0], v[1]), u01(v[2], v[3])
        prod[idx] *= ua
        done = ~(prod[idx] > enlam)
        running[idx[done]] = False
        idx, ub = idx[~done], ub[~done]     # the second uniform: pixels still running only
        out[idx] += 1.0
        prod[idx] *= ub
        done = ~(prod[idx] > enlam)
        running[idx[done]] = False
        out[idx[~done]] += 1.0
    assert not running.any()
    return out


@pytest.mark.parametrize('lam', [0.5, 5.0, 9.9])
def test_small_lambda_sampler_is_the_multiplication_method(lam):
    lib = emul_support.emulator_library()
    lib.cdll.emul_poisson.argtypes = [ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_int, dp]
    n, seed, first = 50000, 987654321, 12345
    got = np.empty(n)
    lib.cdll.emul_poisson(lam, seed, first, n, got.ctypes.data_as(dp))
    want = textbook_multiplication(lam, seed, first + np.arange(n, dtype=np.int64))
    assert np.array_equal(got, want)
