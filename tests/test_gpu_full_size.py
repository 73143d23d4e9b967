"""BASELINE config 4 at its full size (2048^2 object, 16 orientations of the figure-2
rescan PSF, fp32) through the C ABI: direct parity with the oracle where the oracle
finishes in seconds, and size-independent properties of the operators for the rest
(impulse response, linearity, adjointness for point-symmetric PSFs, Richardson-Lucy
fixed point, Poisson moments).  Tolerance of BASELINE.json: rel-L2 <= 1e-5 in fp32."""
import os

import numpy as np
import pytest

import bench
from oracle import line_sted_oracle as orc

pytestmark = pytest.mark.gpu
N, K, TOL = 2048, 16, 1e-5


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.fixture(scope='module')
def setup(golden_dir):
    from rescan_line_sted_b200 import _lib, line_sted_tools as st, orientations
    base = np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))['base_psf']
    psfs = orientations.line_orientation_psfs(base, K, bench.EMISSION_2P0X_LR)
    ref = orc.orientation_psfs(base, K, bench.EMISSION_2P0X_LR)
    for p, r in zip(psfs, ref):
        assert rel_l2(p, r) < 1e-12
    h = _lib.DeconvHandle(_lib.get(), st._stack_psfs(psfs), (N, N), precision=32)
    info = h.info()
    assert (info.Ly, info.Lx) == (2160, 2160)
    yield h, psfs, _lib
    h.close()


def test_forward_model_against_oracle_on_two_orientations(setup):
    h, psfs, _lib = setup
    obj = bench.synthetic_object(N)
    h.create_data(obj, bench.total_brightness(N), 0)
    scaled = obj * (bench.total_brightness(N) / obj.sum())
    assert rel_l2(h.get(_lib.TRUE_OBJECT), scaled) < 1e-6
    for k in (0, 7):
        want = np.clip(orc.fftconvolve_same(scaled, psfs[k]), 0, None)
        assert rel_l2(h.get(_lib.NOISELESS, k), want) < TOL, k


def test_impulse_response_is_the_psf(setup):
    h, psfs, _lib = setup
    x = np.zeros((1, N, N))
    spots = [(1000, 1100), (53, 53), (N - 54, 70), (0, 0), (N - 1, N - 1)]
    for (r, c) in spots:
        x[0, r, c] = 1.0
    out = h.H(x)
    n = psfs[0].shape[-1]
    s = (n - 1) // 2
    for k in range(K):
        want = np.zeros((N + 2 * s, N + 2 * s))
        for (r, c) in spots:
            want[r:r + n, c:c + n] += psfs[k][0]
        want = want[s:s + N, s:s + N]
        assert np.abs(out[k] - want).max() <= 2e-6 * psfs[k].max(), k


def test_linearity_and_adjointness(setup):
    h, psfs, _lib = setup
    rng = np.random.default_rng(1)
    x, z = rng.random((1, N, N)), rng.random((1, N, N))
    Hx, Hz = h.H(x), h.H(z)
    assert rel_l2(h.H(2.0 * x + 0.5 * z), 2.0 * Hx + 0.5 * Hz) < TOL
    # H_t convolves with the un-flipped PSFs (ref:586); it is the adjoint of H because the
    # PSFs are point-symmetric about their centre
    for p in psfs:
        assert np.abs(p - p[:, ::-1, ::-1]).max() < 1e-12 * p.max()
    y = rng.random((K, N, N))
    lhs = float(np.sum(Hx.astype(np.float64) * y))
    rhs = float(np.sum(x * h.Ht(y, False)))
    assert abs(lhs - rhs) <= TOL * abs(lhs)
    ones = h.Ht(np.ones((K, N, N)), False)
    assert rel_l2(h.get(_lib.NORMALIZATION), ones) < TOL


def test_richardson_lucy_fixed_point_and_one_oracle_iteration(setup):
    h, psfs, _lib = setup
    rng = np.random.default_rng(2)
    x = rng.random((1, N, N)) + 0.5
    h.create_data(x, None, 0)
    for k in range(K):                       # noise-free measurement: x is a fixed point
        h.set(_lib.NOISY, k, h.get(_lib.NOISELESS, k))
    h.set(_lib.ESTIMATE, 0, x)
    h.iterate(3)
    assert rel_l2(h.get(_lib.ESTIMATE), x) < 10 * TOL
    # one full RL iteration from the flat start against the oracle (2K = 32 convolutions)
    noisy = [h.get(_lib.NOISELESS, k) * (1.0 + 0.1 * rng.random((1, N, N))) for k in range(K)]
    for k in range(K):
        h.set(_lib.NOISY, k, noisy[k])
    h.set(_lib.ESTIMATE, 0, np.ones((1, N, N)))
    h.iterate(1)
    import scipy.fft
    o = orc.Deconvolver(psfs)
    o.noisy_measurement = noisy
    with scipy.fft.set_workers(os.cpu_count()):
        o.iterate()
    assert rel_l2(h.get(_lib.ESTIMATE), o.estimate) < 10 * TOL


def test_poisson_moments_at_full_size(setup):
    h, psfs, _lib = setup
    obj = bench.synthetic_object(N)
    h.create_data(obj, bench.total_brightness(N), 5)
    for k in (0, 9, 15):
        lam = h.get(_lib.NOISELESS, k)[0, 64:-64, 64:-64].astype(np.float64)
        cnt = h.get(_lib.NOISY, k)[0, 64:-64, 64:-64].astype(np.float64)
        z = (cnt - lam) / np.sqrt(np.maximum(lam, 1.0))
        n = z.size
        # fp32 storage rounds counts above 2^24 to multiples of 2; moments are unaffected
        assert abs(z.mean()) < 5 / np.sqrt(n) + 1e-3
        assert abs(z.var() - 1.0) < 5 * np.sqrt(2.0 / n) + 2e-3
