"""Edge cases around the RL ratio and the Deconvolver state (CPU replay of the
kernel bodies; the GPU twins live in tests/test_gpu_full_size.py).

* expected == 0: with a zero-background object the fp32 forward model is FFT
  round-off in the dark regions, half of it negative -> clipped to 0.  The
  reference divides by that zero (ref:527-528: inf, then NaN in every pixel
  after the next transform).  Decision pinned here: such a pixel contributes a
  ratio of 0, the estimate stays finite.
* Deconvolver state: first iterate() always starts from ones, new data leave
  the estimate alone, replaced PSFs are picked up (ref:496-531, :573, :585).
"""
import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import _lib


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.fixture(scope='module')
def lib():
    return emul_support.emulator_library()


def sparse_object(shape, seed=7, beads=12):
    rng = np.random.default_rng(seed)
    obj = np.zeros((1,) + shape)
    ys = rng.integers(0, shape[0], beads)
    xs = rng.integers(0, shape[1], beads)
    obj[0, ys, xs] = 1.0
    return obj


def gaussian_psfs(K, n=9, sigma=(1.2, 2.5)):
    y, x = np.mgrid[:n, :n] - (n - 1) / 2.0
    out = []
    for k in range(K):
        t = np.pi * k / K
        u, v = x * np.cos(t) + y * np.sin(t), -x * np.sin(t) + y * np.cos(t)
        p = np.exp(-0.5 * (u / sigma[0]) ** 2 - 0.5 * (v / sigma[1]) ** 2)
        out.append(p / p.sum() / K)
    return np.array(out)


@pytest.mark.parametrize('shape,fast', [((96, 96), False), ((6, 2048), True)])
@pytest.mark.parametrize('precision', [32, 64])
def test_dark_background_estimate_stays_finite(lib, shape, fast, precision):
    """Zero-background object, K = 2, 12 iterations: finite, non-zero, no NaN --
    generic kernels (96^2 -> L = 108) and the 2160 fast path (2048 wide)."""
    psfs = gaussian_psfs(2)
    obj = sparse_object(shape)
    h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
    if fast:
        assert h.info().Lx == 2160
    h.create_data(obj, 1e5, 3)
    noisy = np.concatenate([h.get(_lib.NOISY, k) for k in range(2)])
    assert noisy.min() >= 1e-9 * 0.99          # photon counts + 1e-9
    h.iterate(12)
    est = h.get(_lib.ESTIMATE)
    assert np.isfinite(est).all()
    assert est.min() >= 0 and est.sum() > 0
    # the beads are still where the light is: the estimate's mass sits near them
    assert est[obj > 0].sum() > 0
    if precision == 32:
        # fp64 on the same noisy data (the reference's behaviour where it is finite)
        d = _lib.DeconvHandle(lib, psfs, shape, precision=64)
        d.create_data(obj, 1e5, 3, reset_estimate=True)
        for k in range(2):
            d.set(_lib.NOISY, k, noisy[k])
        d.iterate(12)
        ref = d.get(_lib.ESTIMATE)
        # bright pixels agree; the dark background is below fp32 FFT round-off by construction
        bright = ref > 1e-3 * ref.max()
        assert rel_l2(est[bright], ref[bright]) < 1e-2
        d.close()
    h.close()


def test_expected_zero_gives_ratio_zero_not_inf(lib):
    """A measurement with light where the estimate predicts none (estimate forced to 0 on
    half of the image): reference -> inf/NaN (checked on the oracle), here finite."""
    psfs = gaussian_psfs(2)
    shape = (40, 48)
    rng = np.random.default_rng(1)
    meas = [rng.poisson(20.0, (1,) + shape) + 1e-9 for _ in range(2)]
    est0 = np.ones((1,) + shape)
    est0[0, :, :20] = 0.0
    h = _lib.DeconvHandle(lib, psfs, shape, precision=64)
    for k in range(2):
        h.set(_lib.NOISY, k, meas[k])
    h.iterate(1)                   # builds the normalisation, estimate = ones
    h.set(_lib.ESTIMATE, 0, est0)
    h.iterate(1)
    est = h.get(_lib.ESTIMATE)
    assert np.isfinite(est).all()
    assert (est[0, :, :12] == 0).all()
    o = orc.Deconvolver([p[None] for p in psfs])
    o.noisy_measurement = meas
    o.iterate()
    o.estimate = est0.copy()
    with np.errstate(divide='ignore', invalid='ignore'):
        o.iterate()
    assert not np.isfinite(o.estimate).all()     # the reference's answer is inf/NaN
    h.close()


@pytest.mark.parametrize('tile_len', [60, 90, 120])
def test_three_pass_tile_lengths(lib, tile_len):
    """Tile lengths whose crop used to read past the sequence (wrap-around crop index)."""
    rng = np.random.default_rng(0)
    psfs = rng.random((2, 9, 11))
    shape = (130, 101)
    x = rng.random((1,) + shape)
    o = orc.Deconvolver([p[None] for p in psfs])
    for precision, tol in ((64, 1e-12), (32, 1e-5)):
        t = _lib.DeconvHandle(lib, psfs, shape, precision=precision, tile_fft_len=tile_len)
        assert rel_l2(t.H(x), np.concatenate(o.H(x))) < tol
        y = rng.random((2,) + shape)
        assert rel_l2(t.Ht(y, False), o.H_t([v[None] for v in y], normalize=False)) < tol
        t.close()


@pytest.fixture()
def st(monkeypatch):
    from rescan_line_sted_b200 import line_sted_tools
    monkeypatch.setattr(_lib, '_library', emul_support.emulator_library())
    monkeypatch.setenv('LSTED_PRECISION', 'fp64')
    return line_sted_tools


def test_deconvolver_state_follows_the_reference(st, tmp_path):
    rng = np.random.default_rng(2)
    psfs = [p[None] for p in gaussian_psfs(2)]
    obj = rng.random((1, 24, 30)) + 0.1
    d = st.Deconvolver(psfs, output_prefix=str(tmp_path) + '/', verbose=False)
    o = orc.Deconvolver(psfs)
    d.create_data_from_object(obj, total_brightness=1e6, random_seed=0)
    o.create_data_from_object(obj, total_brightness=1e6, random_seed=0)
    d.noisy_measurement = o.noisy_measurement
    # (1) an estimate assigned before the first iterate() is overwritten by ones (ref:521-522)
    d.estimate = 7 * np.ones_like(obj)
    o.estimate = 7 * np.ones_like(obj)
    d.iterate(), o.iterate()
    assert rel_l2(d.estimate, o.estimate) < 1e-12
    # (2) new data keep the estimate and the iteration count (ref:496-512)
    obj2 = rng.random((1, 24, 30)) + 0.1
    d.create_data_from_object(obj2, total_brightness=1e6, random_seed=1)
    o.create_data_from_object(obj2, total_brightness=1e6, random_seed=1)
    d.noisy_measurement = o.noisy_measurement
    assert d.num_iterations == o.num_iterations == 1
    assert rel_l2(d.estimate, o.estimate) < 1e-12
    d.iterate(), o.iterate()
    assert rel_l2(d.estimate, o.estimate) < 1e-12
    # (3) replaced PSFs are used by the next call (ref:573, :585); state carries over
    new = [p[None] for p in gaussian_psfs(2, sigma=(2.0, 2.0))]
    d.psfs = list(new)
    o.psfs = list(new)
    assert rel_l2(np.concatenate(d.H(obj)), np.concatenate(o.H(obj))) < 1e-12
    assert rel_l2(d.estimate, o.estimate) < 1e-12
    assert rel_l2(np.concatenate(d.noisy_measurement), np.concatenate(o.noisy_measurement)) < 1e-15
    d.iterate(), o.iterate()       # (the cached H_t_normalization is kept, stale, like ref:590-592)
    assert rel_l2(d.estimate, o.estimate) < 1e-12
