"""Overlap-save tiling (objects larger than one shared-memory transform) on the
CPU replay of the kernel bodies: the tiled operators must equal the untiled
ones and the oracle, including the Poisson field."""
import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import _lib


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.fixture(scope='module')
def lib():
    return emul_support.emulator_library()


@pytest.mark.parametrize('shape', [(150, 170), (64, 64), (47, 201), (57, 56)])
@pytest.mark.parametrize('precision,tol', [(64, 1e-12), (32, 1e-5)])
def test_tiled_equals_untiled_and_oracle(lib, shape, precision, tol):
    rng = np.random.default_rng(0)
    psfs = rng.random((3, 9, 11))
    x = rng.random((1,) + shape)
    y = rng.random((3,) + shape)
    o = orc.Deconvolver([p[None] for p in psfs])
    tiled = _lib.DeconvHandle(lib, psfs, shape, precision=precision, tile_fft_len=64)
    info = tiled.info()
    assert (info.Ly, info.Lx) == (64, 64)
    assert info.tile_out_y == 64 - 8 and info.tile_out_x == 64 - 10
    assert info.tiles_y == -(-shape[0] // 56) and info.tiles_x == -(-shape[1] // 54)
    assert rel_l2(tiled.H(x), np.concatenate(o.H(x))) < tol
    ylist = [v[None] for v in y]
    assert rel_l2(tiled.Ht(y, False), o.H_t(ylist, normalize=False)) < tol
    assert rel_l2(tiled.Ht(y, True), o.H_t(ylist)) < tol
    plain = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
    assert plain.info().tiles_y == 1
    tiled.create_data(x, 1e7, 5)
    plain.create_data(x, 1e7, 5)
    if precision == 64:  # same Philox counters (global pixel index), same lambda
        for k in range(3):
            assert np.array_equal(tiled.get(_lib.NOISY, k), plain.get(_lib.NOISY, k))
    for k in range(3):
        tiled.set(_lib.NOISY, k, plain.get(_lib.NOISY, k))
    tiled.iterate(3)
    plain.iterate(3)
    assert rel_l2(tiled.get(_lib.ESTIMATE), plain.get(_lib.ESTIMATE)) < 10 * tol
    assert tiled.info().iterations_done == 3
    # record_iteration's error spectrum of a tiled object (direct transform, ref:539-546)
    est, true = tiled.get(_lib.ESTIMATE), tiled.get(_lib.TRUE_OBJECT)
    want = np.log(1 + np.abs(np.fft.fftshift(np.fft.fftn(est - true, axes=(1, 2)), axes=(1, 2))))
    ft_tol = 1e-12 if precision == 64 else 2e-4
    assert np.abs(tiled.ft_error() - want).max() <= ft_tol * np.abs(want).max()
    assert np.abs(tiled.ft_error(est) - want).max() <= ft_tol * np.abs(want).max()
    tiled.close(), plain.close()


def test_tile_shorter_than_psf_is_an_error(lib):
    with pytest.raises(RuntimeError, match='(?i)shorter|too short'):
        _lib.DeconvHandle(lib, np.ones((1, 9, 9)), (40, 40), precision=64, tile_fft_len=8)
