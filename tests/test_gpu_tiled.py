"""Tiled engine on the GPU with the 2160 fast-path tiles: against the oracle at a
two-tile size, and at config-5 scale (8192^2 tiles, fp64) through exact local
checks: a linear convolution is local, so any patch of the tiled result must
equal the oracle's convolution of the patch plus its halo."""
import numpy as np
import pytest

from oracle import line_sted_oracle as orc

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.mark.parametrize('precision,tol', [(64, 1e-12), (32, 1e-5)])
def test_two_tiles_against_oracle(precision, tol):
    from rescan_line_sted_b200 import _lib
    rng = np.random.default_rng(1)
    psfs = rng.random((2, 9, 11))
    shape = (2200, 2300)                      # does not fit one 2160 window -> 2 x 2 tiles
    x = rng.random((1,) + shape)
    h = _lib.DeconvHandle(_lib.get(), psfs, shape, precision=precision, tile_fft_len=2160)
    info = h.info()
    assert (info.tiles_y, info.tiles_x) == (2, 2) and info.Lx == 2160
    o = orc.Deconvolver([p[None] for p in psfs])
    assert rel_l2(h.H(x), np.concatenate(o.H(x))) < tol
    h.create_data(x, 1e9, 3)
    o.create_data_from_object(x, 1e9, 3)
    o.noisy_measurement = [h.get(_lib.NOISY, k) for k in range(2)]
    h.iterate(2)
    o.iterate(), o.iterate()
    assert rel_l2(h.get(_lib.ESTIMATE), o.estimate) < 10 * tol
    h.close()


def test_config5_scale_local_exactness():
    """8192^2 object, 107^2 PSFs, fp64, automatic tiling (4 x 4 tiles of 2054^2)."""
    from rescan_line_sted_b200 import _lib
    rng = np.random.default_rng(2)
    N, K, n = 8192, 4, 107
    g = np.exp(-0.5 * (np.arange(n) - n // 2) ** 2 / 6.0 ** 2)
    psfs = np.stack([np.outer(np.roll(g, s), g) + 0.01 * rng.random((n, n)) for s in range(K)])
    psfs /= psfs.sum(axis=(1, 2), keepdims=True)
    x = rng.random((1, N, N))
    h = _lib.DeconvHandle(_lib.get(), psfs, (N, N), precision=64)     # tiles automatically
    info = h.info()
    assert (info.tiles_y, info.tiles_x) == (4, 4) and info.tile_out_y == 2054
    h.create_data(x, None, 11)
    o = orc.Deconvolver([p[None] for p in psfs])
    half = n // 2
    for (py, px) in ((0, 0), (2000, 2030), (4090, 6100), (N - 160, N - 160), (2054 - 30, 0)):
        size = 160
        y0, x0 = max(py - half, 0), max(px - half, 0)
        y1, x1 = min(py + size + half, N), min(px + size + half, N)
        sub = [v[:, py - y0:py - y0 + size, px - x0:px - x0 + size] for v in o.H(x[:, y0:y1, x0:x1])]
        for k in range(K):
            got = h.get(_lib.NOISELESS, k)[:, py:py + size, px:px + size]
            assert rel_l2(got, sub[k]) < 1e-12, (py, px, k)
    # Poisson field: integer counts, right mean
    noisy = h.get(_lib.NOISY, 0) - 1e-9
    nl = h.get(_lib.NOISELESS, 0)
    assert np.abs(noisy - np.round(noisy)).max() < 1e-6
    assert abs((noisy - nl).mean()) < 5 * np.sqrt(nl.mean() / nl.size)
    h.iterate(1)
    est = h.get(_lib.ESTIMATE)
    assert np.isfinite(est).all() and est.min() >= 0
    h.close()
