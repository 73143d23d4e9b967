"""The compile-time-planned kernels (L = 2160) against the generic kernels
and the oracle at a size the oracle finishes in seconds."""
import numpy as np
import pytest

from oracle import line_sted_oracle as orc

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


@pytest.mark.parametrize('precision,tol', [(64, 1e-12), (32, 1e-5)])
@pytest.mark.parametrize('shape', [(2100, 2100), (2048, 2101), (2089, 2048)])
def test_fast_path_matches_oracle_and_generic(precision, tol, shape):
    from rescan_line_sted_b200 import _lib
    lib = _lib.get()
    rng = np.random.default_rng(0)
    psfs = rng.random((2, 9, 11))
    Ny, Nx = shape
    x = rng.random((1, Ny, Nx))
    y = rng.random((2, Ny, Nx))
    o = orc.Deconvolver([p[None] for p in psfs])
    Hx = np.concatenate(o.H(x))
    Hty = o.H_t([v[None] for v in y], normalize=False)
    h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
    info = h.info()
    assert (info.Ly, info.Lx) == (2160, 2160)
    assert rel_l2(h.H(x), Hx) < tol
    assert rel_l2(h.Ht(y, False), Hty) < tol
    h.create_data(x, 1e9, 0)
    noisy = [h.get(_lib.NOISY, k) for k in range(2)]
    h.iterate(2)
    o.create_data_from_object(x, 1e9, 0)
    o.noisy_measurement = noisy
    o.iterate(), o.iterate()
    assert rel_l2(h.get(_lib.ESTIMATE), o.estimate) < 10 * tol
    if precision == 32:   # the generic kernels fit in shared memory for fp32 only
        g = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
        g.set_option('fast_path', 0)
        assert rel_l2(g.H(x), Hx) < tol
        g.create_data(x, 1e9, 0)
        for k in range(2):
            g.set(_lib.NOISY, k, noisy[k])
        g.iterate(2)
        assert rel_l2(g.get(_lib.ESTIMATE), h.get(_lib.ESTIMATE)) < 10 * tol
        g.close()
    h.close()
