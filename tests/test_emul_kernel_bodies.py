"""CPU replay (tests/host_emul) of the very kernel bodies the sm_100a build
launches: FFT passes, row/column convolution passes, fused RL iteration,
PSF synthesis, Poisson sampler -- checked against numpy, the oracle and the
reference's golden vectors.  Test infrastructure only; see emul.cpp."""
import ctypes
import json
import os

import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import _lib

dp = ctypes.POINTER(ctypes.c_double)


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.fixture(scope='module')
def lib():
    return emul_support.emulator_library()


def emul_fft(lib, x, direction, precision):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    assert lib.cdll.emul_fft(x.shape[1], direction, x.shape[0], precision,
                             x.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
    return out


@pytest.mark.parametrize('L', [8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 36,
                               45, 48, 60, 64, 75, 80, 81, 96, 100, 120, 125, 128,
                               135, 144, 150, 160, 180, 192, 200, 216, 225, 240,
                               243, 250, 256, 270, 288, 300, 320, 360, 375, 384,
                               400, 405, 432, 450, 480, 486, 500, 512, 540, 576,
                               600, 625, 640, 648, 675, 720, 729, 750, 768, 800,
                               810, 864, 900, 960, 972, 1000, 1024, 1080, 2160])
def test_stockham_fft_matches_numpy(lib, L):
    rng = np.random.default_rng(L)
    x = rng.standard_normal((2, L)) + 1j * rng.standard_normal((2, L))
    ref = np.fft.fft(x, axis=1)
    assert np.abs(emul_fft(lib, x, -1, 64) - ref).max() < 1e-13 * np.abs(ref).max() * np.log2(L)
    assert np.abs(emul_fft(lib, x, -1, 32) - ref).max() < 3e-7 * np.abs(ref).max() * np.log2(L)
    inv = np.fft.ifft(x, axis=1) * L
    assert np.abs(emul_fft(lib, x, +1, 64) - inv).max() < 1e-13 * np.abs(inv).max() * np.log2(L)


def test_fft_length_planner(lib):
    for n, want in ((1, 8), (181, 192), (213, 216), (2101, 2160), (8245, 8640),
                    (1025, 1080), (4097, 4320)):
        assert lib.cdll.emul_next_smooth_len(n) == want
    rad = (ctypes.c_int * 12)()
    for L in (2160, 8640, 192, 8):
        n = lib.cdll.emul_fft_plan(L, rad, 12)
        assert n > 0 and int(np.prod(rad[:n])) == L
    assert lib.cdll.emul_fft_plan(14, rad, 12) == -1     # 7 is not a supported radix


@pytest.mark.parametrize('precision,tol', [(64, 1e-12), (32, 1e-5)])
def test_engine_matches_reference_golden(lib, golden_dir, precision, tol):
    g = np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))
    h = _lib.DeconvHandle(lib, g['psfs'], (128, 128), precision=precision)
    h.create_data(g['object_u8'].astype(np.float64), 5e10, 0)
    nl = np.concatenate([h.get(_lib.NOISELESS, k) for k in range(4)])
    assert rel_l2(nl, g['noiseless']) < tol
    for k in range(4):
        h.set(_lib.NOISY, k, g['noisy'][k])
    h.iterate(1)
    assert rel_l2(h.get(_lib.ESTIMATE), g['estimate_1']) < tol
    assert rel_l2(h.get(_lib.NORMALIZATION), g['H_t_normalization']) < tol
    h.iterate(7)
    assert rel_l2(h.get(_lib.ESTIMATE), g['estimate_8']) < 10 * tol
    assert h.info().iterations_done == 8
    # clip-per-term variant (reference order of operations)
    h.set_option('exact_clip', 1)
    h.create_data(g['object_u8'].astype(np.float64), 5e10, 0)
    for k in range(4):
        h.set(_lib.NOISY, k, g['noisy'][k])
    h.iterate(8)
    assert rel_l2(h.get(_lib.ESTIMATE), g['estimate_8']) < 10 * tol
    h.close()


@pytest.mark.parametrize('shape,pshape', [((33, 40), (9, 9)), ((7, 300), (11, 5)),
                                          ((50, 31), (8, 6)), ((1, 64), (1, 9)),
                                          ((5, 9), (11, 11)), ((2, 2), (1, 1))])
def test_H_Ht_ragged_shapes(lib, shape, pshape):
    rng = np.random.default_rng(3)
    psfs = rng.random((3,) + pshape)
    h = _lib.DeconvHandle(lib, psfs, shape, precision=64)
    o = orc.Deconvolver([p[None] for p in psfs])
    x = rng.random((1,) + shape)
    assert rel_l2(h.H(x), np.concatenate(o.H(x))) < 1e-12
    y = rng.random((3,) + shape)
    ylist = [v[None] for v in y]
    assert rel_l2(h.Ht(y, False), o.H_t(ylist, normalize=False)) < 1e-12
    assert rel_l2(h.Ht(y, True), o.H_t(ylist)) < 1e-12
    h.close()


def test_error_paths(lib):
    with pytest.raises(RuntimeError, match='precision'):
        _lib.DeconvHandle(lib, np.ones((1, 3, 3)), (8, 8), precision=16)
    h = _lib.DeconvHandle(lib, np.ones((1, 3, 3)), (8, 8), precision=64)
    with pytest.raises(RuntimeError, match='selector'):
        h.get(_lib.NOISY, 5)
    with pytest.raises(RuntimeError, match='unknown option'):
        h.set_option('nope', 1)
    h.close()


def poisson(lib, lam, n, seed=1, first=0):
    out = np.empty(n)
    lib.cdll.emul_poisson.argtypes = [ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_int, dp]
    lib.cdll.emul_poisson(lam, seed, first, n, out.ctypes.data_as(dp))
    return out


@pytest.mark.parametrize('lam', [0.1, 1.0, 5.0, 9.9, 10.0, 30.0])
def test_poisson_small_lambda_pmf(lib, lam):
    """Chi-square of the sampled histogram against the exact Poisson pmf."""
    from scipy.stats import poisson as sp_poisson, chi2
    n = 200000
    x = poisson(lib, lam, n)
    assert np.all(x == np.round(x)) and x.min() >= 0
    kmax = int(sp_poisson.ppf(1 - 1e-5, lam)) + 1
    obs = np.bincount(x.astype(int), minlength=kmax + 1)
    obs = np.append(obs[:kmax], obs[kmax:].sum())
    exp = np.append(sp_poisson.pmf(np.arange(kmax), lam), sp_poisson.sf(kmax - 1, lam)) * n
    keep = exp > 5
    stat = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stat < chi2.ppf(1 - 1e-6, keep.sum() - 1)


@pytest.mark.parametrize('lam', [100.0, 1e4, 1.8e7])
def test_poisson_large_lambda_moments(lib, lam):
    n = 100000
    x = poisson(lib, lam, n, seed=3)
    assert abs(x.mean() - lam) < 5 * np.sqrt(lam / n)
    assert abs(x.var() / lam - 1) < 5 * np.sqrt(2.0 / n)
    z = (x - lam) / np.sqrt(lam)
    assert abs((z ** 3).mean() - 1 / np.sqrt(lam)) < 0.05      # skewness
    assert poisson(lib, 0.0, 10).max() == 0
    a, b = poisson(lib, lam, 100, seed=3), poisson(lib, lam, 100, seed=4)
    assert np.array_equal(a, x[:100]) and not np.array_equal(a, b)


def test_fast_path_noise_field_is_the_sampler_field(lib):
    """ROW_INV_SIM of the 2160 fast path draws the noise in two steps (first PTRS attempt,
    then a shared-memory queue of the pixels the squeeze rejected); the field has to be
    exactly poisson_sample(noiseless[pixel], seed, pixel, image) -- also for lambda < 10
    and lambda == 0 -- and identical to the generic kernels' field."""
    rng = np.random.default_rng(5)
    ny, nx, seed = 4, 2048, 11
    psfs = rng.random((2, 3, 9))
    obj = rng.random((1, ny, nx)) * 40.0
    obj[0, 1, :700] *= 1e4              # large lambda (PTRS) ...
    obj[0, :, 100:400] = 0.0            # ... dark stretches (lambda ~ 0) ...
    obj[0, :, 1200:1500] *= 0.01        # ... and lambda < 10 (multiplication method)
    fields = []
    for fast in (1, 0):
        h = _lib.DeconvHandle(lib, psfs, (ny, nx), precision=64)
        h.set_option('fast_path', fast)
        h.create_data(obj, None, seed)
        assert (h.info().Lx == 2160) and (h.info().row_pairs_per_cta >= 1)
        noiseless = np.stack([h.get(_lib.NOISELESS, k) for k in range(2)])
        noisy = np.stack([h.get(_lib.NOISY, k) for k in range(2)])
        fields.append(noisy)
        h.close()
    assert np.array_equal(fields[0], fields[1])
    lib.cdll.emul_poisson.argtypes = [ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_int, dp]
    one = np.empty(1)
    flat = noiseless.reshape(2, -1)
    for k in range(2):
        for pix in list(range(0, ny * nx, 97)) + [ny * nx - 1]:
            # emul_poisson samples image 0; image k is covered by the fast/generic equality
            if k == 0:
                lib.cdll.emul_poisson(float(flat[k, pix]), seed, pix, 1, one.ctypes.data_as(dp))
                assert fields[0].reshape(2, -1)[k, pix] == one[0] + 1e-9
    assert (noiseless < 10).any() and (noiseless > 1e4).any()


def test_dual_row_mid_matches_single_pair_kernel_and_oracle(lib):
    """row_mid_dual_body (two row pairs per thread group, c2 elements) against the
    single-pair fast kernel and the oracle on a 2160-wide geometry, fp32."""
    rng = np.random.default_rng(8)
    ny, nx = 8, 2048
    psfs = rng.random((3, 5, 107)) + 0.1
    psfs /= psfs.sum(axis=(1, 2), keepdims=True)
    obj = rng.random((1, ny, nx)) + 0.05
    meas = [rng.poisson(50.0, (1, ny, nx)).astype(np.float64) + 1e-9 for _ in range(3)]
    est = {}
    launches = {}
    for dual in (1, 0):
        before = lib.cdll.emul_dual_launches()
        h = _lib.DeconvHandle(lib, psfs, (ny, nx), precision=32)
        h.set_option('row_plan2', 0)        # the three-pass row kernels
        h.set_option('row_dual', dual)
        h.create_data(obj, 1e7, 1)
        assert h.info().Lx == 2160
        for k in range(3):
            h.set(_lib.NOISY, k, meas[k])
        h.iterate(2)
        est[dual] = h.get(_lib.ESTIMATE)
        launches[dual] = lib.cdll.emul_dual_launches() - before
        h.close()
    assert launches == {1: 2, 0: 0}
    o = orc.Deconvolver([p[None] for p in psfs])
    o.create_data_from_object(obj, total_brightness=1e7, random_seed=1)
    o.noisy_measurement = meas
    o.iterate(); o.iterate()
    assert rel_l2(est[1], est[0]) < 1e-6
    assert rel_l2(est[1], o.estimate) < 1e-5
    assert rel_l2(est[0], o.estimate) < 1e-5


@pytest.mark.parametrize('shape', [(7, 2100), (6, 2048), (2, 2101)])
def test_fast_rows_odd_and_ragged_shapes(lib, shape):
    """2160-wide fast row kernels on images with an odd / non-multiple-of-4 number of
    rows (XB2 spectrum layout: the missing last row has a never-read slot)."""
    rng = np.random.default_rng(12)
    psfs = rng.random((2, 3, 21))
    x = rng.random((1,) + shape)
    y = rng.random((2,) + shape)
    for precision, tol in ((64, 1e-12), (32, 1e-5)):
        o = orc.Deconvolver([p[None] for p in psfs])
        h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
        assert h.info().Lx == 2160
        assert rel_l2(h.H(x), np.concatenate(o.H(x))) < tol
        assert rel_l2(h.Ht(y, True), o.H_t([v[None] for v in y])) < tol
        h.create_data(x, 1e8, 3)
        o.create_data_from_object(x, 1e8, 3)
        noisy = [h.get(_lib.NOISY, k) for k in range(2)]
        o.noisy_measurement = noisy
        h.iterate(2)
        o.iterate(); o.iterate()
        assert rel_l2(h.get(_lib.ESTIMATE), o.estimate) < 10 * tol
        h.close()


def test_two_pass_plan_48x45_matches_numpy(lib):
    """Fft2E (one exchange, one twiddle stage; Good-Thomas 3x16 and 9x5 butterflies)."""
    rng = np.random.default_rng(2160)
    x = rng.standard_normal(2160) + 1j * rng.standard_normal(2160)
    xin = np.ascontiguousarray(x)
    out = np.empty_like(xin)
    for d, ref in ((-1, np.fft.fft(x)), (+1, np.fft.ifft(x) * 2160)):
        for precision, tol in ((64, 1e-13), (32, 3e-6)):
            assert lib.cdll.emul_fft2_2160(d, precision, xin.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
            assert np.abs(out - ref).max() < tol * np.abs(ref).max(), (d, precision)


@pytest.mark.parametrize('shape', [(8, 2048), (5, 2100), (2, 2101)])
def test_two_pass_row_kernels_match_three_pass_and_oracle(lib, shape):
    """row2_fast_body (48 x 45 plan, 48 threads per row pair, all five row modes) against
    the three-pass row kernels and the oracle: forward model + noise field + 2 RL
    iterations + H / H_t, fp32."""
    rng = np.random.default_rng(21)
    ny, nx = shape
    psfs = rng.random((3, 3, 107)) + 0.1
    psfs /= psfs.sum(axis=(1, 2), keepdims=True)
    obj = rng.random((1, ny, nx)) + 0.05
    y = rng.random((3, ny, nx)) + 0.1
    res = {}
    for plan2 in (1, 0):
        before = lib.cdll.emul_row2_launches()
        h = _lib.DeconvHandle(lib, psfs, shape, precision=32)
        h.set_option('row_plan2', plan2)
        assert h.info().Lx == 2160
        h.create_data(obj, 1e5 * obj.size, 4)
        noisy = np.stack([h.get(_lib.NOISY, k) for k in range(3)])
        noiseless = np.stack([h.get(_lib.NOISELESS, k) for k in range(3)])
        h.iterate(2)
        res[plan2] = dict(noisy=noisy, noiseless=noiseless, est=h.get(_lib.ESTIMATE),
                          H=h.H(obj), Ht=h.Ht(y, True),
                          launches=lib.cdll.emul_row2_launches() - before)
        h.close()
    # (the PSF rows are transformed at handle creation, before the option is read: 1 launch)
    assert res[1]['launches'] > 8 and res[0]['launches'] <= 1
    o = orc.Deconvolver([p[None] for p in psfs])
    o.create_data_from_object(obj, total_brightness=1e5 * obj.size, random_seed=4)
    assert rel_l2(res[1]['noiseless'], np.concatenate(o.noiseless_measurement)[:, None]) < 1e-5
    assert rel_l2(res[1]['noiseless'], res[0]['noiseless']) < 2e-6
    # same noiseless image up to rounding -> same Poisson field except where lambda differs in
    # the last bits; compare the iterated estimates on the two-pass kernel's own field
    o.noisy_measurement = [v for v in res[1]['noisy']]
    o.iterate(); o.iterate()
    assert rel_l2(res[1]['est'], o.estimate) < 1e-4
    assert rel_l2(res[1]['H'], np.concatenate(o.H(obj))) < 1e-5
    assert rel_l2(res[1]['Ht'], o.H_t([v[None] for v in y])) < 1e-5
    assert rel_l2(res[1]['H'], res[0]['H']) < 2e-6


@pytest.mark.parametrize('precision,tol', [(64, 1e-12), (32, 1e-5)])
def test_centred_real_otfs_for_point_symmetric_psfs(lib, monkeypatch, precision, tol):
    """Point-symmetric PSFs (all the reference's are) on a 2160-long column geometry: the
    OTFs are stored centred and real, the crop offsets drop out of the geometry; against
    the oracle and against the same handle with the option off (complex OTFs)."""
    rng = np.random.default_rng(31)
    shape = (2100, 8)
    half = rng.random((3, 11, 3))
    psfs = np.concatenate([half, rng.random((3, 1, 3)), half[:, ::-1, ::-1]], axis=1)   # 23 x 3
    psfs[:, 11, :] = 0.5 * (psfs[:, 11, :] + psfs[:, 11, ::-1])
    assert np.abs(psfs - psfs[:, ::-1, ::-1]).max() == 0
    obj = rng.random((1,) + shape) + 0.05
    y = rng.random((3,) + shape) + 0.1
    o = orc.Deconvolver([p[None] for p in psfs])
    res = {}
    launches = {}
    for real in (1, 0):
        monkeypatch.setenv('LSTED_REAL_OTF', str(real))   # read when the handle is created
        before = lib.cdll.emul_real_otf_launches()
        h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
        assert h.info().Ly == 2160
        res[real] = dict(H=h.H(obj), Ht=h.Ht(y, True))
        h.create_data(obj, 1e5 * obj.size, 9)
        noisy = [h.get(_lib.NOISY, k) for k in range(3)]
        h.iterate(2)
        res[real]['est'] = h.get(_lib.ESTIMATE)
        res[real]['noisy'] = noisy
        launches[real] = lib.cdll.emul_real_otf_launches() - before
        h.close()
    assert launches[1] >= 6 and launches[0] == 0
    assert rel_l2(res[1]['H'], np.concatenate(o.H(obj))) < tol
    assert rel_l2(res[1]['Ht'], o.H_t([v[None] for v in y])) < tol
    o.create_data_from_object(obj, total_brightness=1e5 * obj.size, random_seed=9)
    o.noisy_measurement = res[1]['noisy']
    o.iterate(); o.iterate()
    assert rel_l2(res[1]['est'], o.estimate) < 10 * tol
    assert rel_l2(res[1]['H'], res[0]['H']) < tol


def test_sub_block_column_ctas_are_bit_identical(lib):
    """fp32 column kernels on real OTFs run as sub-block CTAs (2 of the 4 columns of a block,
    two CTAs per SM on the GPU; option `col_sub`): same butterflies on the same numbers, so
    the results agree bit for bit with the whole-block CTAs, and with the oracle."""
    rng = np.random.default_rng(5)
    shape = (2090, 10)
    half = rng.random((2, 4, 5))
    psfs = np.concatenate([half, rng.random((2, 1, 5)), half[:, ::-1, ::-1]], axis=1)   # 9 x 5
    psfs[:, 4, :] = 0.5 * (psfs[:, 4, :] + psfs[:, 4, ::-1])
    obj = rng.random((1,) + shape) + 0.05
    y = rng.random((2,) + shape) + 0.1
    res, launches = {}, {}
    for sub in (1, 0):
        before = lib.cdll.emul_col_sub_launches()
        h = _lib.DeconvHandle(lib, psfs, shape, precision=32)
        h.set_option('col_sub', sub)
        assert h.info().Ly == 2160
        res[sub] = dict(H=h.H(obj), Ht=h.Ht(y, True))
        h.create_data(obj, 1e5 * obj.size, 9)
        h.iterate(2)
        res[sub]['est'] = h.get(_lib.ESTIMATE)
        launches[sub] = lib.cdll.emul_col_sub_launches() - before
        h.close()
    assert launches[1] >= 6 and launches[0] == 0
    for key in ('H', 'Ht', 'est'):
        assert np.array_equal(res[1][key], res[0][key]), key
    o = orc.Deconvolver([p[None] for p in psfs])
    assert rel_l2(res[1]['H'], np.concatenate(o.H(obj))) < 1e-5
    assert rel_l2(res[1]['Ht'], o.H_t([v[None] for v in y])) < 1e-5


def test_peer_reduction_block_order(lib):
    """Every rank walks all column blocks exactly once, the ones it does not own first
    (ascending), then its own (xb % world == rank): the order the fused H_t reduction relies
    on to have the peers' partial sums in place when it reaches the blocks it finishes."""
    for world in (2, 3, 4, 8):
        for nxb in (271, 8, 9, 541):
            if nxb < world:
                continue
            for me in range(world):
                seq = [lib.cdll.emul_p2p_block_at(p, me, world, nxb) for p in range(nxb)]
                assert sorted(seq) == list(range(nxb))
                owned = [x for x in range(nxb) if x % world == me]
                others = [x for x in range(nxb) if x % world != me]
                assert seq == others + owned


@pytest.mark.parametrize('shape', [(8, 2048), (5, 2100)])
def test_row_kernels_with_staged_spectrum_chunks(lib, shape):
    """ROW_MID / ROW_FINAL with the pair's spectrum chunks staged through the second exchange
    buffer (tensor-map bulk copies on the GPU, the same box copies as loops on the host):
    same estimates as with per-thread loads / stores, and as the oracle."""
    rng = np.random.default_rng(41)
    psfs = rng.random((3, 3, 107)) + 0.1
    psfs /= psfs.sum(axis=(1, 2), keepdims=True)
    obj = rng.random((1,) + shape) + 0.05
    meas = [rng.poisson(80.0, (1,) + shape).astype(np.float64) + 1e-9 for _ in range(3)]
    est, launches = {}, {}
    for tma in (2, 1, 0):   # 2: also the two-buffer ROW_MID (measurement rows read in place)
        before = lib.cdll.emul_row_tma_launches()
        h = _lib.DeconvHandle(lib, psfs, shape, precision=32)
        h.set_option('row_tma', tma)
        h.create_data(obj, 1e7, 1)
        for k in range(3):
            h.set(_lib.NOISY, k, meas[k])
        h.iterate(3)
        est[tma] = h.get(_lib.ESTIMATE)
        launches[tma] = lib.cdll.emul_row_tma_launches() - before
        h.close()
    assert launches == {2: 6, 1: 6, 0: 0}
    o = orc.Deconvolver([p[None] for p in psfs])
    o.noisy_measurement = meas
    for _ in range(3):
        o.iterate()
    assert np.array_equal(est[1], est[0])      # same arithmetic, only the data path differs
    assert np.array_equal(est[2], est[0])
    assert rel_l2(est[1], o.estimate) < 1e-4


@pytest.mark.parametrize('shape', [(96, 128), (45, 50), (128, 75), (14, 128), (77, 91)])
def test_error_spectrum_on_device(lib, shape):
    """record_iteration's log(1 + |fftshift(fft2(estimate - true_object))|) (ref:539-546)
    from the un-padded device transform: even / odd sides, the estimate in HBM or a host
    image; sides with a prime factor above 5 go through the direct transform (no host path)."""
    rng = np.random.default_rng(3)
    psfs = rng.random((2, 5, 5))
    obj = rng.random((1,) + shape) + 0.1
    for precision, tol in ((64, 1e-12), (32, 2e-4)):   # fp32: bins near zero carry the transform's noise
        h = _lib.DeconvHandle(lib, psfs, shape, precision=precision)
        h.create_data(obj, 1e4 * obj.size, 5)
        h.iterate(2)
        est, true = h.get(_lib.ESTIMATE), h.get(_lib.TRUE_OBJECT)
        got = h.ft_error()
        assert got is not None
        want = np.log(1 + np.abs(np.fft.fftshift(np.fft.fftn(est - true, axes=(1, 2)), axes=(1, 2))))
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= tol * np.abs(want).max()
        other = rng.random((1,) + shape) * true.mean()
        got = h.ft_error(other)
        want = np.log(1 + np.abs(np.fft.fftshift(np.fft.fftn(other - true, axes=(1, 2)), axes=(1, 2))))
        assert np.abs(got - want).max() <= tol * np.abs(want).max()
        h.iterate(1)        # the scratch image it used does not disturb the iteration state
        assert np.isfinite(h.get(_lib.ESTIMATE)).all()
        h.close()


def test_small_frames_spread_orientations_over_ctas():
    """Generic col_h on a small frame: the K orientations of a column block go to several CTAs
    (each repeating the forward transform) -- same numbers as the oracle, and the path is taken."""
    import ctypes
    lib = emul_support.emulator_library()
    lib.cdll.emul_k_split_launches.restype = ctypes.c_int
    before = lib.cdll.emul_k_split_launches()
    rng = np.random.default_rng(21)
    for K in (2, 4, 5):
        psfs = rng.random((K, 9, 7))
        x = rng.random((1, 40, 52)) + 0.1
        h = _lib.DeconvHandle(lib, psfs, (40, 52), precision=64)
        o = orc.Deconvolver([p[None] for p in psfs])
        got, want = h.H(x), np.concatenate(o.H(x))
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-12
        h.close()
    assert lib.cdll.emul_k_split_launches() > before
