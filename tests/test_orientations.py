"""Orientation step (line_sted_figure_2.py:264-272): the oracle's restatement of
scipy.ndimage.rotate against scipy itself and against the committed golden PSF
set, and the kernel body (CPU replay / GPU) against both.  fp64 throughout:
rel-L2 <= 1e-12, elementwise <= 1e-12 of the peak."""
import os

import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import _lib

ANGLES = list(np.arange(0, 180, 180 / 16))


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.fixture(scope='module')
def base(golden_dir):
    return np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))['base_psf']


def test_restatement_matches_scipy(base):
    unit = base / base.sum()
    for a in ANGLES + [7.0, 123.4]:
        ref = orc.rotate_psf(unit, a)               # scipy, as the reference calls it
        got = orc.rotate_psf_restated(unit, a)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), a


def test_restatement_reproduces_golden_psfs(base, golden_dir):
    """tests/golden/fig2_2p0x_lr.npz['psfs'] was produced by the unmodified reference
    pipeline (make_golden.py): 4 orientations of the base PSF."""
    gold = np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))['psfs']
    K = gold.shape[0]
    unit = base / base.sum()
    mine = [orc.rotate_psf_restated(unit, a) for a in np.arange(0, 180, 180 / K)]
    scale = gold[0].sum() / mine[0].sum()
    for k in range(K):
        assert rel_l2(scale * mine[k], gold[k].reshape(mine[k].shape)) < 1e-12


def check_backend(base):
    from rescan_line_sted_b200 import orientations
    unit = base / base.sum()
    many = orientations.rotate_many(unit, ANGLES)
    for a, got in zip(ANGLES, many):
        ref = orc.rotate_psf(unit, a)
        assert got.shape == ref.shape == (1,) + base.shape[1:] and got.dtype == np.float64
        assert rel_l2(got, ref) < 1e-12, a
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), a
        assert got.min() >= 0 and got.max() <= 1.1 * unit.max()
    # single-angle form, non-square plane, an angle whose corners leave the source
    x = np.random.default_rng(3).random((1, 31, 45))
    for a in (0, 90, 33.3, 200.0):
        assert rel_l2(orientations.rotate(x, a), orc.rotate_psf(x, a)) < 1e-12, a
    psfs = orientations.line_orientation_psfs(base, 8, 3.0227)
    ref = orc.orientation_psfs(base, 8, 3.0227)
    assert len(psfs) == 8
    for p, r in zip(psfs, ref):
        assert rel_l2(p, r) < 1e-12


def test_kernel_body_on_cpu_replay(base, monkeypatch):
    monkeypatch.setattr(_lib, '_library', emul_support.emulator_library())
    check_backend(base)


@pytest.mark.gpu
def test_kernel_on_gpu(base):
    check_backend(base)
