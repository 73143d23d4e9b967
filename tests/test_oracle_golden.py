"""The oracle must reproduce the committed outputs of the unmodified
reference (tests/golden/, made by make_golden.py) on any machine."""
import json
import os

import numpy as np
import pytest

from oracle import line_sted_oracle as orc


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.fixture(scope='module')
def scalars(golden_dir):
    with open(os.path.join(golden_dir, 'scalars.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def psf_store(golden_dir):
    return np.load(os.path.join(golden_dir, 'psf_reports.npz'))


@pytest.fixture(scope='module')
def fig2(golden_dir):
    return np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))


@pytest.mark.parametrize('closed_form', [False, True])
def test_psf_report_matches_reference_golden(scalars, psf_store, closed_form):
    for name, g in scalars['psf_report'].items():
        rep = orc.psf_report(*g['args'], use_closed_form=closed_form)
        for k in ('excitation_dose', 'depletion_dose', 'expected_emission'):
            assert rep[k] == pytest.approx(g[k], rel=1e-12), (name, k)
        for k in ('resolution_improvement_descanned',
                  'resolution_improvement_rescanned'):
            if k in g:  # LM fit tolerance 1.5e-8 limits this one
                assert rep[k] == pytest.approx(g[k], rel=1e-6), (name, k)
        for arr_name, arr in rep['psfs'].items():
            gold = psf_store['%s/%s' % (name, arr_name)]
            assert arr.shape == gold.shape
            assert rel_l2(arr, gold) < 1e-12, (name, arr_name)


def test_published_numbers(scalars):
    pub = scalars['published']['point_1_9_8_1']
    rep = orc.psf_report('point', 1, 9, 8, 1)
    assert round(rep['excitation_dose'], 2) == pub['excitation_dose']
    assert round(rep['depletion_dose'], 2) == pub['depletion_dose']
    assert round(rep['expected_emission'], 2) == pub['expected_emission']
    assert round(rep['resolution_improvement_descanned'], 1) == pub['resolution']


def test_tune_psf_point_matches_reference_golden(scalars):
    g = scalars['tune_psf']['point_R2_E4']
    res = orc.tune_psf(**g['kwargs'])
    for k in ('excitation_dose', 'depletion_dose', 'expected_emission',
              'excitation_brightness', 'depletion_brightness',
              'pulses_per_position', 'resolution_improvement_descanned'):
        assert res[k] == pytest.approx(g[k], rel=1e-5), k
    # published appendix row R=2.0: doses 18.8 / 1696.8, emissions 4.00
    row = scalars['published']['fig2_point_rows']['2.0']
    assert round(res['excitation_dose'], 1) == row[0]
    assert round(res['depletion_dose'], 1) == row[1]
    assert round(res['expected_emission'], 2) == 4.00


def test_forward_model_and_rl_match_reference_golden(fig2, scalars):
    g = scalars['fig2_2p0x_lr']
    base = orc.psf_report('line', use_closed_form=True,
                          **g['psf_report_args'])['psfs']['rescan_sted']
    assert rel_l2(base, fig2['base_psf']) < 1e-12
    psfs = orc.orientation_psfs(base, g['K'], 3.0227)
    assert rel_l2(np.concatenate(psfs, 0), fig2['psfs']) < 1e-12
    for engine in ('scipy', 'numpy'):
        d = orc.Deconvolver([p[None] for p in fig2['psfs']], engine=engine)
        d.create_data_from_object(fig2['object_u8'].astype(np.float64),
                                  total_brightness=g['total_brightness'],
                                  random_seed=g['seed'])
        assert rel_l2(np.concatenate(d.noiseless_measurement, 0),
                      fig2['noiseless']) < 1e-13
        # legacy MT19937 Poisson is deterministic given the seed
        assert np.array_equal(np.concatenate(d.noisy_measurement, 0),
                              fig2['noisy'])
        assert d.noisy_measurement[0].sum() == pytest.approx(
            g['noisy0_sum'], rel=1e-14)
        d.iterate()
        assert rel_l2(d.estimate, fig2['estimate_1']) < 1e-12
        assert rel_l2(d.H_t_normalization, fig2['H_t_normalization']) < 1e-13
        for _ in range(7):
            d.iterate()
        assert rel_l2(d.estimate, fig2['estimate_8']) < 1e-11
        assert d.estimate.sum() == pytest.approx(g['estimate8_sum'], rel=1e-11)


def test_restated_primitives_match_scipy():
    from scipy.ndimage import gaussian_filter
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(5)
    a = rng.random((1, 17, 23))
    for sigma in (1.7, (0, 0, 3.4), 4.9):  # radius > half width on purpose
        assert np.abs(gaussian_filter(a, sigma) -
                      orc.gaussian_blur_nd(a, sigma)).max() < 1e-14
    for shape_x, shape_p in (((1, 31, 40), (1, 7, 9)), ((1, 16, 16), (1, 8, 6)),
                             ((1, 5, 9), (1, 11, 11))):
        x, p = rng.random(shape_x), rng.random(shape_p)
        assert np.abs(fftconvolve(x, p, 'same') -
                      orc.fftconvolve_same(x, p)).max() < 1e-12


def test_closed_form_rescan_equals_scan_loop():
    for args in (('line', 1, 9, 8, 1), ('line', 0.3, 20, 10, 2)):
        a = orc.psf_report(*args, use_closed_form=False)
        b = orc.psf_report(*args, use_closed_form=True)
        assert a['_line_rescan_ratio'] == b['_line_rescan_ratio']
        for k in ('rescan_sted', 'descan_sted'):
            assert rel_l2(b['psfs'][k], a['psfs'][k]) < 1e-13


def test_logarithmic_save_points():
    assert orc.logarithmic_save_points(2 ** 10 + 1) == [2 ** i for i in range(10)] + [1024]
    assert orc.logarithmic_save_points(0) == []
    assert orc.logarithmic_save_points(1) == [0]
