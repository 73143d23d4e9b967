"""The host-side mirror of the reference interface (names, return shapes,
prints, error behaviour) exercised on the CPU replay of the kernel bodies."""
import json
import os
import pickle

import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import _lib


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.fixture()
def st(monkeypatch):
    from rescan_line_sted_b200 import line_sted_tools
    monkeypatch.setattr(_lib, '_library', emul_support.emulator_library())
    monkeypatch.setenv('LSTED_PRECISION', 'fp64')
    return line_sted_tools


def test_public_surface_matches_reference(st):
    import inspect
    want = {
        'psf_report': ['psf_type', 'excitation_brightness', 'depletion_brightness',
                       'steps_per_excitation_psf_width', 'pulses_per_position',
                       'verbose', 'output_dir'],
        'generate_psfs': ['shape', 'excitation_brightness', 'depletion_brightness',
                          'blur_sigma', 'psf_type', 'output_dir', 'verbose'],
        'tune_psf': ['psf_type', 'scan_type', 'desired_resolution_improvement',
                     'desired_emissions_per_molecule', 'max_excitation_brightness',
                     'steps_per_improved_psf_width', 'relative_error',
                     'verbose_results', 'verbose_iterations'],
        'logarithmic_progress': ['iterable', 'verbose'],
        'get_width': ['x'],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(st, name)).parameters) == params
    sig = inspect.signature(st.Deconvolver.__init__)
    assert list(sig.parameters) == ['self', 'psfs', 'output_prefix', 'verbose']
    sig = inspect.signature(st.Deconvolver.create_data_from_object)
    assert list(sig.parameters) == ['self', 'obj', 'total_brightness', 'random_seed']
    for m in ('load_data_from_tif', 'iterate', 'record_iteration', 'record_data', 'H', 'H_t'):
        assert callable(getattr(st.Deconvolver, m))
    assert inspect.signature(st.tune_psf).parameters['max_excitation_brightness'].default == 0.5
    assert inspect.signature(st.Deconvolver.H_t).parameters['normalize'].default is True


def test_psf_report_golden_and_prints(st, golden_dir, capsys):
    with open(os.path.join(golden_dir, 'scalars.json')) as f:
        scalars = json.load(f)
    store = np.load(os.path.join(golden_dir, 'psf_reports.npz'))
    for name, g in scalars['psf_report'].items():
        rep = st.psf_report(*g['args'], verbose=False)
        for k in ('excitation_dose', 'depletion_dose', 'expected_emission'):
            assert rep[k] == pytest.approx(g[k], rel=1e-12)
        for k, v in rep['psfs'].items():
            assert v.shape == store[name + '/' + k].shape and v.flags.c_contiguous
            assert rel_l2(v, store[name + '/' + k]) < 1e-12
        pickle.loads(pickle.dumps(rep))            # figure 2 pickles these
    assert capsys.readouterr().out == ''
    st.psf_report('line', 1, 9, 8, 1)
    out = capsys.readouterr().out
    assert ' Ideal line rescan ratio: 13.21440\n Neareset integer: 13\n' in out
    assert 'Excitation dose: 8.516 half-saturations' in out
    assert 'Expected emissions per molecule: 1.2341\n\n' in out
    with pytest.raises(ValueError):
        st.generate_psfs((1, 9, 11), 1, 1, 2.0)


def test_tune_psf_golden(st, golden_dir):
    with open(os.path.join(golden_dir, 'scalars.json')) as f:
        g = json.load(f)['tune_psf']['line_rescanned_R2p2_E3']
    res = st.tune_psf(**g['kwargs'])
    for k, v in g.items():
        if k != 'kwargs':
            assert res[k] == pytest.approx(v, rel=1e-5), k
    with pytest.raises(AssertionError):
        st.tune_psf('point', 'rescanned', 2., 4.)


@pytest.mark.parametrize('psf_type,steps', [('line', 8), ('point', 8), ('line', 12)])
def test_device_fit_restates_scipy(st, monkeypatch, psf_type, steps):
    """The fused single-launch psf_report: its widths come from a step-for-step restatement
    of MINPACK's lmdif (what scipy's curve_fit runs) inside the kernel: same iterates, same
    stopping point.  Given the same row it reproduces scipy bit for bit (test below); through
    the whole report (rows that differ in the last place between the two kernels) no
    resolution factor of the sweep grid is off by more than 1e-9; sub-pixel profiles go to
    the host fit."""
    import _psf_fit_cases as cases
    fallbacks, exact = cases.check_fused_reports_against_host_fit(
        st, monkeypatch, psf_type, steps, tol_R=1e-9, min_exact=0)
    print('fallbacks %d, bit-identical resolution factors %d' % (fallbacks, exact))
    # (steps = 8: the two strongest depletions of the grid give sub-0.35-pixel profiles)
    assert fallbacks <= 16


def test_device_lmdif_follows_scipy(st):
    """The in-kernel fit against scipy.optimize.curve_fit on the same rows: same iterates and
    stopping point, so the coefficients agree to ~1e-10 (a Python transcription of the same
    restatement that calls numpy's exp is bit-identical to scipy; libm's and CUDA's exp differ
    from numpy's in the last place, which the forward-difference Jacobian amplifies by
    1/sqrt(eps)) -- against up to 7e-6 between scipy's stopping point and the true minimum.
    Gaussian, saturated, doughnut-suppressed, noisy and off-centre profiles, several lengths."""
    from scipy.optimize import curve_fit
    rng = np.random.default_rng(12)
    for n in (17, 35, 107, 135):
        x = np.arange(n)
        rows = [np.exp(-(x - n // 2) ** 2 / (2 * 2.5 ** 2)),
                0.3 * np.exp(-(x - n / 2 - 1.3) ** 2 / (2 * 4.0 ** 2)),
                (1 - 2 ** -(8 * np.exp(-(x - n // 2) ** 2 / 18.))) *
                2 ** -(9 * (1 - np.exp(-(x - n // 2) ** 2 / 30.))),
                np.exp(-(x - n // 2) ** 2 / (2 * 1.1 ** 2)) + 0.01 * rng.random(n),
                2 * np.exp(-np.abs(x - n // 2) / 3.0)]
        widths, coeff = st.get_width_batch(np.array(rows))
        for r, w, c in zip(rows, widths, coeff):
            ref, _ = curve_fit(lambda t, A, mu, s: A * np.exp(-(t - mu) ** 2 / (2. * s ** 2)),
                               range(n), r, p0=[1., n / 2., 1.])
            assert np.allclose(c, ref, rtol=2e-9, atol=0) and w == c[2], (n, c, ref)


def test_tune_psf_batch_equals_sequential(st):
    import _psf_fit_cases as cases
    cases.check_tune_psf_batch(st, [
        dict(psf_type='point', scan_type='descanned', desired_resolution_improvement=2.0,
             desired_emissions_per_molecule=4.0),
        dict(psf_type='point', scan_type='descanned', desired_resolution_improvement=1.5,
             desired_emissions_per_molecule=4.0),
        dict(psf_type='line', scan_type='descanned', desired_resolution_improvement=1.7,
             desired_emissions_per_molecule=3.0, steps_per_improved_psf_width=2.0),
    ])


def test_psf_report_batch_mixed_samplings(st):
    """One launch, every point its own grid size."""
    reps = st.psf_report_batch('line', [1, 0.5, 2], [9, 3, 27], [8, 12, 6], [1, 2, 3])
    for rep, args in zip(reps, [(1, 9, 8, 1), (0.5, 3, 12, 2), (2, 27, 6, 3)]):
        one = st.psf_report('line', *args, verbose=False)
        assert rep['pulses_per_position'] == args[3]
        for k in one:
            if k == 'psfs':
                for kk in one[k]:
                    assert np.array_equal(rep[k][kk], one[k][kk])
            else:
                assert rep[k] == one[k], k
    scal = st.psf_report_batch('point', [1, 2], 9, 8, 1, psfs=False)
    assert 'psfs' not in scal[0] and scal[1]['expected_emission'] > 0


def test_deconvolver_mirror(st, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, 'fig2_2p0x_lr.npz'))
    d = st.Deconvolver([p[None] for p in g['psfs']],
                       output_prefix=str(tmp_path / 'sub') + '/run_', verbose=False)
    assert os.path.isdir(str(tmp_path / 'sub'))     # mkdir of the prefix dir (ref:487)
    assert not hasattr(d, 'estimate')
    d.create_data_from_object(g['object_u8'].astype(np.float64), total_brightness=5e10,
                              random_seed=0)
    assert isinstance(d.noiseless_measurement, list) and len(d.noisy_measurement) == 4
    assert d.noisy_measurement[0].shape == (1, 128, 128)
    d.noisy_measurement = [m[None] for m in g['noisy']]
    for _ in range(8):
        d.iterate()
    assert rel_l2(d.estimate, g['estimate_8']) < 1e-11
    d.record_data()
    d.record_iteration()
    from rescan_line_sted_b200 import np_tif
    hist = np_tif.tif_to_array(str(tmp_path / 'sub') + '/run_estimate_history.tif')
    assert hist.shape == (1, 128, 128) and hist.dtype == np.float32
    assert d.saved_iterations == [8] and len(d.estimate_history) == 1
    # the error-spectrum history (ref:539-548) after more saves equals a from-scratch
    # transform of the whole history (the mirror only transforms the new estimates)
    d.iterate()
    d.record_iteration()
    d.iterate()
    d.record_iteration()
    spec = np_tif.tif_to_array(str(tmp_path / 'sub') + '/run_estimate_FT_error_history.tif')
    eh = np.concatenate(d.estimate_history, axis=0)
    want = np.log(1 + np.abs(np.fft.fftshift(np.fft.fftn(eh - d.true_object, axes=(1, 2)),
                                             axes=(1, 2))))
    assert spec.shape == (3, 128, 128)
    assert np.allclose(spec, want.astype(np.float32), rtol=1e-6)
    d.create_data_from_object(g['object_u8'].astype(np.float64) * 2, total_brightness=5e10,
                              random_seed=1)                      # new object: cache restarts
    d.estimate_history, d.saved_iterations = [], []
    d.iterate()
    d.record_iteration()
    spec = np_tif.tif_to_array(str(tmp_path / 'sub') + '/run_estimate_FT_error_history.tif')
    want = np.log(1 + np.abs(np.fft.fftshift(np.fft.fftn(d.estimate - d.true_object, axes=(1, 2)),
                                             axes=(1, 2))))
    assert spec.shape == (1, 128, 128) and np.allclose(spec, want.astype(np.float32), rtol=1e-6)
    # load_data_from_tif round trip (dead code in the reference, works here)
    d2 = st.Deconvolver([p[None] for p in g['psfs']], output_prefix=str(tmp_path) + '/l_',
                        verbose=False)
    d2.load_data_from_tif(str(tmp_path / 'sub') + '/run_noisy_measurement.tif')
    assert np.allclose(d2.noisy_measurement[1], g['noisy'][1][None], rtol=1e-6)


def test_logarithmic_progress(st, capsys):
    for n in (0, 1, 2, 3, 9, 17, 1025):
        flags = [s for _, s in st.logarithmic_progress(range(n), verbose=False)]
        assert [i for i, s in enumerate(flags) if s] == orc.logarithmic_save_points(n)
    assert [x for x, _ in st.logarithmic_progress(range(5))] == list(range(5))


def test_np_tif_round_trip(tmp_path):
    from rescan_line_sted_b200 import np_tif
    rng = np.random.default_rng(0)
    for dtype in (np.uint8, np.uint16, np.uint32, np.int16, np.float32, np.float64):
        a = (rng.random((3, 5, 7)) * 200).astype(dtype)
        fn = str(tmp_path / ('t_%s.tif' % np.dtype(dtype).name))
        np_tif.array_to_tif(a, fn)
        b = np_tif.tif_to_array(fn)
        assert b.shape == a.shape
        assert np.allclose(a, b)
        assert b.dtype == (np.float32 if dtype == np.float64 else dtype)
    np_tif.array_to_tif(rng.random((4, 6)), str(tmp_path / '2d.tif'))
    assert np_tif.tif_to_array(str(tmp_path / '2d.tif')).shape == (1, 4, 6)
    with open(str(tmp_path / 'bad.tif'), 'wb') as f:
        f.write(b'not a tiff at all')
    with pytest.raises(UserWarning):
        np_tif.tif_to_array(str(tmp_path / 'bad.tif'))
