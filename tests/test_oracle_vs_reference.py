"""Pin the oracle against the unmodified reference, run side by side in the
build container.  Skipped where /root/reference is absent (GPU box)."""
import numpy as np
import pytest

from _reference_loader import load_reference, reference_available
from oracle import line_sted_oracle as orc

pytestmark = pytest.mark.skipif(not reference_available(),
                                reason='reference sources not present')


@pytest.fixture(scope='module')
def ref():
    return load_reference()


def rel_l2(a, b):
    den = np.linalg.norm(np.ravel(b))
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / (den if den > 0 else 1.0)


@pytest.mark.parametrize('args', [('point', 1, 9, 8, 1), ('line', 1, 9, 8, 1),
                                  ('line', 0.05, 27, 8, 3), ('point', 4, 54, 6, 2),
                                  ('line', 2, 0, 8, 1)])
def test_psf_report_side_by_side(ref, args):
    r = ref.psf_report(*args, verbose=False)
    for closed in (False, True):
        o = orc.psf_report(*args, use_closed_form=closed)
        for k, v in r.items():
            if k == 'psfs':
                for name, arr in v.items():
                    assert rel_l2(o['psfs'][name], arr) < 1e-12, name
            elif k.startswith('resolution'):
                assert o[k] == pytest.approx(v, rel=1e-6)
            else:
                assert o[k] == pytest.approx(v, rel=1e-12)


def test_get_width_and_progress(ref):
    x = np.exp(-(np.arange(41) - 20.3) ** 2 / (2 * 3.3 ** 2)) * 0.7
    w_r, fit_r = ref.get_width(x)
    w_o, fit_o = orc.gaussian_fit_width(x)
    assert w_o == pytest.approx(w_r, rel=1e-12)
    assert np.allclose(fit_o, fit_r, rtol=0, atol=1e-14)
    for n in (0, 1, 2, 3, 9, 17, 1025):
        flags = [s for _, s in ref.logarithmic_progress(range(n), verbose=False)]
        assert [i for i, s in enumerate(flags) if s] == orc.logarithmic_save_points(n)


def test_deconvolver_side_by_side(ref, tmp_path):
    rng = np.random.default_rng(11)
    psfs = [rng.random((1, 9, 9)), rng.random((1, 9, 9)), rng.random((1, 7, 11))]
    obj = rng.random((1, 33, 40)) + 0.1
    r = ref.Deconvolver(psfs, output_prefix=str(tmp_path) + '/x_', verbose=False)
    o = orc.Deconvolver(psfs, engine='numpy')
    r.create_data_from_object(obj, total_brightness=1e6, random_seed=3)
    o.create_data_from_object(obj, total_brightness=1e6, random_seed=3)
    for a, b in zip(o.noiseless_measurement, r.noiseless_measurement):
        assert rel_l2(a, b) < 1e-13
    for a, b in zip(o.noisy_measurement, r.noisy_measurement):
        assert np.array_equal(a, b)
    for _ in range(5):
        r.iterate(), o.iterate()
    assert rel_l2(o.estimate, r.estimate) < 1e-11
    assert rel_l2(o.H_t_normalization, r.H_t_normalization) < 1e-13
    y = [rng.random((1, 33, 40)) for _ in psfs]
    assert rel_l2(o.H_t(y, normalize=False), r.H_t(y, normalize=False)) < 1e-13
