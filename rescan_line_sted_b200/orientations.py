"""Orientation step between `psf_report` and `Deconvolver`.

The reference's figure-2 driver turns one system PSF into K line orientations
with a local helper (figure_generation/line_sted_figure_2.py:264-272, used at
:244-247):

    def rotate(x, degrees):
        0 -> x; 90 -> np.rot90; else np.clip(scipy.ndimage.rotate(
            x, angle=degrees, axes=(1, 2), reshape=False), 0, 1.1 * x.max())

`rotate` here has the same signature and results; the cubic-spline prefilter
and the interpolation run on the GPU (`lsted_psf_rotate`, one CTA per
orientation), `rotate_many` does all K angles in one launch.  No CPU fallback.
"""
import numpy as np
from scipy.special import cosdg, sindg

from . import _lib
from .line_sted_tools import _device


def _xform(degrees, shape):
    """Matrix and offset scipy.ndimage.rotate builds for one plane
    (scipy/ndimage/_interpolation.py:rotate: [[c, s], [-s, c]] about the centre)."""
    c, s = cosdg(degrees), sindg(degrees)
    m = np.array([[c, s], [-s, c]], dtype=np.float64)
    centre = (np.asarray(shape, dtype=np.float64) - 1) / 2
    off = centre - m @ centre
    return [m[0, 0], m[0, 1], m[1, 0], m[1, 1], off[0], off[1]]


def rotate_many(x, angles):
    """[rotate(x, a) for a in angles] with one kernel launch for all angles that
    need interpolation.  x: (1, n0, n1) float array (as the reference passes it)."""
    x = np.asarray(x)
    assert x.ndim == 3, "expected a (planes, rows, columns) array like the reference's PSFs"
    out = [None] * len(angles)
    todo = []
    for k, a in enumerate(angles):
        if a == 0:
            out[k] = x
        elif a == 90:
            out[k] = np.rot90(np.squeeze(x)).reshape(x.shape)
        else:
            todo.append(k)
    if todo:
        lib = _lib.get()
        p = _lib.c_double_p
        planes = np.ascontiguousarray(x, dtype=np.float64)
        n0, n1 = planes.shape[1:]
        xf = np.ascontiguousarray([_xform(angles[k], (n0, n1)) for k in todo], dtype=np.float64)
        clip_hi = float(1.1 * x.max())
        res = np.empty((planes.shape[0], len(todo), n0, n1), dtype=np.float64)
        for i in range(planes.shape[0]):   # the rotation acts on every plane of axis 0
            lib.call('lsted_psf_rotate', _device(), len(todo), n0, n1,
                     planes[i].ctypes.data_as(p), xf.ctypes.data_as(p), clip_hi,
                     res[i].ctypes.data_as(p))
        for j, k in enumerate(todo):
            out[k] = np.ascontiguousarray(res[:, j])
    return out


def rotate(x, degrees):
    """Drop-in for line_sted_figure_2.py:264-272."""
    return rotate_many(x, [degrees])[0]


def line_orientation_psfs(fine_psf, num_orientations, expected_emission):
    """line_sted_figure_2.py:235-247: the K PSFs handed to Deconvolver."""
    unit = fine_psf / fine_psf.sum()
    angles = list(np.arange(0, 180, 180 / num_orientations))
    return [1 / num_orientations * expected_emission * r
            for r in rotate_many(unit, angles)]
