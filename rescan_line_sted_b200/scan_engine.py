"""Figure-3 scan-position engine on the GPU.

Drop-in for the simulation half of the reference's figure-3 script
(figure_generation/line_sted_figure_3.py): `simulate_imaging` :76-273 and its helpers
`rotate` :382-391, `shift` :393-396, `scale_y` :398-409.  Same names, argument order and
meaning; the per-scan-position loop runs in liblsted.so (`csrc/scan_kernels.cuh`) with every
scan position of an orientation in flight at once.  No CPU fallback: without the CUDA
library or a GPU every call raises.

The reference function draws through `generate_figure` and returns nothing.  Here the figure
code is a callback (`generate_figure=`, called with exactly the reference's positional
arguments, in the reference's order of orientations and positions) and the numbers the
reference only prints or draws are returned as a dict.
"""
import ctypes
import os
import warnings

import numpy as np
from scipy.special import cosdg, sindg

from . import _lib
from .line_sted_tools import _device, _gaussian_taps

IMAGING_TYPES = ('descan_point', 'nondescan_multipoint', 'descan_line', 'rescan_line')
_P = _lib.c_double_p


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def _rotation_xform(angle_degrees, shape):
    """Matrix and offset scipy.ndimage.rotate(reshape=False) builds for one plane."""
    c, s = cosdg(angle_degrees), sindg(angle_degrees)
    m = np.array([[c, s], [-s, c]], dtype=np.float64)
    centre = (np.asarray(shape, dtype=np.float64) - 1) / 2
    off = centre - m @ centre
    return np.array([m[0, 0], m[0, 1], m[1, 0], m[1, 1], off[0], off[1]], dtype=np.float64)


def _spline(planes, xforms, out_shape, mode, clip):
    planes = np.ascontiguousarray(planes, dtype=np.float64)
    xforms = np.ascontiguousarray(xforms, dtype=np.float64).reshape(planes.shape[0], 6)
    out = np.empty((planes.shape[0],) + tuple(out_shape), dtype=np.float64)
    _lib.get().call('lsted_img_spline', _device(), planes.shape[0], planes.shape[1],
                    planes.shape[2], _ptr(planes), _ptr(xforms), out_shape[0], out_shape[1],
                    {'constant': 0, 'nearest': 1}[mode], int(bool(clip)), _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# The three resampling helpers of the figure-3 script
# ---------------------------------------------------------------------------------------------
def rotate(x, angle_degrees):
    """line_sted_figure_3.py:382-391: cubic-spline rotation of every plane of a (planes, rows,
    columns) array about its centre, mode='nearest', clipped to [0, 1.1 max]."""
    x = np.asarray(x)
    assert x.ndim == 3
    if angle_degrees == 0:
        return x.copy()
    xf = _rotation_xform(angle_degrees, x.shape[1:])
    out = _spline(x, np.tile(xf, (x.shape[0], 1)), x.shape[1:], 'nearest', clip=False)
    return np.clip(out, 0, 1.1 * x.max())


def shift(x, shift):
    """line_sted_figure_3.py:393-396: cubic-spline shift (mode='constant'), clipped."""
    x = np.asarray(x)
    assert x.ndim == 3 and len(shift) == 3 and shift[0] == 0
    xf = np.array([1, 0, 0, 1, -float(shift[1]), -float(shift[2])], dtype=np.float64)
    out = _spline(x, np.tile(xf, (x.shape[0], 1)), x.shape[1:], 'constant', clip=False)
    return np.clip(out, 0, 1.1 * x.max())


def scale_y(x, scaling_factor):
    """line_sted_figure_3.py:398-409: zoom rows by `scaling_factor`, zero-pad back."""
    x = np.asarray(x)
    assert len(x.shape) == 3 and x.shape[0] == 1 and x.shape[1] > 1
    assert float(scaling_factor) == scaling_factor
    n0, n1 = x.shape[1:]
    m0 = int(round(n0 * scaling_factor))
    zoom = (n0 - 1) / (m0 - 1) if m0 > 1 else 1.0
    scaled = _spline(x, [[zoom, 0, 0, 1, 0, 0]], (m0, n1), 'constant', clip=False)[0]
    y_dif = n0 - m0
    return np.pad(scaled, ((y_dif // 2, y_dif - y_dif // 2), (0, 0)), 'constant').reshape(x.shape)


def gaussian_filter(x, sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter(mode='reflect') for (1, rows, columns) arrays, as the
    figure-3 script calls it (:139, :173, :202, :221)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.ndim == 3
    sig = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (3,))
    assert x.shape[0] == 1 or sig[0] <= 1e-15, "blur across planes is not supported"
    taps = [(_gaussian_taps(s, truncate) if s > 1e-15 else (None, 0)) for s in sig[1:]]
    out = np.empty_like(x)
    _lib.get().call('lsted_img_gauss', _device(), x.shape[0], x.shape[1], x.shape[2], _ptr(x),
                    _ptr(taps[0][0]), taps[0][1], _ptr(taps[1][0]), taps[1][1], _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# simulate_imaging
# ---------------------------------------------------------------------------------------------
def scan_plan(shape, imaging_type, psf_width, R):
    """Step, scan positions, spot separation (:91, :105-137): integer arithmetic only."""
    _, n_y, n_x = shape
    step = int(np.round(psf_width / (4 * R)))
    assert step >= 1, "psf_width / (4 R) rounds to a zero scan step"
    exc_sep = 0
    if imaging_type in ('descan_line', 'rescan_line'):
        positions = [(int(y), 0) for y in np.arange(-n_y // 2, n_y // 2 + 1, step)]
    elif imaging_type == 'descan_point':
        positions = [(int(y), int(x)) for y in np.arange(-n_y // 2, n_y // 2 + 1, step)
                     for x in np.arange(-n_x // 2, n_x // 2 + 1, step)]
    else:
        exc_sep = int(step * np.round(psf_width * 1.4 / step))
        positions = [(int(y), int(x)) for y in np.arange(0, exc_sep, step)
                     for x in np.arange(0, exc_sep, step)]
    return step, positions, exc_sep


def frames_drawn(num_positions):
    """Scan positions the reference draws: about 150 per orientation plus the last (:246-249)."""
    skip = max(int(np.round(num_positions / 150)), 1)
    return [i for i in range(num_positions) if i % skip == 0 or i == num_positions - 1]


class ScanHandle:
    """RAII wrapper of lsted_scan_* for one (imaging type, object shape, scan)."""

    def __init__(self, imaging_type, obj_shape, pad, psf_width, R, chunk_bytes=0):
        self.lib = _lib.get()
        self.imaging_type = imaging_type
        _, self.n_y, self.n_x = obj_shape
        self.pad = pad
        self.step, self.positions, self.exc_sep = scan_plan(obj_shape, imaging_type, psf_width, R)
        psf_sigma = psf_width / (2 * np.sqrt(2 * np.log(2)))
        blur_taps, blur_radius = _gaussian_taps(psf_sigma, 4.0)        # :173,:202,:221
        exc_taps, exc_radius = _gaussian_taps(psf_sigma / R, 8.0)      # :139
        prm = _lib.ScanParams(IMAGING_TYPES.index(imaging_type), self.n_y, self.n_x, pad, self.step,
                              self.exc_sep, len(self.positions), 1 / (R ** 2 + 1), blur_radius,
                              exc_radius, chunk_bytes)
        pos = np.ascontiguousarray(self.positions, dtype=np.int32)
        self.padded_shape = (self.n_y + 2 * pad, self.n_x + 2 * pad)
        self.h = ctypes.c_void_p()
        self.lib.call('lsted_scan_create', ctypes.byref(self.h), _device(), ctypes.byref(prm),
                      pos.ctypes.data_as(_lib.c_int_p), _ptr(blur_taps), _ptr(exc_taps))
        self.device_ms = 0.0

    def close(self):
        if self.h:
            self.lib.call('lsted_scan_destroy', self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def excitation(self):
        out = np.empty(self.padded_shape, dtype=np.float64)
        self.lib.call('lsted_scan_excitation', self.h, _ptr(out))
        return out

    def run(self, obj_padded, rot, frame_positions=(), want_images=True):
        """One orientation.  Returns (maxima [P][5], reconstruction, cum_detector_sig)."""
        obj_padded = np.ascontiguousarray(obj_padded, dtype=np.float64).reshape(self.padded_shape)
        xf = None if rot == 0 else _rotation_xform(rot, self.padded_shape)
        fp = np.ascontiguousarray(frame_positions, dtype=np.int32)
        maxima = np.empty((len(self.positions), 5), dtype=np.float64)
        recon = np.empty(self.padded_shape, dtype=np.float64) if want_images else None
        cum = np.empty(self.padded_shape, dtype=np.float64) if want_images else None
        ms = ctypes.c_double()
        self.lib.call('lsted_scan_run', self.h, _ptr(obj_padded), _ptr(xf),
                      fp.ctypes.data_as(_lib.c_int_p) if len(fp) else None, len(fp),
                      _ptr(maxima), _ptr(recon), _ptr(cum),
                      ctypes.byref(ms))
        self.device_ms = ms.value
        return maxima, recon, cum

    def frames(self, first, count, display_max, rot):
        """Display planes [count][6][n_y][n_x] of kept frames first .. first+count-1."""
        dm = np.ascontiguousarray(display_max, dtype=np.float64)
        xf = None if rot == 0 else _rotation_xform(-rot, self.padded_shape)
        out = np.empty((count, 6, self.n_y, self.n_x), dtype=np.float64)
        self.lib.call('lsted_scan_frames', self.h, first, count, _ptr(dm), _ptr(xf), _ptr(out))
        return out


def simulate_imaging(obj, imaging_type, psf_width, R, num_orientations, pulses_per_position, pad,
                     comparison_name='', generate_figure=None, verbose=True, frame_batch=32):
    """line_sted_figure_3.py:76-273 with the scan loop on the GPU.

    generate_figure: None (no frames are formed) or a callable with the reference's signature
    (filename, obj, excitation, glow, instantaneous_detector_signal, cumulative_detector_signal,
    new_signal, reconstruction, pulses_delivered, camera_exposures) :275-286; it is called for
    the same scan positions, in the same order, with the same display scaling as the reference.
    Returns a dict: step, scan_positions, exc_sep, maxima (the display scaling of :141-142,
    :231-235), orientations (per orientation, in the reference's order: rot, reconstruction,
    cum_detector_sig, pulses_delivered, camera_exposures), filenames, device_ms."""
    output_filename = imaging_type + '_'
    if num_orientations > 1:
        output_filename += '%02iangles_' % num_orientations
    output_filename += comparison_name
    if verbose:
        print("\nSimulating:", output_filename)
    obj = np.asarray(obj)
    assert len(obj.shape) == 3 and obj.shape[0] == 1
    assert imaging_type in IMAGING_TYPES
    assert psf_width >= 1
    assert R >= 1
    assert num_orientations >= 1 and int(num_orientations) == num_orientations
    assert pad > 0 and int(pad) == pad
    if imaging_type in ('descan_point', 'nondescan_multipoint'):
        num_orientations = 1                                            # :124,:137
    h = ScanHandle(imaging_type, obj.shape, int(pad), psf_width, R)
    try:
        return _simulate(h, obj, imaging_type, num_orientations, pulses_per_position, int(pad),
                         comparison_name, output_filename, generate_figure, verbose, frame_batch)
    finally:
        h.close()


def _simulate(h, obj, imaging_type, num_orientations, pulses_per_position, pad, comparison_name,
              output_filename, generate_figure, verbose, frame_batch):
    P = len(h.positions)
    padded = np.pad(obj, ((0, 0), (pad, pad), (pad, pad)), 'constant')
    crop = (slice(pad, -pad), slice(pad, -pad))
    centered_exc = h.excitation()
    mx = dict(exc=centered_exc[crop].max(), glow=0, inst_sig=0, cum_sig=0, reconst=0, new_sig=0)
    out = dict(step=h.step, scan_positions=h.positions, exc_sep=h.exc_sep or None,
               orientations=[], filenames=[], device_ms=0.0, centered_exc=centered_exc)
    keep = frames_drawn(P) if generate_figure is not None else []
    obj_display = padded[0][crop] / padded.max()
    rotations = np.arange(0, 180, 180 / num_orientations)[::-1]
    for which_run in ('find_maxima', 'generate_figures'):
        camera_exposures, pulses_delivered = 0, 0
        for rot in rotations:
            if which_run == 'find_maxima' and rot > 0:
                continue
            if verbose:
                print("Orientation:", rot, "degrees")
            frames_wanted = keep if which_run == 'generate_figures' else []
            maxima, recon, cum = h.run(padded, rot, frames_wanted,
                                       want_images=which_run == 'generate_figures')
            out['device_ms'] += h.device_ms
            if which_run == 'find_maxima':
                if verbose:
                    print("Calculating maxima for display scaling...")
                for k, name in enumerate(('glow', 'inst_sig', 'cum_sig', 'reconst', 'new_sig')):
                    mx[name] = max(maxima[:, k].max(), mx[name])
                continue
            # counters exactly as the reference advances them per scan position
            pulses_before = pulses_delivered
            exposures_before = camera_exposures

            def counters(which_pos):
                pulses = pulses_before + (which_pos + 1) * pulses_per_position
                if imaging_type == 'descan_point':
                    return pulses, 'N/A'
                if imaging_type == 'rescan_line':
                    return pulses, exposures_before + (1 if which_pos == P - 1 else 0)
                return pulses, exposures_before + which_pos + 1
            if generate_figure is not None:
                if verbose:
                    print("Generating figures...", end='')
                dm = [mx[k] for k in ('exc', 'glow', 'inst_sig', 'cum_sig', 'new_sig', 'reconst')]
                for first in range(0, len(keep), frame_batch):
                    count = min(frame_batch, len(keep) - first)
                    planes = h.frames(first, count, dm, rot)
                    for f in range(count):
                        which_pos = keep[first + f]
                        filename = os.path.join(
                            os.getcwd(), os.pardir, os.pardir, 'big_images', 'Figure_3_temp',
                            imaging_type + '_%03ideg_' % rot + comparison_name + '_%06i.svg' % which_pos)
                        out['filenames'].append(filename)
                        if which_pos == P - 1:
                            out['filenames'].extend([filename] * 10)
                        if verbose:
                            print('.', end='')
                            if which_pos == P - 1:
                                print()
                        pulses, exposures = counters(which_pos)
                        generate_figure(filename, obj_display, planes[f, 0], planes[f, 1],
                                        planes[f, 2], planes[f, 3], planes[f, 4], planes[f, 5],
                                        pulses, exposures)
            pulses_delivered, camera_exposures = counters(P - 1)
            out['orientations'].append(dict(
                rot=float(rot), reconstruction=recon[None], cum_detector_sig=cum[None],
                pulses_delivered=pulses_delivered, camera_exposures=camera_exposures))
    out['maxima'] = {k: float(v) for k, v in mx.items()}
    out['output_filename'] = output_filename
    return out
