"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the explicit scan-position simulation of
figure 3: `simulate_imaging`, figure_generation/line_sted_figure_3.py:76-273,
with its helpers `rotate` :382-391, `shift` :393-396, `scale_y` :398-409.  Only tests/ may import
this module; the product path (rescan_line_sted_b200/scan_engine.py + liblsted.so) never does.

Same scipy calls as the reference (cubic-spline `shift` / `rotate` / `zoom`, `gaussian_filter`),
same two-pass structure (first pass: display maxima on orientation 0; second pass: every
orientation), but the arrays the reference only hands to its figure code are returned.
Pinned: tests/golden/fig3_*.npz hold frames captured from the UNMODIFIED reference function
(tests/golden/make_golden_fig3.py runs it with the plotting calls stubbed)."""
import warnings

import numpy as np
from scipy.ndimage import gaussian_filter
from scipy.ndimage import rotate as nd_rotate, shift as nd_shift, zoom as nd_zoom

IMAGING_TYPES = ('descan_point', 'nondescan_multipoint', 'descan_line', 'rescan_line')


def rotate(x, angle_degrees):                                          # :382-391
    if angle_degrees == 0:
        return x.copy()
    return np.clip(nd_rotate(x, angle=angle_degrees, axes=(1, 2), mode='nearest', reshape=False),
                   0, 1.1 * x.max())


def shift(x, s):                                                       # :393-396
    return np.clip(nd_shift(x, s), 0, 1.1 * x.max())


def scale_y(x, scaling_factor):                                        # :398-409
    original_shape = x.shape
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        scaled = nd_zoom(x[0, :, :], zoom=(scaling_factor, 1))
    y_dif = original_shape[-2] - scaled.shape[-2]
    return np.pad(scaled, ((y_dif // 2, y_dif - y_dif // 2), (0, 0)), 'constant').reshape(original_shape)


def scan_plan(shape, imaging_type, psf_width, R, pad):
    """Step, scan positions, spot separation (:105-137): pure integer arithmetic."""
    _, n_y, n_x = shape
    step = int(np.round(psf_width / (4 * R)))
    exc_sep = None
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


def frames_to_keep(num_positions):
    """Indices of the scan positions the reference draws (:246-249)."""
    skip = max(int(np.round(num_positions / 150)), 1)
    return [i for i in range(num_positions) if i % skip == 0 or i == num_positions - 1]


def simulate_imaging(obj, imaging_type, psf_width, R, num_orientations, pulses_per_position, pad,
                     keep_frames=True):
    assert len(obj.shape) == 3 and obj.shape[0] == 1 and imaging_type in IMAGING_TYPES
    psf_sigma = psf_width / (2 * np.sqrt(2 * np.log(2)))
    _, n_y, n_x = obj.shape
    step, scan_positions, exc_sep = scan_plan(obj.shape, imaging_type, psf_width, R, pad)
    obj = np.pad(obj, ((0, 0), (pad, pad), (pad, pad)), 'constant')
    centered_exc = np.zeros(obj.shape)
    if imaging_type in ('descan_line', 'rescan_line'):
        centered_exc[0, obj.shape[1] // 2, :] = 1
        sted_sigma = (0, psf_sigma / R, 0)
    elif imaging_type == 'descan_point':
        centered_exc[0, obj.shape[1] // 2, obj.shape[2] // 2] = 1
        sted_sigma = (0, psf_sigma / R, psf_sigma / R)
        num_orientations = 1
    else:
        centered_exc[0, pad:-pad:exc_sep, pad:-pad:exc_sep] = 1
        sted_sigma = (0, psf_sigma / R, psf_sigma / R)
        num_orientations = 1
    centered_exc = gaussian_filter(centered_exc, sted_sigma, truncate=8)
    crop = (0, slice(pad, -pad), slice(pad, -pad))
    mx = dict(exc=centered_exc[crop].max(), glow=0, inst_sig=0, cum_sig=0, reconst=0, new_sig=0)
    out = dict(step=step, scan_positions=scan_positions, exc_sep=exc_sep, orientations=[], frames=[])
    keep = set(frames_to_keep(len(scan_positions)))
    for which_run in ('find_maxima', 'generate_figures'):
        camera_exposures, pulses_delivered = 0, 0
        for rot in np.arange(0, 180, 180 / num_orientations)[::-1]:
            if which_run == 'find_maxima' and rot > 0:
                continue
            rot_obj = rotate(obj, rot)
            cum_detector_sig = np.zeros(obj.shape)
            reconstruction = np.zeros(obj.shape)
            for which_pos, (shift_y, shift_x) in enumerate(scan_positions):
                pulses_delivered += pulses_per_position
                last_reconstruction = reconstruction.copy()
                exc = shift(centered_exc, (0, shift_y, shift_x))
                glow = rot_obj * exc
                descanned_glow = shift(glow, (0, -shift_y, -shift_x))
                if imaging_type in ('descan_line', 'descan_point'):
                    inst_detector_sig = gaussian_filter(descanned_glow, psf_sigma)
                    cum_detector_sig = inst_detector_sig
                    y0 = shift_y + n_y // 2 + pad
                    if imaging_type == 'descan_line':
                        reconstruction[0, y0:y0 + step, :] = inst_detector_sig.sum(axis=1, keepdims=True)
                        camera_exposures += 1
                    else:
                        x0 = shift_x + n_x // 2 + pad
                        reconstruction[0, y0:y0 + step, x0:x0 + step] = inst_detector_sig.sum()
                        camera_exposures = 'N/A'
                elif imaging_type == 'nondescan_multipoint':
                    inst_detector_sig = gaussian_filter(glow, psf_sigma)
                    cum_detector_sig = inst_detector_sig
                    for y_sp in range(pad + shift_y, pad + shift_y + n_y, exc_sep):
                        for x_sp in range(pad + shift_x, pad + shift_x + n_x, exc_sep):
                            region = inst_detector_sig[0, max(y_sp - exc_sep // 3, 0):y_sp + exc_sep // 3,
                                                       max(x_sp - exc_sep // 3, 0):x_sp + exc_sep // 3]
                            reconstruction[0, y_sp - step // 2:y_sp - step // 2 + step,
                                           x_sp - step // 2:x_sp - step // 2 + step] = region.sum()
                    camera_exposures += 1
                else:
                    scaled = scale_y(gaussian_filter(descanned_glow, psf_sigma),
                                     scaling_factor=1 / (R ** 2 + 1))
                    inst_detector_sig = shift(scaled, (0, shift_y, shift_x))
                    cum_detector_sig += inst_detector_sig
                    if (shift_y, shift_x) == scan_positions[-1]:
                        reconstruction = cum_detector_sig
                        camera_exposures += 1
                new_signal = reconstruction - last_reconstruction
                if which_run == 'find_maxima':
                    mx['glow'] = max(glow.max(), mx['glow'])
                    mx['inst_sig'] = max(inst_detector_sig.max(), mx['inst_sig'])
                    mx['cum_sig'] = max(cum_detector_sig.max(), mx['cum_sig'])
                    mx['reconst'] = max(reconstruction.max(), mx['reconst'])
                    mx['new_sig'] = max(new_signal.max(), mx['new_sig'])
                elif keep_frames and which_pos in keep:
                    out['frames'].append(dict(
                        rot=float(rot), which_pos=which_pos,
                        excitation=rotate(exc, -rot)[crop] / mx['exc'],
                        glow=rotate(glow, -rot)[crop] / mx['glow'],
                        inst_sig=inst_detector_sig[crop] / mx['inst_sig'],
                        cum_sig=cum_detector_sig[crop] / mx['cum_sig'],
                        new_sig=new_signal[crop] / mx['new_sig'],
                        reconstruction=reconstruction[crop] / mx['reconst'],
                        pulses_delivered=pulses_delivered, camera_exposures=camera_exposures))
            if which_run == 'generate_figures':
                out['orientations'].append(dict(
                    rot=float(rot), reconstruction=reconstruction.copy(),
                    cum_detector_sig=np.array(cum_detector_sig), pulses_delivered=pulses_delivered,
                    camera_exposures=camera_exposures))
    out['maxima'] = {k: float(v) for k, v in mx.items()}
    out['padded_shape'] = obj.shape
    return out
