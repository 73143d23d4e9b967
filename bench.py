#!/usr/bin/env python3
"""Headline benchmark: simulated rescan-line-STED frames/s at 2048^2.

One step = one frame(N_iter) of BASELINE.json config 4:
    create_data_from_object (K forward convolutions + Poisson)
    + H_t_normalization + N_iter Richardson-Lucy iterations,
on a synthetic 2048^2 object (astronaut tiled 16x16), K = 16 line
orientations of the figure-2 '2p0x_lr' rescan PSF (107^2), fp32.

  python bench.py [--gpus N] [--steps K] [--warmup W]       this repo (CUDA)
  python bench.py --impl reference ...                      CPU reference arm

With N > 1 (torchrun) every rank simulates and deconvolves its own frames
(independent frames = the reference's own unit of parallel work: figure 2
runs 24 deconvolvers x 4 images); no data-path collective, weak scaling.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FIG2_2P0X_LR = dict(excitation_brightness=0.21672180512595912,
                    depletion_brightness=11.766131198861775,
                    steps_per_excitation_psf_width=25, pulses_per_position=5)
EMISSION_2P0X_LR = 3.0227


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--orientations', type=int, default=16)
    ap.add_argument('--iterations', type=int, default=64)
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'fp64'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--shard', default='frames', choices=['frames', 'orientations'],
                    help='N>1: independent frames per GPU (weak scaling, no collective) or the '
                         'K orientations of every frame split over the GPUs (strong scaling, one '
                         'NCCL all-reduce per RL iteration)')
    return ap.parse_args()


# ---------------------------------------------------------------------------
# Workload (host-side, one-off, outside every timed region)
# ---------------------------------------------------------------------------
def rotate_psf(x, degrees):
    """Orientation step of the caller (line_sted_figure_2.py:264-272)."""
    from scipy.ndimage import rotate
    if degrees == 0:
        return x
    if degrees == 90:
        return np.rot90(np.squeeze(x)).reshape(x.shape)
    return np.clip(rotate(x, angle=degrees, axes=(1, 2), reshape=False),
                   0, 1.1 * x.max())


def orientation_psfs(base, K):
    unit = base / base.sum()
    return [1 / K * EMISSION_2P0X_LR * rotate_psf(unit, a)
            for a in np.arange(0, 180, 180 / K)]


def synthetic_object(N):
    """O1 of SURVEY.md 8d: the 128^2 astronaut target tiled to N x N."""
    tile = np.load(os.path.join(ROOT, 'tests', 'golden', 'fig2_2p0x_lr.npz'))['object_u8']
    reps = (N + 127) // 128
    return np.ascontiguousarray(
        np.tile(tile, (1, reps, reps))[:, :N, :N].astype(np.float64))


def total_brightness(N):
    return 5e10 * (N / 128.0) ** 2


# ---------------------------------------------------------------------------
# Clock sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (same data as the
    nvidia-smi clocks line of B200_PROFILING.md, without spawning a process
    that stalls the driver) every 20 ms from a thread."""
    REASONS = (('hw_slowdown', 0x8), ('sw_thermal_slowdown', 0x20),
               ('hw_thermal_slowdown', 0x40), ('sw_power_cap', 0x4))

    def __init__(self, device):
        self.device, self.rows, self.thread = device, [], None
        self.running, self.error = False, None
        self.t0 = self.t1 = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            index = self.device
            if visible:
                ids = [v for v in visible.split(',') if v.strip()]
                if all(v.strip().isdigit() for v in ids) and self.device < len(ids):
                    index = int(ids[self.device])
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.error = repr(exc)
            return
        self.running = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nvml
        while self.running:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:  # noqa: BLE001
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                except Exception:  # noqa: BLE001
                    reasons = 0
            self.rows.append((time.perf_counter(), float(mhz), int(reasons)))
            time.sleep(0.02)

    def window(self, t0, t1):
        """Only samples taken inside [t0, t1] (the timed region) count."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None,
                    'reasons': ['NVML unavailable: %s' % self.error]}
        self.running = False
        self.thread.join(timeout=1.0)
        rows = [r for r in self.rows
                if self.t0 is None or self.t0 <= r[0] <= self.t1 + 0.03] or self.rows[-1:]
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(n for n, b in self.REASONS if bits & b),
                'samples': len(sm)}


# ---------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the only place bench.py runs oracle/)
# ---------------------------------------------------------------------------
def cpu_frame_seconds(obj, psfs, n_iter, workers, timed_iterations):
    """Time the oracle port of the reference Deconvolver on a bounded sample:
    one forward model (+Poisson), H_t_normalization, `timed_iterations` RL
    iterations; frame(n_iter) = forward + normalisation + n_iter * mean."""
    import scipy.fft
    from oracle import line_sted_oracle as orc
    N = obj.shape[-1]
    with scipy.fft.set_workers(workers):
        d = orc.Deconvolver(psfs, engine='scipy')
        t0 = time.perf_counter()
        d.create_data_from_object(obj, total_brightness=total_brightness(N), random_seed=0)
        t_forward = time.perf_counter() - t0
        t0 = time.perf_counter()
        d.H_t(d.noisy_measurement)          # builds H_t_normalization (ref:521-522)
        t_norm = (time.perf_counter() - t0) / 2.0  # two H_t passes inside
        d.estimate = np.ones_like(d.noisy_measurement[0])
        d.num_iterations = 1
        its = []
        for _ in range(timed_iterations):
            t0 = time.perf_counter()
            d.iterate()
            its.append(time.perf_counter() - t0)
    t_iter = float(np.mean(its))
    return {'forward_s': t_forward, 'normalization_s': t_norm, 'iteration_s': t_iter,
            'frame_s': t_forward + t_norm + n_iter * t_iter}


def run_reference_arm(args, psfs, obj):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import scipy.fft  # noqa: F401
    from oracle import line_sted_oracle as orc
    import scipy.fft as sfft
    workers = os.cpu_count() or 1
    N, K = args.size, args.orientations
    with sfft.set_workers(workers):
        d = orc.Deconvolver(psfs, engine='scipy')
        t0 = time.perf_counter()
        d.create_data_from_object(obj, total_brightness=total_brightness(N), random_seed=0)
        t_forward = time.perf_counter() - t0
        t0 = time.perf_counter()
        d.H_t_normalization = d.H_t([np.ones(obj.shape)] * K, normalize=False)
        t_norm = time.perf_counter() - t0
        d.estimate = np.ones(obj.shape)
        d.num_iterations = 1
        for _ in range(args.warmup):
            d.iterate()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            d.iterate()
        t_iter = (time.perf_counter() - t0) / max(1, args.steps)
    frame_s = t_forward + t_norm + args.iterations * t_iter
    value = 1.0 / frame_s
    sample = ('each step = 1 RL iteration (2K fftconvolve) at full size; forward+Poisson '
              '(%.2f s) and H_t_normalization (%.2f s) timed once; frame(%d) = forward + '
              'norm + %d x mean iteration (%.3f s)'
              % (t_forward, t_norm, args.iterations, args.iterations, t_iter))
    print(json.dumps({
        'impl': 'reference', 'metric': 'frames_per_sec', 'value': value, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': frame_s * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': workers, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}))


def workload_config(args):
    return {'workload': 'config 4: synthetic %d^2 object (astronaut tiled), %d line orientations '
                        '(fig-2 2p0x_lr rescan PSF 107^2), Poisson noise + %d RL iterations '
                        '= frame(%d)' % (args.size, args.orientations, args.iterations,
                                         args.iterations),
            'size': args.size, 'orientations': args.orientations,
            'rl_iterations': args.iterations, 'psf_side': 107,
            'sharding': ('independent frames per GPU (no collective)' if args.shard == 'frames'
                         else 'orientations of each frame split over the GPUs, one NCCL all-reduce '
                              'of the Fourier-domain partial H_t sum per RL iteration'),
            'cache': 'working set (K measurements + K spectra + K OTFs, >1 GB) exceeds the '
                     '126 MB L2; no explicit flush'}


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    N, K, n_iter = args.size, args.orientations, args.iterations

    if args.impl == 'reference':
        if rank != 0:
            return
        from oracle import line_sted_oracle as orc
        psfs = orc.benchmark_psfs(K)
        run_reference_arm(args, psfs, synthetic_object(N))
        return

    os.environ['LSTED_DEVICE'] = str(local_rank)
    os.environ['LSTED_PRECISION'] = args.precision
    from rescan_line_sted_b200 import _lib, line_sted_tools as st
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    # PSFs through the product's own PSF path (GPU): psf_report + on-device orientation step.
    from rescan_line_sted_b200 import orientations
    base = st.psf_report('line', verbose=False, **FIG2_2P0X_LR)['psfs']['rescan_sted']
    psfs = orientations.line_orientation_psfs(base, K, EMISSION_2P0X_LR)
    obj_host = synthetic_object(N)
    brightness = total_brightness(N)

    lib = _lib.get()
    by_orientation = args.shard == 'orientations' and world > 1
    if by_orientation:
        from rescan_line_sted_b200 import sharded
        sd = sharded.OrientationShardedDeconvolver(
            st._stack_psfs(psfs), (N, N), precision=32 if args.precision == 'fp32' else 64,
            device=local_rank)
        h = sd.handle
    else:
        h = _lib.DeconvHandle(lib, st._stack_psfs(psfs), (N, N),
                              precision=32 if args.precision == 'fp32' else 64, device=local_rank)
    jobs = 1 if by_orientation else world   # frames in flight across the node
    info = h.info()
    pinned_in = _lib.pinned_empty((1, N, N))
    pinned_in[...] = obj_host
    pinned_out = _lib.pinned_empty((1, N, N))

    def frame_resident(seed):
        h.set_option('forget_normalization', 1)
        h.simulate(brightness, seed)
        h.iterate(n_iter)

    def frame_e2e(seed):
        h.upload_object(pinned_in)                       # H2D, pinned
        frame_resident(seed)
        h.get_into(_lib.ESTIMATE, 0, pinned_out)         # D2H (syncs)

    def barrier():
        h.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (headline `value`) ----
    h.upload_object(pinned_in)
    sampler = ClockSampler(local_rank)
    sampler.start()     # before the warm-up: nvidia-smi start-up stalls the driver briefly
    for w in range(args.warmup):
        frame_resident(1000 + w)
    barrier()
    t_begin = time.perf_counter()
    h.timer_start()
    for s in range(args.steps):
        frame_resident(s + 1 + (0 if by_orientation else 7919 * rank))
    ms_total = h.timer_stop()
    sampler.window(t_begin, time.perf_counter())
    barrier()
    # Same K steps again with a CUDA-event pair around every launch (costs a few
    # microseconds of stream time per launch, so it is kept out of `value`): the
    # per-kernel durations behind `roofline` and `kernels`.
    h.set_option('profile', 1)
    h.profile(reset=True)
    for s in range(args.steps):
        frame_resident(s + 1 + (0 if by_orientation else 7919 * rank))
    prof = h.profile(reset=True)
    h.set_option('profile', 0)
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = jobs * 1000.0 / ms_step

    # ---- end-to-end timing through host buffers ----
    for w in range(min(2, args.warmup)):
        frame_e2e(2000 + w)
    barrier()
    t0 = time.perf_counter()
    h.timer_start()
    for s in range(args.steps):
        frame_e2e(s + 1)
    h.sync()
    wall = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(h.timer_stop(), wall)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e) / args.steps
    est = np.array(pinned_out)
    if not np.isfinite(est).all() or est.min() < 0:
        raise RuntimeError('bench produced a non-finite estimate')

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timings) ----
    peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks_file):
        with open(peaks_file) as f:
            peak, peak_src = float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    elem = 4 if args.precision == 'fp32' else 8
    A = elem * N * N
    K_total = K
    if by_orientation:
        K = sd.k1 - sd.k0       # per-GPU accounting: this rank's orientations
    # algorithmic bytes charged to each launch (DESIGN.md "Roofline accounting")
    alg = {'row_mid': K * A,        # streams the K measurements once
           'row_final': 4 * A,      # norm + estimate read, estimate write (+ estimate for H)
           'col_h': 0.0, 'col_ht': 0.0,   # spectra only: no algorithmic array touched
           'row_inv_sim': 2 * K * A, 'row_fwd': A, 'row_inv_store': A}
    total_kernel_ms = sum(v[0] for v in prof.values())
    launches = int(sum(v[1] for v in prof.values()))
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms, dom_n = prof[dom]
    kernels = {k: {'ms_per_step': v[0] / args.steps, 'launches_per_step': v[1] / args.steps,
                   'share': v[0] / total_kernel_ms if total_kernel_ms else 0.0}
               for k, v in prof.items() if v[1]}
    # The iteration is one unit of (K+4)A bytes spread over its four launches:
    it_ms = sum(prof[k][0] for k in ('col_h', 'row_mid', 'col_ht', 'row_final')) / (
        args.steps * n_iter)
    step_bytes = (2 * K + 2 + n_iter * (K + 4)) * A
    roofline = {
        'bound': 'hbm', 'unit': 'GB/s', 'peak': peak, 'peak_source': peak_src,
        'kernel': dom, 'kernel_share_of_step': dom_ms / total_kernel_ms,
        'kernel_avg_ms': dom_ms / max(1, dom_n),
        'iteration_avg_ms': it_ms,
        'achieved': (K + 4) * A / (it_ms * 1e-3) / 1e9,
        'traffic': None,
        'note': 'achieved = (K+4)*A algorithmic bytes of one RL iteration / the summed average '
                'durations of its 4 launches (col_h,row_mid,col_ht,row_final); the path is '
                'FFT (FP32/SMEM) bound, not HBM bound: see DESIGN.md',
    }
    # DRAM traffic of the same four launches from the committed ncu --set full capture
    traffic_file = os.path.join(ROOT, 'profiles', 'r01_dram_traffic.json')
    if (os.path.isfile(traffic_file) and args.precision == 'fp32' and N == 2048 and K_total == 16
            and not by_orientation):
        with open(traffic_file) as f:
            t = json.load(f)
        roofline['traffic'] = sum(t[k]['dram_bytes_read'] + t[k]['dram_bytes_write']
                                  for k in ('col_h', 'row_mid', 'col_ht', 'row_final'))
        roofline['traffic_source'] = t['_source']
        roofline['algorithmic_bytes'] = (K + 4) * A
    roofline['frac'] = roofline['achieved'] / peak
    roofline['step_achieved'] = step_bytes / (ms_step * 1e-3) / 1e9
    roofline['step_frac'] = roofline['step_achieved'] / peak

    line = {
        'metric': 'frames_per_sec', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong' if by_orientation else 'weak',
        'vs_baseline': None,
        'dtype': 'f32' if args.precision == 'fp32' else 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'e2e': {'value': jobs * 1000.0 / ms_e2e, 'unit': 'frames/s',
                'h2d_bytes_per_step': int(pinned_in.nbytes),
                'd2h_bytes_per_step': int(pinned_out.nbytes), 'ms_per_step': ms_e2e},
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'kernels': kernels,
        'rl_iterations_per_sec': jobs * n_iter * 1000.0 / ms_step,
        'geometry': {'Ly': info.Ly, 'Lx': info.Lx, 'cols_per_cta': info.cols_per_cta,
                     'row_pairs_per_cta': info.row_pairs_per_cta,
                     'device_bytes': int(info.device_bytes)},
    }
    if world == 1 and not args.no_cpu_baseline:
        sample_iters = 2
        cpu = cpu_frame_seconds(obj_host, psfs, n_iter, os.cpu_count() or 1, sample_iters)
        line['cpu_baseline'] = {
            'value': 1.0 / cpu['frame_s'], 'unit': 'frames/s', 'cores': os.cpu_count() or 1,
            'kind': 'port',
            'sample': '1 forward+Poisson (%.2f s) + H_t_normalization (%.2f s) + %d RL iterations '
                      '(%.2f s each) of the same 2048^2/K=16 workload with scipy.fft workers = '
                      'all cores; frame(%d) extrapolated as forward + norm + %d x iteration'
                      % (cpu['forward_s'], cpu['normalization_s'], sample_iters,
                         cpu['iteration_s'], n_iter, n_iter)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
