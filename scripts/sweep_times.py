"""Config 3: the 64-point line-STED sweep (8 excitation x 8 depletion
brightnesses) at steps 8 (n = 35) and 25 (n = 107), timed through the public API, plus
`tune_psf` / `tune_psf_batch`.  Under torchrun the points are dealt round-robin over the
ranks (`sharded.shard_items`), every rank runs its share as one fused launch, and the
reports are gathered once (`sharded.gather_reports`): no data-path collective.

    python scripts/sweep_times.py                 # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29517 scripts/sweep_times.py

`sweep_benchmark()` is also what bench.py records under extra.config3_sweep."""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

EXC = [0.05, 0.1, 0.25, 0.5, 1, 2, 4, 8]
DEP = [0, 1, 3, 9, 27, 54, 81, 108]


def _grid(repeat=1):
    E, D = np.meshgrid(EXC, DEP, indexing='ij')
    return np.tile(E.ravel(), repeat), np.tile(D.ravel(), repeat)


def sweep_benchmark(rank=0, world=1, dist=None, repeats=5, big=64):
    """points/s of the sweep on `world` GPUs: the 64-point grid of config 3 and the same grid
    repeated `big` times (4096 points, enough to fill the GPUs).  Wall clock around
    shard -> one fused launch per rank -> gather, max over ranks."""
    from rescan_line_sted_b200 import line_sted_tools as st
    from rescan_line_sted_b200 import sharded
    out = {}

    def barrier():
        if dist is not None:
            dist.barrier()

    for steps in (8, 25):
        for label, rep in (('grid64', 1), ('grid%d' % (64 * big), big)):
            exc, dep = _grid(rep)
            mine = sharded.shard_items(exc.size, rank, world)
            for psfs in (False, True):
                if psfs and rep > 1:
                    continue
                best = None
                for it in range(repeats + 1):
                    barrier()
                    t0 = time.perf_counter()
                    local = st.psf_report_batch('line', exc[mine], dep[mine], steps, 1, psfs=psfs) \
                        if mine else []
                    if dist is not None:
                        reports = sharded.gather_reports(local, exc.size)
                    else:
                        reports = local
                    barrier()
                    dt = time.perf_counter() - t0
                    if dist is not None:
                        import torch
                        t = torch.tensor([dt], dtype=torch.float64,
                                         device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        dt = float(t.item())
                    if it > 0:                      # first pass is the warm-up
                        best = dt if best is None else min(best, dt)
                assert len(reports) == exc.size and all(r is not None for r in reports)
                key = 'steps%d_%s_%s' % (steps, label, 'with_psfs' if psfs else 'scalars')
                out[key] = {'points': int(exc.size), 'seconds': best,
                            'points_per_s': exc.size / best}
    return out


def single_gpu_details():
    """Latency of one psf_report (fused launch vs the three-launch host-fit path), tune_psf
    and tune_psf_batch; the unmodified reference beside them where baseline/_ref exists."""
    from rescan_line_sted_b200 import line_sted_tools as st
    out = {}
    exc, dep = _grid()
    for steps in (8, 25):
        st.psf_report('line', 1, 9, steps, 1, verbose=False)
        t0 = time.perf_counter()
        for e, d in zip(exc[:32], dep[:32]):
            st.psf_report('line', e, d, steps, 1, verbose=False)
        fused = (time.perf_counter() - t0) / 32
        os.environ['LSTED_HOST_FIT'] = '1'
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            t0 = time.perf_counter()
            for e, d in zip(exc[:32], dep[:32]):
                st.psf_report('line', e, d, steps, 1, verbose=False)
            host = (time.perf_counter() - t0) / 32
        os.environ.pop('LSTED_HOST_FIT')
        out['psf_report_line_steps%d_ms' % steps] = {'fused_device_fit': fused * 1e3,
                                                     'host_fit_path': host * 1e3}
    target = dict(psf_type='line', scan_type='rescanned', desired_resolution_improvement=4.07614,
                  desired_emissions_per_molecule=3.0227, max_excitation_brightness=0.25,
                  steps_per_improved_psf_width=4.)
    t0 = time.perf_counter()
    st.tune_psf(**target)
    out['tune_psf_line_rescanned_R4.08_s'] = time.perf_counter() - t0
    # the twelve line-STED operating points of figure 2 (line_sted_figure_2.py:77-162)
    fig2 = [(1.5, 2.8289), (1.5, 2.61804), (2.0, 3.0127), (2.0, 3.0227), (2.5, 3.7863),
            (2.5, 3.8018), (3.0, 5.0307), (3.0, 5.0371), (4.0, 7.3720), (4.0, 7.3720),
            (1.0, 4.0), (1.0, 4.0)]
    targets = [dict(psf_type='line', scan_type='rescanned', desired_resolution_improvement=r,
                    desired_emissions_per_molecule=em, max_excitation_brightness=0.25,
                    steps_per_improved_psf_width=3.) for r, em in fig2 if r > 1.0]
    t0 = time.perf_counter()
    seq = [st.tune_psf(**t) for t in targets]
    t_seq = time.perf_counter() - t0
    t0 = time.perf_counter()
    bat = st.tune_psf_batch(targets)
    t_bat = time.perf_counter() - t0
    same = all(a['depletion_brightness'] == b['depletion_brightness'] and
               a['excitation_brightness'] == b['excitation_brightness'] for a, b in zip(seq, bat))
    out['tune_psf_10_fig2_line_targets_s'] = {'one_by_one': t_seq, 'tune_psf_batch': t_bat,
                                              'identical_results': bool(same)}
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                           'baseline', '_ref')
    if os.path.isfile(os.path.join(ref_dir, 'line_sted_tools.py')):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(ref_dir)), 'tests'))
        from _reference_loader import load_reference
        ref = load_reference()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for steps, npts in ((8, 16), (25, 4)):
                t0 = time.perf_counter()
                for e, d in zip(exc[:npts], dep[:npts]):
                    ref.psf_report('line', e, d, steps, 1, verbose=False)
                out['reference_cpu_psf_report_line_steps%d_ms' % steps] = \
                    (time.perf_counter() - t0) / npts * 1e3
    return out


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', rank))
        os.environ['LSTED_DEVICE'] = str(local)
        torch.cuda.set_device(local)
        dist.init_process_group('nccl')
    out = {'n_gpus': world, 'sweep': sweep_benchmark(rank, world, dist)}
    if world == 1:
        out['details'] = single_gpu_details()
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
