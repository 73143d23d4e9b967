#!/usr/bin/env python3
"""BASELINE config 5: synthetic 8192^2 object, 32 orientations, fp64, processed in
overlap-save tiles (2160-point transforms) -- on one GPU, or sharded into row bands
over the ranks of a torchrun launch.  Prints one JSON line on rank 0.

  python scripts/config5.py [--size 8192] [--orientations 32] [--iterations 4]
  python -m torch.distributed.run --nproc-per-node 8 ... scripts/config5.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=8192)
    ap.add_argument('--orientations', type=int, default=32)
    ap.add_argument('--iterations', type=int, default=4)
    ap.add_argument('--precision', default='fp64')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    os.environ['LSTED_DEVICE'] = str(local)
    from rescan_line_sted_b200 import _lib, sharded, line_sted_tools as st
    N, K = args.size, args.orientations
    prec = 64 if args.precision == 'fp64' else 32
    base = st.psf_report('line', verbose=False, **bench.FIG2_2P0X_LR)['psfs']['rescan_sted']
    psfs = st._stack_psfs(bench.orientation_psfs(base, K))
    obj = bench.synthetic_object(N)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        d = sharded.TileShardedDeconvolver(psfs, (N, N), precision=prec, device=local)
        h = d.handle
    else:
        h = _lib.DeconvHandle(_lib.get(), psfs, (N, N), precision=prec, device=local,
                              tile_fft_len=2160)
    info = h.info()
    t0 = time.perf_counter()
    h.create_data(obj, bench.total_brightness(N), 1)
    h.sync()
    t_forward = time.perf_counter() - t0
    h.iterate(1)             # includes H_t_normalization
    h.sync()
    if world > 1:
        dist.barrier()
    h.timer_start()
    h.iterate(args.iterations)
    ms = h.timer_stop() / args.iterations
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    est = h.get(_lib.ESTIMATE)
    ok = bool(np.isfinite(est).all() and est.min() >= 0)
    if rank == 0:
        elem = prec // 8
        A = elem * N * N
        print(json.dumps({
            'config': 'config 5: %d^2 tiled object, %d orientations, %s' % (N, K, args.precision),
            'n_gpus': world, 'tiles': [info.tiles_y, info.tiles_x],
            'tile_fft': info.Lx, 'tile_out': info.tile_out_y,
            'forward_s': t_forward, 'ms_per_rl_iteration': ms,
            'rl_iterations_per_sec': 1000.0 / ms,
            'algorithmic_GBps': (K + 4) * A / (ms * 1e-3) / 1e9,
            'device_bytes_rank0': int(info.device_bytes), 'estimate_finite_nonneg': ok,
            'estimate_sum_over_brightness': float(est.sum() / bench.total_brightness(N))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
