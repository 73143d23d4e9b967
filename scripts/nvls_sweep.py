#!/usr/bin/env python3
"""Launch shape of the NVLS all-reduce kernel (orientation sharding, fp32 2048^2, K = 16): CTAs x
threads against the kernel's CUDA-event time.  Run under torchrun on N GPUs:
  python -m torch.distributed.run --nproc-per-node 8 ... scripts/nvls_sweep.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    os.environ['LSTED_DEVICE'] = str(local)
    os.environ['LSTED_NVLS'] = '1'
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from rescan_line_sted_b200 import _lib, sharded, orientations, line_sted_tools as st
    N, K = 2048, 16
    base = st.psf_report('line', verbose=False, **bench.FIG2_2P0X_LR)['psfs']['rescan_sted']
    psfs = st._stack_psfs(orientations.line_orientation_psfs(base, K, bench.EMISSION_2P0X_LR))
    obj = bench.synthetic_object(N)
    out = {}
    for ctas, threads, probe in ((148, 512, 0), (148, 512, 1), (32, 512, 0), (32, 512, 1), (16, 256, 0), (8, 512, 0)):
        d = sharded.OrientationShardedDeconvolver(psfs, (N, N), precision=32, device=local)
        assert d.nvls
        h = d.handle
        h.set_option('nvls_ctas', ctas)
        h.set_option('nvls_threads', threads)
        h.set_option('nvls_probe', probe)
        d.create_data(obj, bench.total_brightness(N), 1)
        h.iterate(8)
        h.sync()
        dist.barrier()
        h.set_option('profile', 1)
        h.profile(reset=True)
        h.timer_start()
        h.iterate(64)
        ms = h.timer_stop() / 64
        prof = h.profile(reset=True)
        h.set_option('profile', 0)
        out['%dx%d%s' % (ctas, threads, ' barriers only' if probe else '')] = {'iteration_ms': ms, 'allreduce_ms': prof['elementwise'][0] / max(1, prof['elementwise'][1]),
                                          'col_ht_ms': prof['col_ht'][0] / max(1, prof['col_ht'][1])}
        d.close()
        dist.barrier()
    if dist.get_rank() == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
