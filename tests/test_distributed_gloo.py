"""N > 1 path on CPU: world_size 2 over gloo.  The orientation-sharded engine
(same orchestration as the NCCL build, collectives routed to gloo) must
reproduce the reference's golden vectors, and sweep sharding must return the
reports in order."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import emul_support
from oracle import line_sted_oracle as orc
from rescan_line_sted_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_orientation_slices_partition_everything():
    for K in (1, 4, 10, 16, 32):
        for world in (1, 2, 3, 4, 8):
            if world > K:
                with pytest.raises(ValueError):
                    sharded.orientation_slice(K, 0, world)
                continue
            got = []
            for r in range(world):
                a, b = sharded.orientation_slice(K, r, world)
                assert b > a
                got += list(range(a, b))
            assert got == list(range(K))
    assert sorted(sum((sharded.shard_items(7, r, 3) for r in range(3)), [])) == list(range(7))


def test_world_size_2_gloo(tmp_path):
    emul_support.build_emulator()
    out = str(tmp_path / 'result.json')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', '29517',
           os.path.join(ROOT, 'tests', '_gloo_worker.py'), out]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    with open(out) as f:
        r = json.load(f)
    assert r['fp64']['noiseless'] < 1e-12 and r['fp64']['norm'] < 1e-12
    assert r['fp64']['est1'] < 1e-12 and r['fp64']['est8'] < 1e-11
    assert r['fp32']['noiseless'] < 1e-5 and r['fp32']['est1'] < 1e-5 and r['fp32']['est8'] < 1e-4
    assert r['fp64']['replica_diff'] == 0.0 and r['fp32']['replica_diff'] == 0.0
    # 3 x 4 tiles of 56 x 54 pixels over 2 ranks: a 1 x 2 rank grid, rank 0 owns columns 0..107
    assert r['tiles']['rows'] == [0, 150] and r['tiles']['cols'] == [0, 108] and r['tiles']['noisy_same']
    assert r['tiles']['est'] < 1e-11
    exc = [0.1, 0.5, 1, 2, 4, 8]
    dep = [1, 3, 9, 27, 54, 81]
    want = [orc.psf_report('line', e, d, 8, 1, use_closed_form=True)['expected_emission']
            for e, d in zip(exc, dep)]
    assert np.allclose(r['sweep_emission'], want, rtol=1e-12)


def test_tile_grid_world_size_4_gloo(tmp_path):
    """Tiles of a tiled object dealt to a 2 x 2 grid of ranks: halo exchange with edge and
    corner neighbours, odd and even PSF sizes, fp64 and fp32."""
    emul_support.build_emulator()
    out = str(tmp_path / 'tiles.json')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=4',
           '--master-addr', '127.0.0.1', '--master-port', '29519',
           os.path.join(ROOT, 'tests', '_gloo_tiles_worker.py'), out]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    with open(out) as f:
        r = json.load(f)
    for name, tol in (('odd_fp64', 1e-11), ('even_fp64', 1e-11), ('odd_fp32', 1e-4)):
        c = r[name]
        assert c['noisy_same'] and c['outside_zero'], name
        assert c['est'] < tol and c['est_injected'] < tol, (name, c)
    # 4 x 4 tiles: a 2 x 2 grid of ranks; 4 x 5 tiles (the even PSF): four row bands of one
    # tile row each; either way the rectangles tile the image
    for name in ('odd_fp64', 'odd_fp32'):
        rects = r[name]['all_rects']
        assert len({(a, b) for a, b, _, _ in rects}) == 2 and len({(x, y) for _, _, x, y in rects}) == 2
    assert r['odd_fp64']['tiles'] == [4, 4] and r['even_fp64']['tiles'] == [4, 5]
    assert sum((b - a) * (d - c) for a, b, c, d in r['odd_fp64']['all_rects']) == 200 * 200
    assert sum((b - a) * (d - c) for a, b, c, d in r['even_fp64']['all_rects']) == 190 * 215
