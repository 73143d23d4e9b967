"""Orientation sharding over NCCL on 2 GPUs (skipped on a 1-GPU box): same
golden vectors as the single-GPU tests, identical replicas on every rank, and
a noise field that does not depend on the number of GPUs."""
import ctypes
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    from rescan_line_sted_b200 import _lib
    n = ctypes.c_int(0)
    _lib.get().cdll.lsted_device_count(ctypes.byref(n))
    return n.value


def test_orientation_sharding_nccl(tmp_path):
    if device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'result.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', '29519',
           os.path.join(ROOT, 'tests', '_nccl_worker.py'), out]
    res = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR='127.0.0.1'),
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    with open(out) as f:
        r = json.load(f)
    assert r['fp64']['est1'] < 1e-12 and r['fp64']['est8'] < 1e-11
    assert r['fp32']['est1'] < 1e-5 and r['fp32']['est8'] < 1e-4
    assert r['tiles']['noisy_same'] and r['tiles']['est'] < 1e-11
    for tag in ('fp64', 'fp32'):
        assert r[tag]['replica_diff'] == 0.0
        assert r[tag]['noise_independent_of_world']
    # fast path: fused peer-memory reduction ('1') and NCCL all-reduce ('0') against one GPU
    # (fp32 also through the NVSwitch multicast object, 'nvls', where the box supports it)
    assert 'nvls' in r['p2p_fp32'] or r['p2p_fp32'].get('nvls_unavailable')
    for tag in ('p2p_fp32', 'p2p_fp64'):
        for mode in [m for m in ('nvls', '1', '0') if m in r[tag]]:
            assert r[tag][mode] < 10 * r[tag]['tol'], (tag, mode, r[tag])
            assert r[tag][mode + '_replica_diff'] == 0.0, (tag, mode, r[tag])
