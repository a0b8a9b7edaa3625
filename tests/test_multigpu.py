"""Sharded product path on real GPUs: N ranks under torchrun, each
integrating its slice of the global packet ids, ModelImage / LOSResult combined with one NCCL
all-reduce per product == the single-process products.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import REPO

WORKER = os.path.join(REPO, 'tests', '_multigpu', 'worker.py')


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, name, npackets, seed, dest, packs_per_it=None):
    env = dict(os.environ)
    env.pop('NEXOCLOM_B200_SAVEPATH', None)
    args = [name, str(npackets), str(seed), dest] + ([str(packs_per_it)] if packs_per_it else [])
    if world == 1:
        cmd = [sys.executable, WORKER] + args
    else:
        port = 29600 + os.getpid() % 300
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               f'--nproc-per-node={world}', '--master-addr', '127.0.0.1', '--master-port',
               str(port), WORKER] + args
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-4000:]
    return np.load(dest)


@pytest.mark.gpu
@pytest.mark.parametrize('name, npackets, packs_per_it', [
    ('Na.maxwellian.radpres.input', 400_000, None),       # adaptive, configs[1] physics
    ('Na.bounce.input', 6_000, 1_000),                    # constant step + bounce, two chunks
])
def test_sharded_products_equal_single_process(tmp_path, name, npackets, packs_per_it):
    world = min(_gpus(), 8)
    if world < 2:
        pytest.skip('needs at least 2 GPUs')
    one = _run(1, name, npackets, 4, str(tmp_path / 'one.npz'), packs_per_it)
    many = _run(world, name, npackets, 4, str(tmp_path / 'many.npz'), packs_per_it)
    assert int(many['world']) == world and int(one['world']) == 1
    assert int(many['mine']) < npackets                    # rank 0 ran only its share
    assert float(many['totalsource']) == float(one['totalsource'])
    assert float(many['atoms_per_packet']) == float(one['atoms_per_packet'])
    assert np.array_equal(many['packet_image'], one['packet_image'])       # counts: exact
    assert one['packet_image'].sum() > 1000
    for key in ('image', 'column'):
        a, b = many[key], one[key]
        nz = b > 0
        assert np.array_equal(nz, a > 0) and nz.sum() > 50
        assert np.max(np.abs(a[nz] - b[nz]) / b[nz]) < 1e-12              # f64 sums, other order
    assert np.array_equal(many['npackets_los'], one['npackets_los'])
    assert one['npackets_los'].sum() > 100
    r1, rn = one['radiance'], many['radiance']
    nz = r1 > 0
    assert np.array_equal(nz, rn > 0)
    assert np.max(np.abs(rn[nz] - r1[nz]) / r1[nz]) < 1e-11
    assert float(many['sourcerate']) == pytest.approx(float(one['sourcerate']), rel=1e-11)
