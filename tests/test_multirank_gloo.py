"""N > 1 host logic on CPU: world_size-2 gloo.  Shard ranges tile the global id
space; per-rank products all-reduce to the single-rank product (the packet-level
work is covered by the oracle: Philox ids are global)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import REPO
from nexoclom_b200.sharding import shard_range


def test_shard_ranges_tile_the_id_space():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (100_000_000, 8)):
        got = [shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and sum(c for _, c in got) == n
        for (f0, c0), (f1, _) in zip(got, got[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in got) - min(c for _, c in got) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from common import workload
    from nexoclom_b200.runsetup import RunSetup
    from nexoclom_b200.sharding import shard_range, allreduce_products
    from oracle import initial_state, imaging
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    n_total = 6000
    first, n = shard_range(n_total, rank, world)
    x0 = initial_state.draw_x0(setup, n, 3, first_id=first)
    img, cnt, _, _ = imaging.create_image(x0[:, 1] * 2, x0[:, 2] * 2, x0[:, 3] * 2, x0[:, 5], x0[:, 7],
                                          vrplanet=0.0, M=np.eye(3), dims=[64, 64],
                                          xrange=(-4, 4), zrange=(-4, 4), apix=1.0,
                                          quantity='column')
    t_img, t_cnt = torch.from_numpy(img), torch.from_numpy(cnt).to(torch.int64)
    allreduce_products(t_img, t_cnt)
    if rank == 0:
        q.put((t_img.numpy(), t_cnt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank():
    from common import workload
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state, imaging
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    img2, cnt2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    x0 = initial_state.draw_x0(setup, 6000, 3)
    img, cnt, _, _ = imaging.create_image(x0[:, 1] * 2, x0[:, 2] * 2, x0[:, 3] * 2, x0[:, 5], x0[:, 7],
                                          vrplanet=0.0, M=np.eye(3), dims=[64, 64],
                                          xrange=(-4, 4), zrange=(-4, 4), apix=1.0,
                                          quantity='column')
    assert np.array_equal(cnt2, cnt.astype(np.int64))
    assert np.allclose(img2, img, rtol=1e-13, atol=0)
