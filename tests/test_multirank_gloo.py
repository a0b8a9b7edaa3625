"""N > 1 host logic on CPU: world_size-2 gloo.  Shard ranges tile the global id space;
per-rank products all-reduce (the product's own ``sharding.allreduce_sum``) to the single-rank
product (the packet-level work is covered by the oracle: Philox ids are global); ``Input.run``
under a process group hands every rank its own id range, one seed, and stops when the ranks
TOGETHER have run the requested packets.  The GPU side of the same path:
tests/test_multigpu.py."""
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
    from nexoclom_b200.sharding import shard_range, allreduce_sum
    from oracle import initial_state, imaging
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    n_total = 6000
    first, n = shard_range(n_total, rank, world)
    x0 = initial_state.draw_x0(setup, n, 3, first_id=first)
    img, cnt, _, _ = imaging.create_image(x0[:, 1] * 2, x0[:, 2] * 2, x0[:, 3] * 2, x0[:, 5], x0[:, 7],
                                          vrplanet=0.0, M=np.eye(3), dims=[64, 64],
                                          xrange=(-4, 4), zrange=(-4, 4), apix=1.0,
                                          quantity='column')
    img = np.ascontiguousarray(img, dtype=np.float64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int64)
    allreduce_sum(img, cnt)                       # what ModelImage does under gloo
    if rank == 0:
        q.put((img, cnt))
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


def _run_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    os.environ.pop('NEXOCLOM_B200_SAVEPATH', None)
    from common import workload
    from nexoclom_b200 import sharding, catalogue
    import nexoclom_b200                                   # noqa: F401
    output_mod = sys.modules['nexoclom_b200.Output']         # the module (the package attribute is the class)
    assert sharding.init('gloo') == (rank, world)
    calls = []

    class FakeOutput:
        """Records what Input.run asks for and registers itself like a real Output."""

        def __init__(self, inputs, npackets, compress=True, seed=None, first_id=0, **kw):
            self.inputs, self.npackets, self.totalsource = inputs, int(npackets), float(npackets)
            calls.append((int(first_id), int(npackets), seed))
            catalogue.register(inputs, self)
    output_mod.Output = FakeOutput
    inputs = workload('Ca.isotropic.flat.input')
    inputs.run(10_001, packs_per_it=3000)                       # seed=None: one seed for all
    first_pass = list(calls)
    inputs.run(12_000, packs_per_it=3000)                       # 1 999 more, ids continue
    _, files, mine, _ = inputs.search()
    q.put((rank, first_pass, calls[len(first_pass):], mine))
    dist.barrier()
    dist.destroy_process_group()


def test_input_run_shards_packet_ids_over_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, a1, a2, mine0), (_, b1, b2, mine1) = got
    # first run: 10 001 packets -> 5 001 + 5 000, chunks of <= 3 000, contiguous global ids
    ranges = sorted((f, f + n) for f, n, _ in a1 + b1)
    assert ranges[0][0] == 0 and ranges[-1][1] == 10_001
    assert all(e0 == s1 for (_, e0), (s1, _) in zip(ranges, ranges[1:]))
    assert max(n for _, n, _ in a1 + b1) <= 3000
    assert sum(n for _, n, _ in a1) == 5001 and sum(n for _, n, _ in b1) == 5000
    seeds = {s for _, _, s in a1 + b1}
    assert len(seeds) == 1 and None not in seeds                # one broadcast seed
    # second run tops up to 12 000: the ids continue after the 10 001 already there
    ranges2 = sorted((f, f + n) for f, n, _ in a2 + b2)
    assert ranges2[0][0] == 10_001 and ranges2[-1][1] == 12_000
    assert mine0 + mine1 == 12_000
