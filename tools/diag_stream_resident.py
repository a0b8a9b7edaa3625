#!/usr/bin/env python
"""The streaming (class-ordered) K2 kernel over a RESIDENT X0 slab (option schedule = 2) against
the sorted kernel and the host-buffer call: how much of the end-to-end kernel time is the
schedule itself and how much the late arrival of long packets."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import torch
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
n = 10_000_000
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    res = {}
    for name, sched in (('sorted', 0), ('stream resident', 2)):
        eng.set_option('schedule', sched)
        best = 1e9
        for rep in range(3):
            eng.init_state(sp, 0, r * n, n)
            att, _ = eng.integrate_adaptive()
            best = min(best, eng.last_kernel_ms())
        res[name] = best
    eng.set_option('schedule', 1)
    eng.init_state(sp, 0, r * n, n)
    host.numpy()[:] = eng.export_x0()[:8]
    cols = [host.numpy()[k] for k in range(8)]
    best = 1e9
    for rep in range(3):
        eng.integrate_adaptive_host(cols, nchunks=32)
        best = min(best, eng.last_kernel_ms())
    res['host buffers'] = best
    print(f'shard {r}: ' + ', '.join(f'{k} {v:.2f} ms' for k, v in res.items()), flush=True)
