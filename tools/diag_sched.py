#!/usr/bin/env python
"""K2 class-ordered schedule: timing + (debug builds) phase timeline."""
import os, sys, ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup

eng = Engine(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
ncls = int(sys.argv[2]) if len(sys.argv) > 2 else 5
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
import torch
eng.init_state(sp, 0, 0, n)
X0 = eng.export_x0()[:8]
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
host.numpy()[:] = X0
cols = [host.numpy()[k] for k in range(8)]
for nchunks, model in [tuple(int(c) for c in a.split(':')) for a in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['1:1', '32:1'])]:
    eng.set_option('order_packets', model)
    best = 1e9
    for rep in range(3):
        att, _ = eng.integrate_adaptive_host(cols, nchunks=nchunks)
        best = min(best, eng.last_kernel_ms())
    q = (C.c_ulonglong * 512)()
    eng.lib.nx_debug_queue(eng.ctx, q, 512)
    d = list(q)
    line = f'nchunks={nchunks} model={model} host path: {best:.2f} ms, {att} steps'
    if d[0]:
        t0 = d[0]
        line += ' | class exhausted at ms: ' + ' '.join(f'{(d[1 + c] - t0) / 1e6:.2f}' for c in range(ncls) if d[1 + c])
        line += f' | end {(d[16] - t0) / 1e6:.2f} | warp-iters {d[17]} lane-util {att / (32 * max(d[17], 1)):.3f} scans {d[18]}'
        h = np.array(d[32:32 + 200], dtype=float)
        nb = int(np.nonzero(h)[0].max()) + 1 if h.any() else 0
        line += '\n   lane-steps per 1 ms (1e6): ' + ' '.join(f'{v / 1e6:.0f}' for v in np.add.reduceat(h[:nb], np.arange(0, nb, 4)))
    print(line, flush=True)
