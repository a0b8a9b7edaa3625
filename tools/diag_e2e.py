#!/usr/bin/env python
"""End-to-end (host buffers) timing of the adaptive path: H2D alone, K2 alone,
nx_integrate_adaptive_host for several chunk counts."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup

eng = Engine(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
chunks = [int(c) for c in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3, 4, 6, 8, 16]
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
att0 = None
scheds = [int(c) for c in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1]
for sched in scheds:
    eng.set_option('schedule', sched)
    for rep in range(3):
        eng.init_state(sp, 0, 0, n)
        att, _ = eng.integrate_adaptive()
    print(f'schedule={sched} K2 resident: {eng.last_kernel_ms():.2f} ms, {att} steps', flush=True)
    assert att0 is None or att == att0
    att0 = att
eng.init_state(sp, 0, 0, n)
X0 = eng.export_x0()[:8]
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
host.numpy()[:] = X0
cols = [host.numpy()[k] for k in range(8)]
dev = torch.empty((8, n), dtype=torch.float64, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print(f'H2D {host.numel() * 8 / 1e6:.0f} MB: {(t1 - t0) * 1e3:.2f} ms = {host.numel() * 8 / (t1 - t0) / 1e9:.1f} GB/s', flush=True)
ref = None
for sched in scheds:
    eng.set_option('schedule', sched)
    for c in chunks:
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            att, _ = eng.integrate_adaptive_host(cols, nchunks=c)
            eng.sync()
            t1 = time.perf_counter()
            best = min(best, (t1 - t0) * 1e3)
        assert att == att0, (att, att0)
        X = eng.export_state()
        if ref is None:
            ref = X
        same = bool(np.array_equal(X, ref))
        print(f'schedule={sched} nchunks={c}: wall {best:.2f} ms  (device-timed {eng.last_kernel_ms():.2f} ms)  '
              f'{att / best * 1e3:.4g} steps/s  identical_state={same}', flush=True)
