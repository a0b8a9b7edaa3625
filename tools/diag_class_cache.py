#!/usr/bin/env python
"""A/B of the streaming kernel's per-packet class cache (option `class_cache`): end-to-end
host-buffer time and the resident streaming kernel, 1e7 packets."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np, torch
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0); n = 10_000_000
setup = RunSetup(workload('Na.maxwellian.radpres.input')); setup.upload(eng)
sp = setup.source_params(eng)
eng.init_state(sp, 0, 0, n)
att0, _ = eng.integrate_adaptive(); ref = eng.export_state()
eng.init_state(sp, 0, 0, n)
X0 = eng.export_x0()[:8]
host = torch.empty((8, n), dtype=torch.float64).pin_memory(); host.numpy()[:] = X0
cols = [host.numpy()[k] for k in range(8)]
for rep in range(2):
  for cc in (0, 1):
    eng.set_option('class_cache', cc)
    best = 1e9
    for r in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        att, _ = eng.integrate_adaptive_host(cols, nchunks=16); eng.sync()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    same = bool(np.array_equal(eng.export_state(), ref))
    print(f'class_cache={cc}: {best:.2f} ms att_ok={att == att0} identical={same}', flush=True)
    eng.set_option('schedule', 2)
    eng.init_state(sp, 0, 0, n); eng.integrate_adaptive(); eng.init_state(sp, 0, 0, n); eng.integrate_adaptive()
    print(f'   resident streaming kernel: {eng.last_kernel_ms():.2f} ms', flush=True)
    eng.set_option('schedule', 1)
