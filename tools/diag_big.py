import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
for n in (10_000_000, 20_000_000, 40_000_000):
    best = 1e9
    for rep in range(2):
        eng.init_state(sp, 0, 0, n)
        att, acc = eng.integrate_adaptive()
        best = min(best, eng.last_kernel_ms())
    print(f'n={n} ms={best:.3f} steps/s={att / best * 1e3:.4g}', flush=True)
