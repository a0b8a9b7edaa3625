#!/usr/bin/env python
"""K2 (sort + integrate) on a few shards; the sort kernels' share comes from the ncu launch list."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
n = 10_000_000
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    best = 1e9
    for rep in range(4):
        eng.init_state(sp, 0, r * n, n)
        att, acc = eng.integrate_adaptive()
        best = min(best, eng.last_kernel_ms())
    print(f'shard {r}: sort + K2 {best:.3f} ms ({att} steps)', flush=True)
