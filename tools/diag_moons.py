#!/usr/bin/env python
"""BASELINE configs[3]: Na from Io at Jupiter (gravity of Jupiter + Io, radiation pressure,
photo-loss, adaptive step) -- K1 + K2 (generic kernel) timing."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
setup = RunSetup(workload('Na.Io.Jupiter.input'))
setup.upload(eng)
sp = setup.source_params(eng)
for rep in range(2):
    eng.init_state(sp, 0, 0, n)
    k1 = eng.last_kernel_ms()
    att, acc = eng.integrate_adaptive()
    ms = eng.last_kernel_ms()
x = eng.export_state()
r = np.sqrt((x[1:4]**2).sum(0))
print(f'n={n} K1 {k1:.2f} ms, K2 {ms:.2f} ms, attempted={att} ({att / n:.1f} per packet) '
      f'steps/s={att / ms * 1e3:.4g}; alive {np.mean(x[7] > 0):.3f}, r median {np.median(r):.2f} R_J')
