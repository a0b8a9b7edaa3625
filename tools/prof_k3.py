#!/usr/bin/env python
"""One K1 + K3 pass for ncu captures.  usage: prof_k3.py [npackets] [fused:0|1]"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
from bench import image_params
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
fused = len(sys.argv) > 2 and sys.argv[2] == '1'
eng = Engine(0)
setup = RunSetup(workload('Na.bounce.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
eng.init_state(setup.source_params(eng), 0, 0, n)
if fused:
    ip, _ = image_params(setup)
    eng.image_begin(800, 800)
    a, b = eng.image_device_ptrs()
    _, nsteps, steps = eng.integrate_constant(seed=1, image_params=ip, image_dev=a, counts_dev=b, n=n)
else:
    _, nsteps, steps = eng.integrate_constant(seed=1, n=n)
eng.sync()
print(f'n={n} fused={fused} steps={steps} ms={eng.last_kernel_ms():.3f}')
