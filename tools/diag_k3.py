#!/usr/bin/env python
"""K3 (constant step + bounce, configs[2] physics) timings: plain, fused image, row sink."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
sys.path.insert(0, os.path.dirname(HERE))
from bench import image_params

eng = Engine(0)
name = sys.argv[2] if len(sys.argv) > 2 else 'Na.bounce.input'
setup = RunSetup(workload(name))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
sp = setup.source_params(eng)
ip, _ = image_params(setup)
for n in [int(float(x)) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['2e6', '8e6'])]:
    for fused in (False, True):
        best = 1e30
        for rep in range(3):
            eng.init_state(sp, 0, 0, n)
            if fused:
                eng.image_begin(800, 800)
                a, b = eng.image_device_ptrs()
                _, nsteps, steps = eng.integrate_constant(seed=1, image_params=ip, image_dev=a, counts_dev=b, n=n)
            else:
                _, nsteps, steps = eng.integrate_constant(seed=1, n=n)
            eng.sync()
            best = min(best, eng.last_kernel_ms())
        print(f'K3 {name} n={n} fused={fused}: {best:.3f} ms  {steps / best / 1e6:.3f} e9 packet-steps/s '
              f'({steps} steps, {steps * 626 / best / 1e9 / 37.225:.3f} of 37.2 TF)', flush=True)
