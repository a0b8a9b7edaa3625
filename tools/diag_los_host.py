#!/usr/bin/env python
"""Where the wall time of eng.los_accumulate goes (bench geometry: 1e5 lines x 1e7 packets)."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
from nexoclom_b200._lib import LosParams
import bench
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
eng.init_state(setup.source_params(eng), 0, 0, n)
eng.integrate_adaptive(n)
los, dplan = bench.synthetic_los(100_000)
lp = LosParams()
lp.dphi, lp.outeredge = float(np.radians(1.0)), 25.0
lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
lp.quantity, lp.round_f32, lp.skip_dead = 1, 1, 0
for it in range(4):
    eng.sync()
    t0 = time.perf_counter()
    rad, npk, inc = eng.los_accumulate(los, dplan, lp, n=n)
    print(f'wall {(time.perf_counter() - t0) * 1e3:.2f} ms, kernels {eng.last_kernel_ms():.2f} ms, hits {int(npk.sum())}', flush=True)
