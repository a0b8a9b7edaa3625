#!/usr/bin/env python
"""K6 timing: source map of n initial states on the reference's default 180 x 90 grid,
smear radius 10 deg; CPU oracle (NumPy + sklearn BallTree, the reference's algorithm) on a
sample for comparison."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import get_engine
from nexoclom_b200.make_source_map import source_map_arrays
from nexoclom_b200.runsetup import RunSetup
from oracle import source_map
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
ncpu = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000
eng = get_engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
eng.init_state(setup.source_params(eng), 0, 0, n)
x0 = eng.export_x0()
X0 = {'frac': x0[7], 'v': x0[8], 'longitude': x0[9], 'latitude': x0[10], 'altitude': x0[12], 'azimuth': x0[13]}
for rep in range(2):
    t0 = time.time()
    res = source_map_arrays(X0, setup.radius_km, {}, 'source')
    wall = time.time() - t0
    print(f'K6 n={n}: kernel {eng.last_kernel_ms():.2f} ms, call (H2D + kernel + D2H) {wall * 1e3:.1f} ms, '
          f'pairs {int(res["n_total"].sum())}', flush=True)
Xs = {k: v[:ncpu] for k, v in X0.items()}
t0 = time.time()
ref = source_map.make_source_map(Xs, setup.radius_km, {}, 'source')
print(f'CPU oracle n={ncpu}: {time.time() - t0:.1f} s', flush=True)
