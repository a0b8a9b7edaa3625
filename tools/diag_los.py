#!/usr/bin/env python
"""K5 timing: synthetic MESSENGER-like LOS sweep over an integrated Na cloud."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200._lib import LosParams
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
from test_gpu_parity import _synthetic_los
eng = Engine(0)
setup = RunSetup(workload('Ca.isotropic.flat.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([4227]))
sp = setup.source_params(eng)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
nlos = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000
eng.init_state(sp, 0, 0, n)
eng.integrate_adaptive()
x = eng.export_state()
print('alive fraction', (x[7] > 0).mean(), 'r max', np.sqrt((x[1:4]**2).sum(0)).max())
los = _synthetic_los(nlos)
sx, sy, sz, bx, by, bz = los.T
dist = np.sqrt(sx**2 + sy**2 + sz**2)
ang = np.arccos((-sx * bx - sy * by - sz * bz) / dist)
dist = np.where(ang > np.arcsin(1. / dist), 1e30, dist)
lp = LosParams()
lp.dphi, lp.outeredge, lp.vrplanet, lp.rp_cm = np.radians(1.0), 15., setup.vrplanet, setup.radius_km * 1e5
lp.quantity, lp.round_f32, lp.skip_dead = 1, 1, 1
ref = None
for mode in (1, 2):
    eng.set_option('los_mode', mode)
    if mode == 1 and n * nlos > 5e10:
        continue
    for rep in range(2):
        rad, npk, inc = eng.los_accumulate(los.T.copy(), dist, lp)
        ms = eng.last_kernel_ms()
    print(f'mode={mode} n={n} nlos={nlos} ms={ms:.2f} pairs/s={n * nlos / ms * 1e3:.4g} hits={npk.sum()} '
          f'-> 1e5 LOS x 1e8 packets would take {1e13 / (n * nlos / ms * 1e3):.2f} s')
    if ref is None:
        ref = (rad, npk, inc)
    else:
        print('  grid == brute: counts', np.array_equal(npk, ref[1]), 'included', np.array_equal(inc, ref[2]),
              'max rel rad', np.max(np.abs(rad - ref[0]) / np.maximum(ref[0], 1e-300)))
