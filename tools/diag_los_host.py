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
for G in [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]:
  eng.set_option('los_grid', G)
  print('los_grid', G)
  for it in range(3):
    eng.sync()
    t0 = time.perf_counter()
    rad, npk, inc = eng.los_accumulate(los, dplan, lp, n=n)
    print(f'  wall {(time.perf_counter() - t0) * 1e3:.2f} ms, kernels {eng.last_kernel_ms():.2f} ms, hits {int(npk.sum())}', flush=True)
# `used` sets: counts from the accumulate pass + indices from its candidate pairs, against the
# two extra searches of nx_los_used (1e4 lines of sight: the CSR is 4 B per used pair)
los2, dplan2 = bench.synthetic_los(10_000)
for it in range(2):
    eng.sync(); t0 = time.perf_counter()
    rad, npk, inc, cnt = eng.los_accumulate(los2, dplan2, lp, n=n, count_used=True)
    t1 = time.perf_counter()
    off, idx = eng.los_used_fill(los2, dplan2, lp, cnt, n=n)
    t2 = time.perf_counter()
    k_fill = eng.last_kernel_ms()
    rad, npk, inc = eng.los_accumulate(los2, dplan2, lp, n=n)
    t3 = time.perf_counter()
    off_b, idx_b = eng.los_used(los2, dplan2, lp, n=n)
    t4 = time.perf_counter()
    assert np.array_equal(off, off_b)
    print(f'counted accumulate {1e3 * (t1 - t0):.2f} ms + fill {1e3 * (t2 - t1):.2f} ms (kernel {k_fill:.2f}) | '
          f'accumulate {1e3 * (t3 - t2):.2f} ms + two-pass used {1e3 * (t4 - t3):.2f} ms; {int(off[-1])} used pairs', flush=True)
