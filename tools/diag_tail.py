#!/usr/bin/env python
"""Diagnostics for the K2 tail: step-count distribution, t(n) fit, ordering on/off."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
for order in (1, 2, 0):
    eng.set_option('order_packets', order)
    ts = []
    for n in (1_250_000, 2_500_000, 5_000_000, 10_000_000):
        best = 1e9
        for rep in range(3):
            eng.init_state(sp, 0, 0, n)
            att, acc = eng.integrate_adaptive()
            best = min(best, eng.last_kernel_ms())
        ts.append((n, best, att))
        print(f'order={order} n={n} ms={best:.3f} steps/s={att / best * 1e3:.4g}', flush=True)
    ns = np.array([t[0] for t in ts], float); ms = np.array([t[1] for t in ts])
    b, a = np.polyfit(ns, ms, 1)
    print(f'order={order}: t = {a:.3f} ms + {b * 1e6:.4f} ms per 1e6 packets; asymptotic {ts[-1][2] / ts[-1][0] / b / 1e-3:.4g} steps/s')
a_, c_ = eng.export_stats()
print('att percentiles 50/90/99/99.9/99.99/max:', np.percentile(a_, [50, 90, 99, 99.9, 99.99]), a_.max(), 'mean', a_.mean())
print('packets with att > 2000:', int((a_ > 2000).sum()), ' > 4000:', int((a_ > 4000).sum()))
