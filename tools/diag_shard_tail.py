#!/usr/bin/env python
"""Which packets end a shard's K2 late?  start time ~ (steps queued before it) / rate, end = start +
attempted x per-step latency of a lone packet."""
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
n = 10_000_000
mu = abs(setup.GM); res = 1e-4
for r in [int(a) for a in sys.argv[1:]] or [0, 6]:
    eng.init_state(sp, 0, r * n, n)
    x0 = eng.export_x0()
    best = 1e9
    for rep in range(2):
        eng.init_state(sp, 0, r * n, n)
        eng.integrate_adaptive()
        best = min(best, eng.last_kernel_ms())
    att, acc = eng.export_stats()
    t, x, y, z, vx, vy, vz, f = x0[:8]
    r2 = x*x+y*y+z*z; rr = np.sqrt(r2); v2 = vx*vx+vy*vy+vz*vz; rv = x*vx+y*vy+z*vz
    en = 0.5*v2 - mu/rr
    tfl = t.copy()
    bound = en < 0
    a = np.where(bound, -mu/(2*np.where(bound, en, -1)), 1e30)
    l2 = np.maximum(r2*v2 - rv*rv, 0)
    e = np.sqrt(np.maximum(1 + 2*en*l2/mu**2, 0))
    hit = bound & (a*(1-e) < 1) & (e > 1e-12)
    c1 = np.clip((1-1/a)/np.maximum(e,1e-300), -1, 1); c0 = np.clip((1-rr/a)/np.maximum(e,1e-300), -1, 1)
    E1 = np.arccos(c1); E0 = np.arccos(c0); E0 = np.where(rv < 0, 2*np.pi-E0, E0); Ei = 2*np.pi - E1
    tk = ((Ei - e*np.sin(Ei)) - (E0 - e*np.sin(E0)))*np.sqrt(np.where(bound, a, 1)**3/mu)
    use = hit & (tk > 0) & (tk < tfl)
    tfl = np.where(use, tk, tfl)
    est = tfl*np.sqrt(v2)/(40*res*(1+rr)) + 4
    b = np.clip((2*np.log2(est)).astype(int), 0, 31)
    b[~((t > res) & (f > 0))] = 0
    order = np.argsort(-b, kind='stable')
    cum = np.cumsum(att[order].astype(np.float64))
    start = np.empty(n); start[order] = (cum - att[order]) / cum[-1] * best
    lat = 2.75e-3
    end = start + att * lat
    worst = np.argsort(end)[-8:][::-1]
    print(f'shard {r}: K2 {best:.3f} ms; bucket histogram of att>3000: {np.bincount(b[att > 3000], minlength=32)[16:].tolist()} (buckets 16..31)')
    for i in worst:
        print(f'   att {att[i]:5d} bucket {b[i]:2d} est {est[i]:7.0f} start~{start[i]:6.2f} ms end~{end[i]:6.2f} ms')
