#!/usr/bin/env python
"""How well does the longest-first cost predictor rank packets?"""
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
n = 2_000_000
eng.init_state(sp, 0, 0, n)
x0 = eng.export_x0()
eng.integrate_adaptive()
att, acc = eng.export_stats()
xf = eng.export_state()
t, x, y, z, vx, vy, vz, f = x0[:8]
mu = abs(setup.GM); res = 1e-4
r2 = x*x+y*y+z*z; r = np.sqrt(r2); v2 = vx*vx+vy*vy+vz*vz; rv = x*vx+y*vy+z*vz
en = 0.5*v2 - mu/r
tfl = t.copy()
bound = en < 0
a = np.where(bound, -mu/(2*np.where(bound, en, -1)), 1e30)
l2 = np.maximum(r2*v2 - rv*rv, 0)
e = np.sqrt(np.maximum(1 + 2*en*l2/mu**2, 0))
hit = bound & (a*(1-e) < 1) & (e > 1e-12)
c1 = np.clip((1-1/a)/np.maximum(e,1e-300), -1, 1); c0 = np.clip((1-r/a)/np.maximum(e,1e-300), -1, 1)
E1 = np.arccos(c1); E0 = np.arccos(c0); E0 = np.where(rv < 0, 2*np.pi-E0, E0); Ei = 2*np.pi - E1
tk = ((Ei - e*np.sin(Ei)) - (E0 - e*np.sin(E0)))*np.sqrt(np.where(bound, a, 1)**3/mu)
use = hit & (tk > 0) & (tk < tfl)
tfl = np.where(use, tk, tfl)
est = tfl*np.sqrt(v2)/(40*res*(1+r)) + 4
b = np.clip((2*np.log2(est)).astype(int), 0, 31)
print('bucket: count, mean att, p99 att, max att')
for k in range(32):
    m = b == k
    if m.any():
        print(f'{k:2d} {m.sum():8d} {att[m].mean():9.1f} {np.percentile(att[m],99):9.1f} {att[m].max():6d}   pred {2**(k/2):8.1f}')
big = att > 1500
print('att>1500:', big.sum(), 'their buckets:', np.bincount(b[big], minlength=32))
worst = np.argsort(att)[-10:]
for i in worst:
    print(f'att={att[i]} bucket={b[i]} est={est[i]:.0f} t={t[i]:.0f} v={np.sqrt(v2[i])*2440.53:.2f}km/s en={en[i]:.3e} bound={bound[i]} hit={hit[i]} tk={tk[i]:.0f} final_r={np.sqrt((xf[1:4,i]**2).sum()):.2f} frac={xf[7,i]:.3g} ratio att/acc={att[i]/max(acc[i],1):.2f}')
