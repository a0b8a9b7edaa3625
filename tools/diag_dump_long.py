#!/usr/bin/env python
"""Dump (X0, attempted) of the long packets + a random sample of a shard for offline work on the cost model."""
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
out = {}
for r in [int(a) for a in sys.argv[1:]] or [0, 6]:
    eng.init_state(sp, 0, r * n, n)
    x0 = eng.export_x0()[:8]
    eng.integrate_adaptive()
    att, acc = eng.export_stats()
    xf = eng.export_state()[:8]
    g = np.random.default_rng(r)
    sel = np.unique(np.concatenate([np.nonzero(att > 1200)[0], g.choice(n, 60_000, replace=False)]))
    out[f"x0_{r}"] = x0[:, sel]
    out[f'xf_{r}'] = xf[:, sel].astype(np.float32)
    out[f'att_{r}'] = att[sel]
    out[f'sel_{r}'] = sel
    out[f'hist_{r}'] = np.bincount(np.minimum(att, 8191), minlength=8192)
    print(r, len(sel), int(att.sum()))
out['GM'] = setup.GM
np.savez_compressed(os.path.join(os.path.dirname(HERE), 'gpurun_out', 'long_packets.npz'), **out)
