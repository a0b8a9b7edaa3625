#!/usr/bin/env python
"""K5 parameter sweep on the bench workload (final state of 1e7 Na packets, ALL packets
in the grid): cells per axis and the scale of the asinh spacing."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200._lib import LosParams
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nlos = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
eng.init_state(setup.source_params(eng), 0, 0, n)
eng.integrate_adaptive()
g = torch.Generator(device='cpu').manual_seed(1)
th = torch.rand(nlos, generator=g, dtype=torch.float64) * 2 * np.pi
rr = 1.1 + 4.9 * torch.rand(nlos, generator=g, dtype=torch.float64)
x_sc = torch.stack([0.3 * rr * torch.cos(th), 0.2 * rr * torch.cos(th) - 0.5, rr * torch.sin(th)], dim=0)
x_sc *= torch.clamp(x_sc.norm(dim=0), min=1.1) / x_sc.norm(dim=0)
tgt = torch.randn(3, nlos, generator=g, dtype=torch.float64)
tgt *= (1 + 3 * torch.rand(nlos, generator=g, dtype=torch.float64)) / tgt.norm(dim=0)
bore = tgt - x_sc
bore /= bore.norm(dim=0)
dplan = x_sc.norm(dim=0)
ang = torch.arccos(-(x_sc * bore).sum(dim=0) / dplan)
dplan = torch.where(ang > torch.arcsin(1. / dplan), torch.full_like(dplan, 1e30), dplan)
los = torch.cat([x_sc, bore], dim=0).numpy().copy()
dist = dplan.numpy().copy()
lp = LosParams()
lp.dphi, lp.outeredge = float(np.radians(1.0)), 25.0
lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
lp.quantity, lp.round_f32 = 1, 1
eng.set_option('los_mode', 2)
combos = [tuple(int(v) for v in a.split(':')) for a in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['128:1000'])]
for skip, order in ((0, 0), (0, 1), (1, 0), (1, 1)):
    lp.skip_dead = skip
    eng.set_option('los_order', order)
    for G, sc in combos:
        eng.set_option('los_grid', G)
        eng.set_option('los_grid_scale_milli', sc)
        for rep in range(2):
            rad, npk, inc = eng.los_accumulate(los, dist, lp)
            ms = eng.last_kernel_ms()
        print(f'skip_dead={skip} order={order} G={G} scale={sc / 1000}: {ms:.2f} ms hits={int(npk.sum())}', flush=True)
