#!/usr/bin/env python
"""Small driver for ncu captures: one K1 + K2 (+K4) pass of the bench workload.
usage: prof_run.py [npackets] [workload]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np                                        # noqa: E402
from common import workload                               # noqa: E402
from nexoclom_b200._lib import ImageParams                # noqa: E402
from nexoclom_b200.engine import Engine                   # noqa: E402
from nexoclom_b200.ModelImage import image_rotation       # noqa: E402
from nexoclom_b200.runsetup import RunSetup               # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
wl = sys.argv[2] if len(sys.argv) > 2 else 'Na.maxwellian.radpres.input'
eng = Engine(0)
setup = RunSetup(workload(wl))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
eng.init_state(setup.source_params(eng), 0, 0, n)
if len(sys.argv) > 3 and sys.argv[3] == 'host':
    # the end-to-end path: streamed H2D behind one class-ordered persistent kernel
    import torch
    X0 = eng.export_x0()[:8]
    host = torch.empty((8, n), dtype=torch.float64).pin_memory()
    host.numpy()[:] = X0
    att, acc = eng.integrate_adaptive_host([host.numpy()[k] for k in range(8)], nchunks=16)
else:
    if len(sys.argv) > 3 and sys.argv[3] == 'stream':
        eng.set_option('schedule', 2)      # streaming kernel over the resident X0 slab
    att, acc = eng.integrate_adaptive()
ms = eng.last_kernel_ms()
ip = ImageParams()
M = image_rotation(0.0, np.pi / 2)
for k in range(9):
    ip.M[k] = float(M.flat[k])
ip.x0, ip.x1, ip.z0, ip.z1 = -4, 4, -4, 4
ip.nx = ip.nz = 800
ip.apix = 5.9e11
ip.vrplanet = setup.vrplanet
ip.quantity = 1
ip.round_f32 = 1
ip.skip_dead = 1
img, cnt = eng.image_accumulate(ip)
print(f'n={n} attempted={att} accepted={acc} k2_ms={ms:.3f} steps/s={att / ms * 1e3:.4g} '
      f'k4_ms={eng.last_kernel_ms():.4f} hits={int(cnt.sum())}')
