#!/usr/bin/env python
"""K4 timing at several sizes (resident packets drawn by K1; no integration:
the initial state on the surface is as good as any for streaming)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200._lib import ImageParams
from nexoclom_b200.engine import Engine
from nexoclom_b200.ModelImage import image_rotation
from nexoclom_b200.runsetup import RunSetup
import torch
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
sp = setup.source_params(eng)
ip = ImageParams()
M = image_rotation(0.0, np.pi / 2)
for k in range(9):
    ip.M[k] = float(M.flat[k])
ip.x0, ip.x1, ip.z0, ip.z1 = -4, 4, -4, 4
ip.nx = ip.nz = 800
ip.apix = 5.9e11
ip.vrplanet = setup.vrplanet
img = torch.zeros((800, 800), dtype=torch.float64, device='cuda')
cnt = torch.zeros((800, 800), dtype=torch.int64, device='cuda')
for n in (10_000_000, 100_000_000):
    eng.init_state(sp, 0, 0, n)
    eng.sync()
    print(f'n={n} init_ms={eng.last_kernel_ms():.3f} ({184.0 * n / eng.last_kernel_ms() / 1e6:.0f} GB/s)')
    for quantity in (0, 1):
        for skip in (0, 1):
            ip.quantity, ip.skip_dead, ip.round_f32 = quantity, skip, 1
            best = 1e9
            for rep in range(5):
                eng.image_accumulate_dev(ip, img.data_ptr(), cnt.data_ptr(), n)
                eng.sync()
                best = min(best, eng.last_kernel_ms())
            print(f'  quantity={quantity} skip_dead={skip}: {best:.4f} ms  {40.0 * n / best / 1e6:.0f} GB/s '
                  f'({best * 1e8 / n:.3f} ms per 1e8)')
