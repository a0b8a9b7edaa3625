#!/usr/bin/env python
"""K3 timing: BASELINE configs[2] (Na, T-dependent sticking + bounce + accommodation,
constant 30 s step, 361 steps) with the image fused into the integrator."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200._lib import ImageParams
from nexoclom_b200.engine import Engine
from nexoclom_b200.ModelImage import image_rotation
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
stricts = (0,) if (len(sys.argv) > 2 and sys.argv[2] == 'fast') else (0, 1)
for strict in stricts:
    setup = RunSetup(workload('Na.bounce.input'), strict_math=bool(strict))
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
    ip.quantity, ip.round_f32, ip.skip_dead = 1, 1, 1
    img = torch.zeros((800, 800), dtype=torch.float64, device='cuda')
    cnt = torch.zeros((800, 800), dtype=torch.int64, device='cuda')
    for fused in (0, 1):
        best, steps = 1e9, 0
        for rep in range(2):
            eng.init_state(sp, 0, 0, n)
            img.zero_(); cnt.zero_()
            _, nsteps, steps = eng.integrate_constant(
                seed=1, image_params=ip if fused else None,
                image_dev=img.data_ptr() if fused else None,
                counts_dev=cnt.data_ptr() if fused else None)
            best = min(best, eng.last_kernel_ms())
        print(f'strict={strict} fused_image={fused} n={n} nsteps={nsteps} ms={best:.2f} packet-steps={steps} '
              f'steps/s={steps / best * 1e3:.4g} rows_in_image={int(cnt.sum())}', flush=True)
