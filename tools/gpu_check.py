#!/usr/bin/env python
"""First-light GPU check (run under gpurun): microbenchmarks, adaptive-driver
parity against the oracle, image parity, a timing run."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from common import workload, oracle_constants, state_parity   # noqa: E402
from nexoclom_b200.engine import Engine                      # noqa: E402
from nexoclom_b200.runsetup import RunSetup                  # noqa: E402
from nexoclom_b200._lib import ImageParams                   # noqa: E402
from oracle import tracking, imaging                         # noqa: E402

res = {}
eng = Engine(0)
res['fp64_peak_tflops'] = eng.measure_fp64_peak()
res['copy_gbs'] = eng.measure_copy_bw(1 << 30)
print(res, flush=True)

inputs = workload('Na.maxwellian.radpres.input')
for strict in (0, 1):
    setup = RunSetup(inputs, strict_math=bool(strict))
    setup.upload(eng)
    sp = setup.source_params(eng)
    n = 4000
    eng.init_state(sp, 0, 0, n)
    x0 = eng.export_x0()
    X0 = x0[:8].T.copy()
    att, acc = eng.integrate_adaptive()
    Xg = eng.export_state().T
    a_g, c_g = eng.export_stats()
    rc = oracle_constants(setup)
    t0 = time.time()
    Xo, a_o, c_o = tracking.integrate_adaptive(X0, rc)
    res[f'oracle_s_{strict}'] = time.time() - t0
    par = state_parity(Xg, Xo)
    par['att_equal'] = float((a_g == a_o).mean())
    par['acc_equal'] = float((c_g == c_o).mean())
    par['att_total'] = (int(att), int(a_o.sum()))
    res[f'adaptive_parity_strict{strict}'] = par
    print(strict, par, flush=True)

# image parity on the final states (fast mode result resident)
setup = RunSetup(inputs)
M = np.asarray(imaging.image_rotation(0.0, np.pi / 2))
for quantity in (0, 1):
    ip = ImageParams()
    for k in range(9):
        ip.M[k] = M.flat[k]
    ip.x0, ip.x1, ip.z0, ip.z1 = -4, 4, -4, 4
    ip.nx = ip.nz = 800
    rcm = setup.radius_km * 1e5
    ip.apix = (8 / 800 * rcm) * (8 / 800 * rcm)
    ip.vrplanet = setup.vrplanet
    ip.quantity = quantity
    ip.round_f32 = 1
    ip.skip_dead = 1
    gt = setup.gtables([5891, 5897])
    eng.upload_gtables(gt)
    img, cnt = eng.image_accumulate(ip)
    X = eng.export_state().T
    X = X[X[:, 7] > 0].astype(np.float32).astype(np.float64)
    oi, oc, _, _ = imaging.create_image(X[:, 1], X[:, 2], X[:, 3], X[:, 5], X[:, 7],
                                        vrplanet=setup.vrplanet, M=M, dims=[800, 800],
                                        xrange=(-4, 4), zrange=(-4, 4), apix=ip.apix,
                                        quantity='radiance' if quantity else 'column', gtables=gt)
    res[f'image_q{quantity}'] = dict(counts_equal=bool(np.array_equal(cnt, oc.astype(np.int64))),
                                     nz=int((oc > 0).sum()),
                                     maxrel=float(np.abs(img - oi).max() / max(oi.max(), 1e-300)))
    print(res[f'image_q{quantity}'], flush=True)

# timing run
setup = RunSetup(inputs)
setup.upload(eng)
sp = setup.source_params(eng)
for n in (1_000_000, 4_000_000):
    eng.init_state(sp, 1, 0, n)
    t_init = eng.last_kernel_ms()
    att, acc = eng.integrate_adaptive()
    ms = eng.last_kernel_ms()
    res[f'timing_{n}'] = dict(init_ms=t_init, ms=ms, attempted=att, accepted=acc,
                              steps_per_s=att / (ms * 1e-3),
                              frac_fp64=att / (ms * 1e-3) * 764 / (res['fp64_peak_tflops'] * 1e12))
    print(res[f'timing_{n}'], flush=True)
    ip.quantity = 1
    img, cnt = eng.image_accumulate(ip)
    res[f'image_ms_{n}'] = eng.last_kernel_ms()
    print('image ms', res[f'image_ms_{n}'], flush=True)

os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/gpu_check.json', 'w') as f:
    json.dump(res, f, indent=1, default=str)
print(json.dumps(res, indent=1, default=str))
