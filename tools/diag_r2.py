#!/usr/bin/env python
"""Round-2 kernel timings: K1 (112 B/packet), K4 global vs privatised counts, K2 with the
split input / output slabs, device-side compaction, K3 with the row sink."""
import os, sys, json, time
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
res = {}
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


def best(fn, reps=5):
    b = 1e9
    for _ in range(reps):
        fn()
        eng.sync()
        b = min(b, eng.last_kernel_ms())
    return b


for n in (10_000_000, 100_000_000):
    t = best(lambda: eng.init_state(sp, 0, 0, n))
    res[f'k1_ms_{n:.0e}'] = t
    res[f'k1_gbs_{n:.0e}'] = 112.0 * n / t / 1e6
    print(f'K1 n={n}: {t:.3f} ms  {112.0 * n / t / 1e6:.0f} GB/s at 112 B/packet', flush=True)
    for quantity in (0, 1):
        for mode in (1, 2):
            eng.set_option('image_mode', mode)
            ip.quantity, ip.skip_dead, ip.round_f32 = quantity, 0, 1
            t = best(lambda: eng.image_accumulate_dev(ip, img.data_ptr(), cnt.data_ptr(), n))
            res[f'k4_alllive_q{quantity}_mode{mode}_ms_{n:.0e}'] = t
            print(f'  K4 all-live n={n} quantity={quantity} mode={mode}: {t:.4f} ms '
                  f'{40.0 * n / t / 1e6:.0f} GB/s', flush=True)
eng.set_option('image_mode', 0)

n = 10_000_000
eng.init_state(sp, 0, 0, n)
for rep in range(3):
    eng.rewind_state()
    att, acc = eng.integrate_adaptive()
    eng.sync()
    print(f'K2 n={n}: {eng.last_kernel_ms():.3f} ms, {att} attempted', flush=True)
res['k2_ms_1e7'] = eng.last_kernel_ms()
res['k2_attempted'] = att
for mode in (1, 2):
    eng.set_option('image_mode', mode)
    ip.quantity, ip.skip_dead, ip.round_f32 = 1, 1, 1
    t = best(lambda: eng.image_accumulate_dev(ip, img.data_ptr(), cnt.data_ptr(), n))
    res[f'k4_benchstate_mode{mode}_ms_1e7'] = t
    print(f'  K4 bench state (1% live) mode={mode}: {t:.4f} ms  {40.0 * n / t / 1e6:.0f} GB/s', flush=True)
eng.set_option('image_mode', 0)
t0 = time.perf_counter()
tab = eng.compact_state(skip_dead=True, round_f32=True)
eng.sync()
t1 = time.perf_counter()
res['compact_ms_1e7'] = eng.last_kernel_ms()
res['compact_wall_ms_1e7'] = (t1 - t0) * 1e3
res['compact_live'] = tab.n
print(f'compact: {eng.last_kernel_ms():.3f} ms kernel, {(t1 - t0) * 1e3:.3f} ms wall, {tab.n} live', flush=True)
t0 = time.perf_counter()
cols, index = tab.export()
print(f'export f32: {(time.perf_counter() - t0) * 1e3:.3f} ms', flush=True)
tab.free()

# K3 (configs[2] physics)
setup3 = RunSetup(workload('Na.bounce.input'))
setup3.upload(eng)
eng.upload_gtables(setup3.gtables([5891, 5897]))
sp3 = setup3.source_params(eng)
n3 = 2_000_000
for rep in range(2):
    eng.init_state(sp3, 0, 0, n3)
    _, nsteps, steps = eng.integrate_constant(seed=1)
    eng.sync()
    t = eng.last_kernel_ms()
print(f'K3 n={n3}: {t:.3f} ms  {steps / t / 1e6:.3f} e9 packet-steps/s', flush=True)
res['k3_ms_2e6'] = t
res['k3_steps_per_s'] = steps / t * 1e3
eng.init_state(sp3, 0, 0, n3)
t0 = time.perf_counter()
tab, _, steps2 = eng.integrate_constant_rows(seed=1, skip_dead=True)
eng.sync()
t1 = time.perf_counter()
print(f'K3 rows: wall {(t1 - t0) * 1e3:.2f} ms (count pass + fill pass), last kernel '
      f'{eng.last_kernel_ms():.3f} ms, {tab.n} rows', flush=True)
res['k3_rows_wall_ms_2e6'] = (t1 - t0) * 1e3
res['k3_rows_fill_ms_2e6'] = eng.last_kernel_ms()
res['k3_rows'] = tab.n
tab.free()
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/diag_r2.json', 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
