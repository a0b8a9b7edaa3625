#!/usr/bin/env python
"""Where the public API spends its host time: Input.run(n) -> ModelImage, and the import-mode
Output(X0=...) -> ModelImage, wall-clock per stage plus a cProfile of one pass."""
import cProfile
import io
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200 import Output, ModelImage
from nexoclom_b200.engine import get_engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
inputs = workload('Na.maxwellian.radpres.input')
params = {'quantity': 'radiance'}
eng = get_engine(0)


def stage(label, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print(f'  {label}: {(time.perf_counter() - t) * 1e3:.2f} ms', flush=True)
    return r


for it in range(3):
    print(f'pass {it} (device-drawn)')
    stage('delete_files', inputs.delete_files)
    stage('Input.run', lambda: inputs.run(n, seed=0, overwrite=True))
    stage('ModelImage', lambda: ModelImage(inputs, params))

out = Output(inputs, n, seed=0)
x0 = eng.export_x0() if eng.n == n else None
inputs.delete_files()
eng.init_state(__import__('nexoclom_b200.runsetup', fromlist=['get_setup']).get_setup(inputs).source_params(eng), 0, 0, n)
x0 = eng.export_x0()[:8]
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
host.copy_(torch.from_numpy(x0))
hn = host.numpy()
cols = {c: hn[k] for k, c in enumerate(('time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac'))}
for it in range(3):
    print(f'pass {it} (import mode)')
    stage('delete_files', inputs.delete_files)
    stage('Output(X0=)', lambda: Output(inputs, n, X0=cols))
    stage('ModelImage', lambda: ModelImage(inputs, params))

inputs.delete_files()
pr = cProfile.Profile()
pr.enable()
inputs.run(n, seed=0, overwrite=True)
im = ModelImage(inputs, params)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
