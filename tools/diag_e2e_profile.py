#!/usr/bin/env python
"""cProfile of the bench's end-to-end call: Output(inputs, n, X0=<pinned host columns>) ->
ModelImage(inputs, {quantity: radiance}) -> image on the host."""
import os, sys, time, cProfile, pstats, io
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200 import Output, ModelImage
from nexoclom_b200.engine import get_engine
from nexoclom_b200.runsetup import get_setup
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
inputs = workload('Na.maxwellian.radpres.input')
eng = get_engine(0)
setup = get_setup(inputs); setup.upload(eng)
eng.init_state(setup.source_params(eng), 0, 0, n)
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
host.numpy()[:] = eng.export_x0()[:8]
cols = {c: host.numpy()[k] for k, c in enumerate(('time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac'))}
params = {'quantity': 'radiance'}
def once():
    inputs.delete_files()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = Output(inputs, n, X0=cols, first_id=0)
    t1 = time.perf_counter()
    im = ModelImage(inputs, params)
    s = float(im.image.sum())
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, out.kernel_ms
for _ in range(4):
    print('Output %.2f ms, ModelImage %.2f ms, kernels %.2f ms' % once(), flush=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    once()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45); print(s.getvalue()[:9000])
