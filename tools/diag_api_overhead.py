#!/usr/bin/env python
"""Host-side overhead of the public Output API (wall time and cProfile of Output(inputs, n))."""
import os, sys, time, cProfile, pstats, io
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200 import Output
inputs = workload('Na.maxwellian.radpres.input')
inputs.delete_files()
Output(inputs, 1000, seed=0)          # warm up (library load, tables)
inputs.delete_files()
for n in (1_000_000, 4_000_000):
    t = time.time(); out = Output(inputs, n, seed=1); dt = time.time() - t
    print(f'Output({n}): {dt:.3f} s wall', flush=True)
    inputs.delete_files()
pr = cProfile.Profile(); pr.enable()
out = Output(inputs, 4_000_000, seed=2)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(25); print(s.getvalue()[:6000])
t = time.time(); o2 = Output.restore(out.filename); print('restore', time.time() - t)
inputs.delete_files()
