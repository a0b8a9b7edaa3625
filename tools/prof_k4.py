#!/usr/bin/env python
"""Driver for ncu captures of K4 on an all-live state: prof_k4.py [npackets] [quantity]"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
from bench import image_params
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
quantity = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
eng.upload_gtables(setup.gtables([5891, 5897]))
eng.init_state(setup.source_params(eng), 0, 0, n)
ip, _ = image_params(setup, quantity=quantity, skip_dead=0)
eng.image_begin(800, 800)
for rep in range(3):
    eng.image_add(ip, n)
    eng.sync()
    print(f'K4 quantity={quantity} n={n}: {eng.last_kernel_ms():.4f} ms', flush=True)
