#!/usr/bin/env python
"""K1 (initial state) warm timing at 1e8 packets and agreement with the oracle's draw."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
from oracle import initial_state
eng = Engine(0)
for wl in ('Na.maxwellian.radpres.input', 'Ca.isotropic.flat.input', 'Na.bounce.input'):
    setup = RunSetup(workload(wl))
    setup.upload(eng)
    sp = setup.source_params(eng)
    eng.init_state(sp, 42, 1000, 200000)
    got = eng.export_x0().T
    ref = initial_state.draw_x0(setup, 200000, 42, first_id=1000)
    err = np.abs(got - ref).max(axis=0)
    n = 100_000_000
    best = 1e9
    for rep in range(4):
        eng.init_state(sp, 0, 0, n); eng.sync()
        best = min(best, eng.last_kernel_ms())
    print(f'{wl}: K1 {best:.3f} ms per 1e8 ({112e8 / best / 1e6:.0f} GB/s); max |device - oracle| per column {err.max():.2e}', flush=True)
