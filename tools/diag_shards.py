#!/usr/bin/env python
"""K2 time per shard of a sharded run (sorted kernel and the streaming host-buffer path):
why ranks differ -- step totals vs the tail of mispredicted long packets."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
import numpy as np
import torch
from common import workload
from nexoclom_b200.engine import Engine
from nexoclom_b200.runsetup import RunSetup
eng = Engine(0)
setup = RunSetup(workload('Na.maxwellian.radpres.input'))
setup.upload(eng)
sp = setup.source_params(eng)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
models = [int(m) for m in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 3]
host = torch.empty((8, n), dtype=torch.float64).pin_memory()
for r in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    line = f'shard {r}:'
    for model in models:
        eng.set_option('order_packets', model)
        best = 1e9
        for rep in range(3):
            eng.init_state(sp, 0, r * n, n)
            att, acc = eng.integrate_adaptive()
            best = min(best, eng.last_kernel_ms())
        if model == models[0]:
            host.numpy()[:] = eng.export_x0()[:8]
            cols = [host.numpy()[k] for k in range(8)]
        bs = 1e9
        for rep in range(3):
            att2, _ = eng.integrate_adaptive_host(cols, nchunks=32)
            bs = min(bs, eng.last_kernel_ms())
        assert att2 == att
        line += f'  model {model}: K2 {best:.3f} ms, host-buffer path {bs:.3f} ms;'
    print(line, f'steps {att}', flush=True)
