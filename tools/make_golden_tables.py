#!/usr/bin/env python
"""tests/golden/host_tables.npz: outputs of the UNMODIFIED reference host-table builders
(build container only; needs /root/reference), through tools/reftables.py:

  SurfaceInteraction.__init__ (SurfaceInteraction.py:10-61) for the three surface blocks the
      parity runs use (temperature-dependent sticking at TAA 3.14 and 1.3, constant sticking
      0.5): ``probgrid`` [201, 101], the temperature / probability axes, the accommodation
      spline ``v_interp`` and the sticking closure evaluated on committed sample points;
  planet_dist (planet_dist.py:29-74): (r, v_r) of Mercury at TAA 0, 1.3, 3.14 (the TAAs of
      the workloads) and on a 64-point sweep, Jupiter and Mars at three TAAs;
  SSObject (SSObject.py:27-71): the constants of Mercury, Jupiter, Io.

tests/test_host_tables.py replays them through nexoclom_b200's ports (bit-exact).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

import reftables                                       # noqa: E402
from common import workload                            # noqa: E402
from nexoclom_b200.units import Quantity               # noqa: E402

GOLD = os.path.join(REPO, 'tests', 'golden')

SURFACE_CASES = (('tdep314', 'Na.bounce.input', None),
                 ('tdep130', 'Na.bounce.input', 1.3),
                 ('c05', 'Na.bounce.stick05.input', None))


def case_inputs(wl, taa):
    inputs = workload(wl)
    if taa is not None:
        inputs.geometry.taa = Quantity(float(taa), 'rad')
    return inputs


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    for tag, wl, taa in SURFACE_CASES:
        inputs = case_inputs(wl, taa)
        si = reftables.surface_interaction(inputs)
        temperature = np.asarray(si.temperature, dtype=float)
        out[f'{tag}_probgrid'] = np.asarray(si.probgrid, dtype=float)
        out[f'{tag}_temperature'] = temperature
        out[f'{tag}_probability'] = np.asarray(si.probability, dtype=float)
        T = rng.uniform(temperature.min(), temperature.max(), 2000)
        P = rng.random(2000)
        out[f'{tag}_sample_T'], out[f'{tag}_sample_P'] = T, P
        out[f'{tag}_v_interp'] = np.asarray(si.v_interp(T, P), dtype=float)
        if hasattr(si, 'stickcoef'):
            lon = rng.random(2000) * 2 * np.pi
            lat = np.arcsin(rng.random(2000) * 2 - 1)
            out[f'{tag}_lon'], out[f'{tag}_lat'] = lon, lat
            out[f'{tag}_stickcoef'] = np.asarray(si.stickcoef(lon, lat), dtype=float)
        print(tag, 'probgrid', out[f'{tag}_probgrid'].shape,
              'T range', temperature.min(), temperature.max())

    taas = np.concatenate([[0., 1.3, 3.14], np.linspace(0, 2 * np.pi, 64, endpoint=False)])
    out['mercury_taa'] = taas
    out['mercury_r_vr'] = np.array([reftables.planet_dist('Mercury', t) for t in taas])
    for planet in ('Jupiter', 'Mars'):
        t3 = np.array([0.5, 2.0, 4.5])
        out[f'{planet.lower()}_taa'] = t3
        out[f'{planet.lower()}_r_vr'] = np.array([reftables.planet_dist(planet, t) for t in t3])
    for name in ('Mercury', 'Jupiter', 'Io'):
        o = reftables.ssobject(name)
        out[f'ss_{name.lower()}'] = np.array(
            [float(np.asarray(o.radius)), float(np.asarray(o.mass)), float(np.asarray(o.a)),
             float(o.e), float(np.asarray(o.orbperiod)), float(np.asarray(o.GM))])
    np.savez_compressed(os.path.join(GOLD, 'host_tables.npz'), **out)
    print('host_tables.npz', {k: v.shape for k, v in out.items() if k.startswith('mercury')})


if __name__ == '__main__':
    main()
