#!/usr/bin/env python
"""Convert the reference's physical DATA tables into this repo's own formats.

Run once in the build container (needs /root/reference).  Only data is
converted -- no reference source code is read or copied:

  nexoclom/data/g-values/g-values.pkl   -> nexoclom_b200/data/gvalues.npz
  nexoclom/data/Loss/photorates.pkl     -> nexoclom_b200/data/photorates.json
  nexoclom/data/PlanetaryConstants.pkl  -> nexoclom_b200/data/planetary_constants.json
  tests/unit_tests/atomicdata/g_value_test_data.pkl
                                        -> tests/golden/gvalue_golden.npz

The golden pickle holds astropy Quantities; astropy is not installed here, so a
stub Unpickler maps Quantity onto a bare ndarray subclass (units are known from
the reference docstrings: km/s, 1/s, km/s**2, AU, Angstrom).
"""
import json
import os
import pickle
import sys

import numpy as np
import pandas as pd

REF = os.environ.get('NEXOCLOM_REFERENCE', '/root/reference')
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Q(np.ndarray):
    def __setstate__(self, state):
        super().__setstate__(state[0])


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        pass

    def __call__(self, *a, **k):
        return _Dummy()


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('astropy'):
            return _Q if name == 'Quantity' else _Dummy
        return super().find_class(module, name)


def main():
    datadir = os.path.join(REPO, 'nexoclom_b200', 'data')
    golddir = os.path.join(REPO, 'tests', 'golden')
    os.makedirs(datadir, exist_ok=True)
    os.makedirs(golddir, exist_ok=True)

    # ---- g-values ----
    g = pd.read_pickle(os.path.join(REF, 'nexoclom/data/g-values/g-values.pkl'))
    files = sorted(g.filename.unique())
    np.savez_compressed(
        os.path.join(datadir, 'gvalues.npz'),
        species=np.array(g.species.values, dtype='U8'),
        wavelength=g.wavelength.values.astype(np.float64),
        velocity=g.velocity.values.astype(np.float64),
        gvalue=g.gvalue.values.astype(np.float64),
        refpoint=g.refpoint.values.astype(np.float64),
        file_id=np.array([files.index(f) for f in g.filename.values], dtype=np.int32),
        file_names=np.array(files, dtype='U64'),
        row_order=np.arange(len(g), dtype=np.int64))
    print('gvalues rows', len(g))

    # ---- photo rates ----
    ph = pd.read_pickle(os.path.join(REF, 'nexoclom/data/Loss/photorates.pkl'))
    rows = [dict(species=str(r.species), reaction=str(r.reaction), kappa=float(r.kappa),
                 reference=str(r.reference), best_version=bool(r.best_version))
            for r in ph.itertuples()]
    with open(os.path.join(datadir, 'photorates.json'), 'w') as f:
        json.dump(rows, f, indent=1)
    print('photorates rows', len(rows))

    # ---- planetary constants ----
    pc = pd.read_pickle(os.path.join(REF, 'nexoclom/data/PlanetaryConstants.pkl'))
    rows = [dict(Object=str(r.Object), orbits=str(r.orbits), radius=float(r.radius),
                 mass=float(r.mass), a=float(r.a), e=float(r.e), tilt=float(r.tilt),
                 rot_period=float(r.rot_period), orb_period=float(r.orb_period))
            for r in pc.itertuples()]
    with open(os.path.join(datadir, 'planetary_constants.json'), 'w') as f:
        json.dump(rows, f, indent=1)
    print('planetary rows', len(rows))

    # ---- golden g-value / radiation-pressure vectors (reference test fixture) ----
    with open(os.path.join(REF, 'tests/unit_tests/atomicdata/g_value_test_data.pkl'), 'rb') as f:
        gv, rp = _StubUnpickler(f).load()
    out = {}
    for i, d in enumerate(gv):
        out[f'g{i}_species'] = np.array(d['species'])
        out[f'g{i}_wavelength'] = np.float64(d['wavelength'])
        out[f'g{i}_aplanet'] = np.float64(d['aplanet'])
        out[f'g{i}_velocity'] = np.asarray(d['velocity'], dtype=np.float64)
        out[f'g{i}_g'] = np.asarray(d['g'], dtype=np.float64)
    for i, d in enumerate(rp):
        out[f'r{i}_species'] = np.array(d['species'])
        out[f'r{i}_aplanet'] = np.float64(d['aplanet'])
        out[f'r{i}_velocity'] = np.asarray(d['velocity'], dtype=np.float64)
        out[f'r{i}_accel'] = np.asarray(d['accel'], dtype=np.float64)
    np.savez_compressed(os.path.join(golddir, 'gvalue_golden.npz'), **out)
    print('golden g-values written')


if __name__ == '__main__':
    sys.exit(main())
