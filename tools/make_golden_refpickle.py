#!/usr/bin/env python
"""tests/golden/reference_output.pkl: an Output file in the REFERENCE's on-disk format
(reference particle_tracking/Output.py:480-548: ``pickle.dump(self)``), for the reader in
nexoclom_b200/refpickle.py.  Build container only (needs /root/reference).

The reference package cannot be imported (PostgreSQL, astropy), so the object graph is put
together here from the two things that ARE available:

* the inputs, the planet (``SSObject``) and every astropy Quantity / unit come out of the
  reference's OWN pickle fixture ``tests/test_data/input_classes_data.pkl`` (written by the
  reference with real astropy): they are unpickled into stand-in classes living in fake
  modules named like the real ones, each re-pickling through the same protocol it was read
  with -- ``astropy.units.core._recreate_irreducible_unit(cls, names, registered)`` + state
  for irreducible units, plain objects with ``_names / _represents`` or ``_scale / _bases /
  _powers`` for the other units, ``(ndarray state, {'_unit': ...})`` for ``Quantity``;
* the ``Output`` instance gets the attributes reference ``Output.__init__`` / ``save`` set
  (Output.py:87-202, 522-543): ``inputs, planet, randgen, compress, unit, GM, aplanet,
  vrplanet, loss_info, radpres, X0, X (float32 / int32), npackets, totalsource, idnum,
  filename``; the packets are an oracle-made cloud, saved the way ``save`` leaves them.

The same packets go to tests/golden/reference_output_packets.npz for the tests.
"""
import os
import pickle
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
REF = os.environ.get('NEXOCLOM_REFERENCE', '/root/reference')
GOLD = os.path.join(REPO, 'tests', 'golden')


def fake_module(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition('.')
    if parent:
        if parent not in sys.modules:
            fake_module(parent)
        setattr(sys.modules[parent], leaf, m)
    return m


def plain_class(module, name):
    def __new__(cls, *args, **kwargs):
        return object.__new__(cls)

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state
    cls = type(name, (), {'__module__': module, '__qualname__': name, '__new__': __new__,
                          '__init__': __init__, '__setstate__': __setstate__})
    setattr(sys.modules.get(module) or fake_module(module), name, cls)
    return cls


def install_fakes():
    core = fake_module('astropy.units.core')
    quantity = fake_module('astropy.units.quantity')

    def _recreate_irreducible_unit(cls, names, registered):
        unit = cls.__new__(cls)
        unit._names = list(names)
        return unit
    _recreate_irreducible_unit.__module__ = 'astropy.units.core'
    _recreate_irreducible_unit.__qualname__ = '_recreate_irreducible_unit'
    core._recreate_irreducible_unit = _recreate_irreducible_unit

    class IrreducibleUnit:
        def __reduce__(self):
            return (_recreate_irreducible_unit, (self.__class__, list(self._names), True),
                    self.__dict__)
    IrreducibleUnit.__module__, IrreducibleUnit.__qualname__ = 'astropy.units.core', 'IrreducibleUnit'
    core.IrreducibleUnit = IrreducibleUnit
    for name in ('Unit', 'PrefixUnit', 'CompositeUnit', 'NamedUnit'):
        plain_class('astropy.units.core', name)

    class Quantity(np.ndarray):
        def __reduce__(self):
            state = list(super().__reduce__())
            state[2] = (state[2], self.__dict__)
            return tuple(state)

        def __setstate__(self, state):
            nd_state, own = state
            super().__setstate__(nd_state)
            self.__dict__.update(own)
    Quantity.__module__, Quantity.__qualname__ = 'astropy.units.quantity', 'Quantity'
    quantity.Quantity = Quantity
    for mod, names in (('astropy.time.core', ['Time']), ('astropy.time.formats', ['TimeISOT']),
                       ('nexoclom.initial_state.Input', ['Input']),
                       ('nexoclom.initial_state.input_classes',
                        ['Geometry', 'SurfaceInteraction', 'Forces', 'SpatialDist', 'SpeedDist',
                         'AngularDist', 'Options']),
                       ('nexoclom.solarsystem.SSObject', ['SSObject']),
                       ('nexoclom.particle_tracking.Output', ['Output']),
                       ('nexoclom.initial_state.LossInfo', ['LossInfo']),
                       ('nexoclom.atomicdata.g_values', ['RadPresConst'])):
        for n in names:
            plain_class(mod, n)
    return Quantity


class FixtureUnpickler(pickle.Unpickler):
    """The reference's fixture was written by an older layout (nexoclom.modelcode.*)."""
    MOVED = {'nexoclom.modelcode.Input': 'nexoclom.initial_state.Input',
             'nexoclom.modelcode.input_classes': 'nexoclom.initial_state.input_classes'}

    def find_class(self, module, name):
        return super().find_class(self.MOVED.get(module, module), name)


def main():
    Quantity = install_fakes()
    with open(os.path.join(REF, 'tests/test_data/input_classes_data.pkl'), 'rb') as f:
        names, inputs_all = FixtureUnpickler(f).load()
    inputs = inputs_all[names.index('test_data/inputfiles/Ca.isotropic.flat.input')]
    planet = inputs.geometry.planet
    km = planet.radius._unit                    # genuine astropy unit objects
    sec = inputs.options.endtime._unit
    au = planet.a._unit
    core = sys.modules['astropy.units.core']

    def q(value, unit):
        out = np.asarray(value, dtype=np.float64).view(Quantity)
        out._unit = unit
        return out

    def composite(scale, bases, powers):
        c = core.CompositeUnit.__new__(core.CompositeUnit)
        c._scale, c._bases, c._powers = scale, list(bases), list(powers)
        return c

    def named(names, represents):
        n = core.Unit.__new__(core.Unit)
        n._names, n._short_names, n._long_names, n._format = list(names), list(names), [], {}
        n._represents = represents
        n.__doc__ = names[0]
        return n

    r_km = float(np.asarray(planet.radius))
    unit = named(['R_' + planet.object], composite(r_km, [km], [1]))   # u.def_unit('R_Mercury', radius)

    # an oracle-made cloud of packets (the reference's own run needs PostgreSQL)
    from nexoclom_b200 import Input as OurInput
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state
    wl = os.path.join(REPO, 'nexoclom_b200', 'workloads', 'Ca.isotropic.flat.input')
    setup = RunSetup(OurInput(wl))
    n = 4000
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac', 'v', 'longitude', 'latitude',
            'local_time', 'altitude', 'azimuth']
    x0 = initial_state.draw_x0(setup, n, 31)
    X0 = pd.DataFrame(x0, columns=cols)
    rng = np.random.default_rng(8)
    t = rng.random(n) * 4000.0
    X = X0.drop(['longitude', 'latitude', 'local_time'], axis=1).copy()
    for c, v in (('x', 'vx'), ('y', 'vy'), ('z', 'vz')):
        X[c] = X0[c] + X0[v] * t
    X['time'] = X0['time'] - t
    X['frac'] = rng.random(n) * (rng.random(n) > 0.25)
    X['lossfrac'] = np.zeros(n)
    X['step_size'] = 1000.0 * rng.random(n)          # the adaptive driver's column
    X['Index'] = np.arange(n, dtype=np.int64)

    out = sys.modules['nexoclom.particle_tracking.Output'].Output()
    out.inputs = inputs
    out.planet = planet
    out.randgen = np.random.default_rng(seed=5)
    out.compress = True
    out.unit = unit
    out.GM = setup.GM                                # plain floats while the run is going on ...
    out.aplanet = setup.aplanet
    out.vrplanet = setup.vrplanet
    loss = sys.modules['nexoclom.initial_state.LossInfo'].LossInfo()
    loss.photo, loss.photo_factor, loss.reactions = 7.0e-5 / setup.aplanet ** 2, 1.0, None
    out.loss_info = loss
    out.radpres = None
    out.npackets = n
    out.totalsource = float(X0['frac'].sum())
    # ... "Add units back in" (Output.py:361-366)
    out.aplanet = q(out.aplanet, au)
    out.vrplanet = q(out.vrplanet * r_km, composite(1.0, [km, sec], [1, -1]))
    out.GM = q(out.GM, composite(1.0, [unit, sec], [3, -2]))
    # save (Output.py:522-548): frac > 0 rows, 32 bit
    out.idnum = 17
    out.filename = '/data/modeloutputs/Mercury/Ca/uniform/flat/000/0000000017.pkl'
    X = X[X.frac > 0]
    for frame in (X0, X):
        for column in frame:
            if frame[column].dtype == np.int64:
                frame[column] = frame[column].astype(np.int32)
            elif frame[column].dtype == np.float64:
                frame[column] = frame[column].astype(np.float32)
    out.X0, out.X = X0, X
    dest = os.path.join(GOLD, 'reference_output.pkl')
    with open(dest, 'wb') as f:
        pickle.dump(out, f, protocol=4)
    np.savez_compressed(os.path.join(GOLD, 'reference_output_packets.npz'),
                        X0=X0.values, X=X[['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']].values,
                        index=X['Index'].values, npackets=n, totalsource=out.totalsource,
                        vrplanet_kms=float(np.asarray(out.vrplanet)), aplanet=setup.aplanet)
    print(dest, os.path.getsize(dest), 'bytes;', len(X), 'rows kept of', n)


if __name__ == '__main__':
    main()
