"""Reading files written by the reference.

The reference keeps a run as ``pickle.dump(self)`` of its ``Output`` object (reference
``particle_tracking/Output.py:480-548``): an object graph of ``nexoclom.*`` classes --
``Output``, ``Input``, the seven ``input_classes``, ``SSObject``, ``LossInfo`` ... -- pandas
DataFrames (``X0``, ``X``, float32 / int32 columns) and astropy ``Quantity`` / unit objects.
Neither ``nexoclom`` (PostgreSQL at import) nor ``astropy`` is importable next to this
package, so the stream is read with an ``Unpickler`` that resolves

* the reference's classes to this package's drop-in classes of the same name (both the
  current module layout and the older ``nexoclom.modelcode.*`` one, which the reference's own
  fixture ``tests/test_data/input_classes_data.pkl`` was written with);
* ``astropy.units.quantity.Quantity`` -- an ndarray subclass pickled as
  ``(ndarray state, {'_unit': unit})`` -- and the unit classes (``IrreducibleUnit`` through
  ``_recreate_irreducible_unit``, named ``Unit`` / ``PrefixUnit`` with ``_names`` and
  ``_represents``, ``CompositeUnit`` with ``_scale / _bases / _powers``) to light stand-ins,
  converted to ``nexoclom_b200.units.Quantity`` once the graph is complete;
* every other class of ``nexoclom`` / ``astropy`` / ``sqlalchemy`` to an attribute bag.

The protocol facts above are pinned against the reference's own pickles (the fixture named
above and ``tests/unit_tests/atomicdata/g_value_test_data.pkl``), see
``tests/test_refpickle.py``.  Files written by this package go through the same reader.
"""
import pickle

import numpy as np

from .units import Quantity, _UNITS, def_unit

# astropy unit name -> name in units._UNITS
_UNIT_ALIASES = {'AU': 'au', 'au': 'au', 'Angstrom': 'AA', 'AA': 'AA', 'angstrom': 'AA',
                 'kilometer': 'km', 'meter': 'm', 'second': 's', 'hour': 'h', 'day': 'd',
                 'radian': 'rad', 'degree': 'deg', 'Kelvin': 'K', 'kilogram': 'kg'}
_SI_BASE = {'m': ('length', 1.0), 's': ('time', 1.0), 'kg': ('mass', 1.0), 'rad': ('angle', 1.0),
            'K': ('temperature', 1.0)}


class _Bag:
    """Stand-in for a class this package has no counterpart of: keeps the state."""

    def __init__(self, *args, **kwargs):
        self._args = args

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            if isinstance(state[0], dict):
                self.__dict__.update(state[0])
            self.__dict__.update(state[1])
        else:
            self._state = state

    def __repr__(self):
        return f'<{type(self).__name__} {sorted(self.__dict__)[:8]}>'


class _RefUnit(_Bag):
    """Any astropy unit object."""

    def name(self):
        """(name, dimension, SI scale) of the unit in this package's vocabulary."""
        names = self.__dict__.get('_names')
        if names:
            nm = _UNIT_ALIASES.get(names[0], names[0])
            if nm in _UNITS:
                return nm, _UNITS[nm][0], _UNITS[nm][1]
            rep = self.__dict__.get('_represents')
            if isinstance(rep, _RefUnit):                 # def_unit('R_Mercury', 2440.53 km)
                _, dim, scale = rep.name()
                return nm, dim, scale
            return nm, 'unknown:' + nm, 1.0
        bases = self.__dict__.get('_bases')
        if bases is not None:                             # CompositeUnit
            powers = self.__dict__.get('_powers', [1] * len(bases))
            scale = float(self.__dict__.get('_scale', 1.0))
            num, den, dims = [], [], []
            for b, p in zip(bases, powers):
                bn, bdim, bscale = b.name() if isinstance(b, _RefUnit) else (str(b), '?', 1.0)
                scale *= bscale ** float(p)
                dims.append(f'{bdim}^{p}')
                p_ = abs(p)
                tag = bn if p_ == 1 else f'{bn}{int(p_) if float(p_).is_integer() else p_}'
                (num if p > 0 else den).append(tag)
            nm = ' '.join(num) if num else ('1' if den else '')
            if den:
                nm += '/' + ' '.join(den)
            if nm in _UNITS:
                return nm, _UNITS[nm][0], _UNITS[nm][1]
            if len(bases) == 1 and powers[0] == 1:        # scale * one base: a length etc.
                return nm, dims[0][:-2], scale
            return nm, ' '.join(dims), scale
        return '', 'dimensionless', 1.0


def _recreate_unit(cls, names, registered=True):
    """astropy.units.core._recreate_irreducible_unit"""
    unit = _RefUnit()
    unit.__dict__['_names'] = list(names)
    return unit


class _RefQuantity(np.ndarray):
    """astropy Quantity as pickled: ndarray.__reduce__ state + the instance dict."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            super().__setstate__(state[0])
            self._own = state[1]
        else:
            super().__setstate__(state)
            self._own = {}

    def __array_finalize__(self, obj):
        self._own = getattr(obj, '_own', {})

    def convert(self):
        unit = self._own.get('_unit')
        name = ''
        if isinstance(unit, _RefUnit):
            name, dim, scale = unit.name()
            if name not in _UNITS:
                def_unit(name, dim, scale)
        return Quantity(np.asarray(self, dtype=np.float64), name)


def _drop_in(name):
    """This package's class for a reference class name, or None."""
    from importlib import import_module
    pkg = __name__.rsplit('.', 1)[0]
    input_classes = import_module(pkg + '.input_classes')
    table = {'Input': import_module(pkg + '.Input').Input,
             'Output': import_module(pkg + '.Output').Output,
             'SSObject': import_module(pkg + '.solarsystem').SSObject}
    for cls in ('Geometry', 'SurfaceInteraction', 'Forces', 'SpatialDist', 'SpeedDist',
                'AngularDist', 'Options'):
        table[cls] = getattr(input_classes, cls)
    return table.get(name)


class RefUnpickler(pickle.Unpickler):
    _bags = {}

    def find_class(self, module, name):
        root = module.split('.')[0]
        if root == 'astropy':
            if name == 'Quantity':
                return _RefQuantity
            if name == '_recreate_irreducible_unit':
                return _recreate_unit
            if module.startswith('astropy.units') and name[:1].isupper():
                return _RefUnit
        if root == 'nexoclom':
            # input_classes.SurfaceInteraction is the input group; the class of the same name
            # in particle_tracking is the run-time accommodation table (not needed to read)
            if not (name == 'SurfaceInteraction' and 'input_classes' not in module):
                cls = _drop_in(name)
                if cls is not None:
                    return cls
        if root in ('astropy', 'nexoclom', 'sqlalchemy', 'refunits'):
            key = module + '.' + name
            if name[:1].islower() or name[:1] == '_':          # a reconstruction function
                def rebuild(*args, **kwargs):
                    bag = _Bag()
                    bag._function, bag._args = key, args
                    return bag
                return rebuild
            if key not in self._bags:
                self._bags[key] = type(name, (_Bag,), {'_reference_class': key})
            return self._bags[key]
        return super().find_class(module, name)


def _convert(obj, seen):
    """Replace every unpickled astropy Quantity below `obj` by this package's Quantity."""
    if isinstance(obj, _RefQuantity):
        return obj.convert()
    if id(obj) in seen:
        return obj
    if isinstance(obj, (str, bytes, int, float, bool, type(None), np.ndarray, np.generic)):
        return obj
    seen.add(id(obj))
    if isinstance(obj, dict):
        for k in list(obj):
            obj[k] = _convert(obj[k], seen)
        return obj
    if isinstance(obj, list):
        for i, v in enumerate(obj):
            obj[i] = _convert(v, seen)
        return obj
    if isinstance(obj, tuple):
        return tuple(_convert(v, seen) for v in obj)
    if isinstance(obj, (set, frozenset)):
        for v in obj:
            _convert(v, seen)
        return obj
    d = getattr(obj, '__dict__', None)
    if isinstance(d, dict) and type(obj).__module__.split('.')[0] in ('nexoclom_b200',):
        for k in list(d):
            d[k] = _convert(d[k], seen)
    return obj


def load(source):
    """Unpickle a file (path or binary file object) written by the reference or by this
    package; astropy Quantities arrive as ``nexoclom_b200.units.Quantity``."""
    if isinstance(source, (str, bytes)) or hasattr(source, '__fspath__'):
        with open(source, 'rb') as f:
            obj = RefUnpickler(f).load()
    else:
        obj = RefUnpickler(source).load()
    return _convert(obj, set())
