"""Functional miniature of ``astropy.units`` / ``astropy.constants`` -- just enough unit
algebra to EXECUTE the unmodified reference modules that the golden-vector
generator drives (``initial_state/source_distribution.py``,
``data_simulation/ModelResult.py``, ``ModelImage.py``, ``compute_iteration.py``,
``math/interpu.py``) in a container without astropy.

Test infrastructure only (used by tools/make_golden_products.py through
tools/refimport.py); nothing under nexoclom_b200/ imports it.  Values are the CODATA
2018 constants astropy 5.3 ships (the version pinned by the reference's poetry.lock).
"""
import sys
import types

import numpy as np

_DIMS = ('m', 's', 'kg', 'K', 'rad')


class Unit:
    """scale (SI) x dimension exponents; supports * / ** == and .to()."""

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        # ndarray (*|/) Unit, also in place (``v0 *= unit`` rebinds v0 to the Quantity)
        name = ufunc.__name__
        if method != '__call__' or name not in ('multiply', 'divide', 'true_divide'):
            return NotImplemented
        a, b = inputs
        if isinstance(b, Unit):
            arr = a
            unit = b if name == 'multiply' else Unit(1.0 / b.scale, tuple(-d for d in b.dims))
        else:
            arr = b if name == 'multiply' else 1.0 / np.asarray(b, dtype=float)
            unit = a
        if isinstance(arr, Quantity):
            return Quantity(np.asarray(arr), arr.unit * unit)
        return Quantity(arr, unit)

    def __init__(self, scale=1.0, dims=(0, 0, 0, 0, 0), name=None):
        self.scale = float(scale)
        self.dims = tuple(float(d) for d in dims)
        self.name = name

    # -- algebra between units
    def _combine(self, other, sign):
        return Unit(self.scale * other.scale ** sign,
                    tuple(a + sign * b for a, b in zip(self.dims, other.dims)))

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, 1)
        if isinstance(other, Quantity):
            return Quantity(np.asarray(other), self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(np.asarray(other), other.unit * self)
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.scale, tuple(-d for d in self.dims))
        if isinstance(other, Quantity):
            return Quantity(np.asarray(other), other.unit * inv)
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit(self.scale ** p, tuple(d * p for d in self.dims))

    def _same_dims(self, other):
        # radians are dimensionless for conversion purposes (astropy equivalency-free
        # code in the reference never mixes them)
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims))

    def __eq__(self, other):
        return (isinstance(other, Unit) and self._same_dims(other)
                and abs(self.scale - other.scale) <= 1e-15 * abs(self.scale))

    def __hash__(self):
        return hash((round(self.scale, 12), self.dims))

    def to(self, other, value=1.0):
        if not self._same_dims(other):
            raise ValueError(f'unit mismatch {self} -> {other}')
        return value * (self.scale / other.scale)

    def __repr__(self):
        return self.name or f'Unit({self.scale:g}, {self.dims})'

    __str__ = __repr__


dimensionless = Unit()


def _unit_of(x):
    return x.unit if isinstance(x, Quantity) else dimensionless


class Quantity(np.ndarray):
    __array_priority__ = 1e5

    def __new__(cls, value, unit=dimensionless):
        obj = np.array(value, dtype=np.float64, copy=True).view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, 'unit', dimensionless)

    @property
    def value(self):
        v = np.asarray(self)
        return float(v) if v.ndim == 0 else v

    def __getitem__(self, key):
        r = super().__getitem__(key)
        return r if isinstance(r, Quantity) else Quantity(r, self.unit)

    def to(self, unit):
        if isinstance(unit, Quantity):               # astropy: 1 / u.s is a unit
            unit = Unit(float(np.asarray(unit)) * unit.unit.scale, unit.unit.dims)
        return Quantity(np.asarray(self) * (self.unit.scale / unit.scale), unit) \
            if self.unit._same_dims(unit) else self.unit.to(unit)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raw = [np.asarray(x) if isinstance(x, Quantity) else x for x in inputs]
        out = kwargs.pop('out', None)
        if out is not None:
            kwargs['out'] = tuple(np.asarray(o) if isinstance(o, Quantity) else o for o in out)
        name = ufunc.__name__
        if method != '__call__':
            res = getattr(ufunc, method)(*raw, **kwargs)
            if isinstance(res, np.ndarray) or (method == 'reduce' and name in
                                               ('add', 'maximum', 'minimum')):
                return Quantity(res, _unit_of(inputs[0]))
            return res
        if name in ('add', 'subtract', 'maximum', 'minimum', 'remainder', 'fmod', 'less',
                    'less_equal', 'greater', 'greater_equal', 'equal', 'not_equal'):
            ua = _unit_of(inputs[0])
            ub = _unit_of(inputs[1])
            if isinstance(inputs[1], Quantity) and isinstance(inputs[0], Quantity) and ua != ub:
                raw[1] = raw[1] * (ub.scale / ua.scale)
                if not ua._same_dims(ub):
                    raise ValueError(f'unit mismatch in {name}: {ua} vs {ub}')
            res = ufunc(*raw, **kwargs)
            if res.dtype == bool:
                return res
            return Quantity(res, ua if isinstance(inputs[0], Quantity) else ub)
        if name == 'multiply':
            return Quantity(ufunc(*raw, **kwargs), _unit_of(inputs[0]) * _unit_of(inputs[1]))
        if name in ('divide', 'true_divide'):
            return Quantity(ufunc(*raw, **kwargs), _unit_of(inputs[0]) / _unit_of(inputs[1]))
        if name == 'sqrt':
            return Quantity(ufunc(*raw, **kwargs), _unit_of(inputs[0]) ** 0.5)
        if name == 'power':
            return Quantity(ufunc(*raw, **kwargs), _unit_of(inputs[0]) ** float(raw[1]))
        if name in ('negative', 'absolute', 'fabs', 'positive'):
            return Quantity(ufunc(*raw, **kwargs), _unit_of(inputs[0]))
        if name in ('arcsin', 'arccos', 'arctan', 'arctan2'):
            return Quantity(ufunc(*raw, **kwargs), rad)
        if name in ('sin', 'cos', 'tan', 'exp', 'log', 'isfinite', 'isnan', 'sign'):
            # angles are kept in radians by every caller in the reference
            res = ufunc(*raw, **kwargs)
            return res if res.dtype == bool else Quantity(res, dimensionless)
        res = ufunc(*raw, **kwargs)
        return res

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(np.asarray(self), self.unit * other)
        return np.multiply(self, other)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(np.asarray(self), self.unit / other)
        return np.divide(self, other)

    def __rtruediv__(self, other):
        return np.divide(other, self)

    def __rmul__(self, other):
        return np.multiply(other, self)

    def __eq__(self, other):
        return np.equal(self, other)

    def __ne__(self, other):
        return np.not_equal(self, other)

    def __hash__(self):
        return hash((float(np.asarray(self).sum()), self.unit.scale))

    def __float__(self):
        return float(np.asarray(self))

    def __repr__(self):
        return f'<Quantity {np.asarray(self)} {self.unit}>'


def _u(scale, m=0, s=0, kg=0, K=0, rad_=0, name=None):
    return Unit(scale, (m, s, kg, K, rad_), name)


m = _u(1, m=1, name='m'); cm = _u(1e-2, m=1, name='cm'); km = _u(1e3, m=1, name='km')
au = _u(1.495978707e11, m=1, name='au'); AA = _u(1e-10, m=1, name='AA')
s = _u(1, s=1, name='s'); h = _u(3600, s=1, name='h'); d = _u(86400, s=1, name='d')
kg = _u(1, kg=1, name='kg'); g = _u(1e-3, kg=1, name='g')
K = _u(1, K=1, name='K')
rad = _u(1, name='rad'); deg = _u(np.pi / 180, name='deg')
J = _u(1, m=2, s=-2, kg=1, name='J'); eV = _u(1.602176634e-19, m=2, s=-2, kg=1, name='eV')
R = _u(1e10 / (4 * np.pi), m=-2, s=-1, name='R'); kR = _u(1e13 / (4 * np.pi), m=-2, s=-1, name='kR')
dimensionless_unscaled = dimensionless
imperial = types.SimpleNamespace(mi=_u(1609.344, m=1, name='mi'))


def def_unit(name, represents=None, **kw):
    if isinstance(represents, Quantity):
        return Unit(float(np.asarray(represents)) * represents.unit.scale, represents.unit.dims,
                    name)
    if isinstance(represents, Unit):
        return Unit(represents.scale, represents.dims, name)
    return Unit(1.0, (0, 0, 0, 0, 0), name)


class _Const(Quantity):
    pass


constants = types.ModuleType('astropy.constants')
constants.k_B = Quantity(1.380649e-23, J / K)
constants.h = Quantity(6.62607015e-34, J * s)
constants.c = Quantity(299792458.0, m / s)
constants.u = Quantity(1.66053906660e-27, kg)
constants.G = Quantity(6.6743e-11, m ** 3 / kg / s ** 2)
constants.au = Quantity(1.495978707e11, m)


def install():
    """Register this module as ``astropy.units`` (and ``astropy.constants``)."""
    ap = types.ModuleType('astropy')
    ap.__path__ = []
    me = sys.modules[__name__]
    ap.units = me
    ap.constants = constants
    sys.modules['astropy'] = ap
    sys.modules['astropy.units'] = me
    sys.modules['astropy.constants'] = constants
    return me
