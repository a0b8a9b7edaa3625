"""Minimal stand-in for the slice of ``astropy.units`` the nexoclom API exposes.

The reference hands astropy Quantities to users (``inputs.options.endtime.value``,
``inputs.geometry.taa``, ``planet.radius`` ...; reference
``initial_state/input_classes.py:83-96, 1060-1075``).  astropy is not available on
the GPU boxes, so this module provides a small ``Quantity`` that keeps the two
things user code touches -- ``.value`` and ``.to(unit)`` between a fixed set of
named units -- and otherwise behaves like a float / ndarray.  All numerical work
on the hot path is done on plain float64, never through this class.
"""
import numpy as np

# SI scale of every named unit this package uses, grouped by physical dimension.
_UNITS = {
    # length
    'm': ('length', 1.0), 'cm': ('length', 1e-2), 'km': ('length', 1e3),
    'au': ('length', 1.495978707e11), 'AA': ('length', 1e-10),
    # time
    's': ('time', 1.0), 'h': ('time', 3600.0), 'd': ('time', 86400.0),
    # angle
    'rad': ('angle', 1.0), 'deg': ('angle', np.pi / 180.0),
    # speed / acceleration
    'km/s': ('speed', 1e3), 'm/s': ('speed', 1.0), 'km/s2': ('accel', 1e3),
    # misc
    'K': ('temperature', 1.0), 'kg': ('mass', 1.0), 'u': ('mass', 1.66053906660e-27),
    'eV': ('energy', 1.602176634e-19), 'J': ('energy', 1.0), '1/s': ('rate', 1.0),
    'cm2': ('area', 1e-4), 'km2': ('area', 1e6), 'm3/s2': ('gm', 1.0),
    '': ('dimensionless', 1.0),
}


def def_unit(name, dimension, si_scale):
    """Register a run-specific unit such as ``R_Mercury`` (reference
    ``particle_tracking/Output.py:102``)."""
    _UNITS[name] = (dimension, float(si_scale))
    return name


class Quantity(np.ndarray):
    """ndarray subclass with ``.value``, ``.unit`` and ``.to()``."""

    def __new__(cls, value, unit=''):
        obj = np.asarray(value, dtype=np.float64).view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, 'unit', '')

    @property
    def value(self):
        v = np.asarray(self)
        return float(v) if v.ndim == 0 else v

    def to(self, unit):
        if unit == self.unit:
            return Quantity(np.asarray(self), unit)
        d0, s0 = _UNITS[self.unit]
        d1, s1 = _UNITS[unit]
        if d0 != d1:
            raise ValueError(f"cannot convert '{self.unit}' ({d0}) to '{unit}' ({d1})")
        return Quantity(np.asarray(self) * (s0 / s1), unit)

    def __eq__(self, other):
        if isinstance(other, Quantity) and other.unit != self.unit:
            try:
                other = other.to(self.unit)
            except (ValueError, KeyError):
                return False
        return np.asarray(self).__eq__(np.asarray(other))

    def __ne__(self, other):
        return np.logical_not(self.__eq__(other))

    def __hash__(self):
        return hash((float(np.asarray(self).sum()), self.unit))

    def __repr__(self):
        return f'{self.value} {self.unit}'.strip()

    __str__ = __repr__

    def __format__(self, spec):
        v = self.value
        if isinstance(v, float):
            return format(v, spec) + (f' {self.unit}' if self.unit else '')
        return repr(self)

    def __reduce__(self):
        return (Quantity, (np.asarray(self), self.unit))


def value_of(x):
    """Plain float / ndarray from a Quantity or a number."""
    return x.value if isinstance(x, Quantity) else x
