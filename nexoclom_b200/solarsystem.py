"""Solar-system bodies and heliocentric distance / radial velocity.

Host-side scalars feeding the kernels (GM, planet radius, a_planet, v_r planet).
Follows the reference ``solarsystem/SSObject.py:28-100`` and
``solarsystem/planet_dist.py:9-74``; the constants table was converted from the
reference's ``data/PlanetaryConstants.pkl`` by ``tools/extract_reference_data.py``.
"""
import functools
import json
import os

import numpy as np

from .atomicdata import G_NEWTON, AU_M
from .units import Quantity

_DATADIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


@functools.lru_cache(maxsize=1)
def _constants():
    with open(os.path.join(_DATADIR, 'planetary_constants.json')) as f:
        return json.load(f)


class SSObject:
    """A solar-system body; ``GM`` is NEGATIVE (``-mass*G``, reference
    ``SSObject.py:53``), ``moons`` is a list of SSObjects or None."""

    def __init__(self, obj):
        rows = [r for r in _constants() if r['Object'].casefold() == obj.casefold()]
        if len(rows) == 1:
            row = rows[0]
            self.object = row['Object']
            self.orbits = row['orbits']
            self.radius = Quantity(row['radius'], 'km')
            self.mass = Quantity(row['mass'], 'kg')
            self.e = row['e']
            self.tilt = Quantity(row['tilt'], 'deg')
            self.rotperiod = Quantity(row['rot_period'], 'h')
            self.orbperiod = Quantity(row['orb_period'], 'd')
            self.GM = Quantity(-row['mass'] * G_NEWTON, 'm3/s2')
            self.moons = [SSObject(r['Object']) for r in _constants()
                          if r['orbits'] == self.object]
            if len(self.moons) == 0:
                self.moons = None
            if self.orbits == 'Milky Way':
                self.type = 'Star'
                self.a = Quantity(row['a'], 'km')
            elif self.orbits == 'Sun':
                self.type = 'Planet'
                self.a = Quantity(row['a'], 'au')
            else:
                self.type = 'Moon'
                self.a = Quantity(row['a'], 'km')
        else:
            print(f'Object {obj} does not exist in table.')
            self.object = None

    def __len__(self):
        return 1 if self.moons is None else len(self.moons) + 1

    def __eq__(self, other):
        return isinstance(other, SSObject) and self.object == other.object

    def __hash__(self):
        return hash((self.object,))

    def __repr__(self):
        return f'SSObject({self.object})'


def planet_dist(planet_, taa=None, time=None):
    """Distance from the Sun [AU] and radial velocity [km/s] at true anomaly
    ``taa`` [rad] (reference ``planet_dist.py:29-74``): Kepler ellipse for r,
    and v_r from finite differences of r over a 1000-point mean-anomaly ladder
    with a third-order series for the true anomaly, linearly interpolated."""
    if isinstance(planet_, str):
        planet = SSObject(planet_)
        if planet.object is None:
            return None
    elif isinstance(planet_, SSObject):
        planet = planet_
    else:
        raise TypeError('solarsystemMB.planet_dist',
                        'Must give a SSObject or a object name.')
    if time is not None:
        raise NotImplementedError
    if taa is None:
        print('Neither a time nor a true anomaly was given.')
        return None

    a = float(planet.a.value)
    eps = planet.e
    if isinstance(taa, Quantity):
        taa_ = float(taa.to('rad').value)
    elif type(taa) in (int, float, np.float64):
        taa_ = float(taa)
    else:
        raise TypeError('taa must be a number or angle quantity')

    if eps > 0:
        r = a * (1 - eps**2) / (1 + eps * np.cos(taa_))
        period = float(planet.orbperiod.to('s').value)
        t = np.linspace(0, 1, 1000) * period
        t = np.concatenate([np.array([t[0] - t[1]]), t])
        mean_anomaly = np.linspace(0, 2 * np.pi, 1000)
        mean_anomaly = np.concatenate(
            [np.array([mean_anomaly[0] - mean_anomaly[1]]), mean_anomaly])
        true_anomaly = (mean_anomaly +
                        (2 * eps - eps**3 / 4) * np.sin(mean_anomaly) +
                        5 / 4 * eps**2 * np.sin(2 * mean_anomaly) +
                        13 / 12 * eps**3 * np.sin(3 * mean_anomaly))
        r_true = a * (1 - eps**2) / (1 + eps * np.cos(true_anomaly))
        drdt = (r_true[1:] - r_true[:-1]) / (t[1:] - t[:-1])   # AU / s
        drdt_kms = drdt * (AU_M / 1e3)
        v_r = float(np.interp(taa_, true_anomaly[1:], drdt_kms))
    else:
        r, v_r = a, 0.
    return Quantity(r, 'au'), Quantity(v_r, 'km/s')
