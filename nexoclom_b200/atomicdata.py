"""Host-side atomic data tables: g-values, radiation-pressure acceleration,
photo-loss rates and atomic masses.

These are tiny (<= 827 points) and are built once per run on the host, then
uploaded to the GPU with ``nx_tables_upload``.  Behaviour follows the reference
``atomicdata/g_values.py:24-160``, ``atomicdata/photolossrates.py:66-86`` and
``atomicdata/atomicmass.py:5-51``; the raw tables in ``data/`` were converted from
the reference's pickles by ``tools/extract_reference_data.py``.

Constants are CODATA 2018 (what astropy 5.3, pinned by the reference's
``poetry.lock``, ships).  Atomic masses are the periodictable 1.6.1 values the
reference resolves through ``atomicmass()`` (Na = 22.98977 u reproduces the
reference's golden radiation-pressure table; see tests/test_host_tables.py).
"""
import json
import os
import functools

import numpy as np

from .units import Quantity

H_PLANCK = 6.62607015e-34      # J s
K_BOLTZMANN = 1.380649e-23     # J / K
AMU = 1.66053906660e-27        # kg
G_NEWTON = 6.6743e-11          # m3 / (kg s2)
AU_M = 1.495978707e11          # m
EV_J = 1.602176634e-19         # J

_DATADIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')

# periodictable 1.6.1 standard atomic weights for the species the model handles.
_ATOMIC_MASS = {
    'H': 1.00794, 'He': 4.002602, 'C': 12.0107, 'N': 14.0067, 'O': 15.9994,
    'Na': 22.98977, 'Mg': 24.305, 'Al': 26.981538, 'Si': 28.0855, 'S': 32.065,
    'K': 39.0983, 'Ca': 40.078, 'Ti': 47.867, 'Mn': 54.938049, 'Fe': 55.845,
}


def atomicmass(species):
    """Atomic mass in u (reference ``atomicdata/atomicmass.py:5-51``).
    Returns None (with the reference's warning) for unknown species."""
    if species in _ATOMIC_MASS:
        return Quantity(_ATOMIC_MASS[species], 'u')
    print(f'WARNING: mathMB.atomicmass: {species} not found')
    return None


@functools.lru_cache(maxsize=1)
def _gvalue_table():
    z = np.load(os.path.join(_DATADIR, 'gvalues.npz'))
    return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=1)
def _photorate_table():
    with open(os.path.join(_DATADIR, 'photorates.json')) as f:
        return json.load(f)


class gValue:
    """g-value [1/s] versus radial velocity [km/s] for one transition, scaled to
    heliocentric distance ``aplanet`` [AU] (reference ``g_values.py:59-99``)."""

    def __init__(self, sp, wavelength, aplanet=1.0):
        self.species = sp
        wavelength = float(wavelength.to('AA').value if isinstance(wavelength, Quantity)
                           else wavelength)
        aplanet = float(aplanet.to('au').value if isinstance(aplanet, Quantity) else aplanet)
        self.wavelength = Quantity(wavelength, 'AA')
        self.aplanet = Quantity(aplanet, 'au')

        tab = _gvalue_table()
        rows = (tab['species'] == sp) & (tab['wavelength'] == wavelength)
        if not rows.any():
            self.velocity = Quantity([0., 1.], 'km/s')
            self.g = Quantity([0., 0.], '1/s')
            self.filename = None
            self.reference = None
            print(f'Warning: g-values not found for species = {sp}')
        elif len(np.unique(tab['file_id'][rows])) == 1:
            vel = tab['velocity'][rows]
            g = tab['gvalue'][rows] * tab['refpoint'][rows]**2 / aplanet**2
            s = np.argsort(vel)
            self.velocity = Quantity(vel[s], 'km/s')
            self.g = Quantity(g[s], '1/s')
            self.filename = str(tab['file_names'][tab['file_id'][rows][0]])
            self.reference = None
        else:
            print('This should never happen')
            raise ValueError()


class RadPresConst:
    """Radiation acceleration [km/s^2] versus radial velocity [km/s]:
    ``a(v) = sum_lambda h g_lambda(v) / (m lambda)`` on the union velocity grid
    (reference ``g_values.py:134-160``, formula at :153)."""

    def __init__(self, species, aplanet):
        self.species = species
        aplanet = float(aplanet.to('au').value if isinstance(aplanet, Quantity) else aplanet)
        self.aplanet = Quantity(aplanet, 'au')

        tab = _gvalue_table()
        rows = tab['species'] == species
        if rows.any():
            waves = np.array(sorted(np.unique(tab['wavelength'][rows])))
            vel = np.array(sorted(np.unique(tab['velocity'][rows])))
            mass = atomicmass(species).value
            # (J s / u / AA) * (1/s) -> km/s^2
            to_kms2 = 1.0 / (AMU * 1e-10) * 1e-3
            rpres = np.zeros_like(vel)
            for wave in waves:
                gval = gValue(species, wave, aplanet)
                g_ = np.interp(vel, gval.velocity.value, gval.g.value)
                rpres_ = H_PLANCK / mass / wave * g_
                rpres += rpres_ * to_kms2
            self.wavelength = Quantity(waves, 'AA')
            self.velocity = Quantity(vel, 'km/s')
            self.accel = Quantity(rpres, 'km/s2')
        else:
            self.velocity = Quantity([0., 1.], 'km/s')
            self.accel = Quantity([0., 0.], 'km/s2')
            print(f'Warning: g-values not found for species = {species}')


class PhotoRate:
    """Total photo-reaction rate [1/s] at ``aplanet`` AU: sum of kappa / a^2 over
    all tabulated reactions (reference ``photolossrates.py:66-86``)."""

    def __init__(self, species, aplanet_=1.0):
        aplanet = float(aplanet_.value if isinstance(aplanet_, Quantity) else aplanet_)
        self.species = species
        self.aplanet = Quantity(aplanet, 'au')
        prates = [r for r in _photorate_table() if r['species'] == species]
        if len(prates) == 0:
            print('No photoreactions found')
            self.reactions = None
            self.rate = Quantity(1e-30, '1/s')
        else:
            rates = np.array([r['kappa'] / aplanet**2 for r in prates])
            self.reactions = prates
            self.rate = Quantity(rates.sum(), '1/s')

    def __str__(self):
        return (f'Species = {self.species}\n'
                f'Distance = {self.aplanet}\n'
                f'Rate = {self.rate}')


class LossInfo:
    """Loss-rate selection (reference ``initial_state/LossInfo.py:6-35``):
    ``lifetime < 0`` -> generic photo rate 1/|lifetime| (shadowed);
    ``lifetime == 0`` -> tabulated photo rate at ``aplanet``."""

    def __init__(self, atom, lifetime, aplanet):
        self.photo = 0.
        self.eimp = 0.
        self.chX = 0.
        self.reactions = []
        lifetime_ = float(lifetime.value if isinstance(lifetime, Quantity) else lifetime)
        if lifetime_ < 0:
            self.photo = abs(1. / lifetime_)
            self.reactions = 'Generic photo reaction'
        elif lifetime_ == 0:
            photo = PhotoRate(atom, aplanet)
            self.photo = float(photo.rate.value)
            self.reactions = ([r['reaction'] for r in photo.reactions]
                              if photo.reactions is not None else [])
        else:
            print('LossInfo objects should not be instantiated with lifetime > 0')
        if len(self.reactions) == 0:
            self.reactions = None

    def __len__(self):
        return len(self.reactions) if self.reactions is not None else 0
