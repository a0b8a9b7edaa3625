"""Base class of ``ModelImage`` and ``LOSResult`` (reference
``data_simulation/ModelResult.py:10-170``): parses ``params``, validates
``quantity``, picks default emission wavelengths, and provides
``packet_weighting`` -- whose arithmetic runs inside the K4 / K5 kernels; the
host method here only prepares the g-value tables the kernels interpolate."""
import copy
import os

from .input_classes import InputError
from .units import Quantity, def_unit


class ModelResult:
    def __init__(self, inputs, params):
        self.inputs = copy.deepcopy(inputs)
        self.outid, self.outputfiles, _, _ = self.inputs.search()
        self.npackets = 0
        self.totalsource = 0.
        self.atoms_per_packet = 0.
        self.sourcerate = Quantity(0., '')          # unit: 10**23 atoms/s
        if isinstance(params, str):
            if os.path.exists(params):
                self.params = {}
                with open(params, 'r') as f:
                    for line in f:
                        if ';' in line:
                            line = line[:line.find(';')]
                        elif '#' in line:
                            line = line[:line.find('#')]
                        if '=' in line:
                            p, v = line.split('=')
                            self.params[p.strip().lower()] = v.strip()
            else:
                raise FileNotFoundError('ModelResult.__init__', 'params file not found.')
        elif isinstance(params, dict):
            self.params = params
        else:
            raise TypeError('ModelResult.__init__', 'params must be a dict or filename.')

        quantities = ('column', 'radiance', 'density', 'difrad')
        self.quantity = self.params.get('quantity', None)
        if (self.quantity is None) or (self.quantity not in quantities):
            raise InputError('ModelImage.__init__', "quantity must be 'column' or 'radiance'")

        self.g = self.params.get('g', None)
        if self.quantity == 'radiance':
            self.mechanism = ['resonant scattering']
            species = self.inputs.options.species
            if 'wavelength' in self.params:
                self.wavelength = tuple(sorted(Quantity(int(m.strip()), 'AA')
                                               for m in self.params['wavelength'].split(',')))
            elif species is None:
                raise InputError('ModelImage.__init__',
                                 'Must provide either species or params.wavelength')
            elif species == 'Na':
                self.wavelength = (Quantity(5891, 'AA'), Quantity(5897, 'AA'))
            elif species == 'Ca':
                self.wavelength = (Quantity(4227, 'AA'),)
            elif species == 'Mg':
                self.wavelength = (Quantity(2852, 'AA'),)
            else:
                raise InputError('ModelResult.__init__',
                                 f'Default wavelengths not available for {species}')
        else:
            self.mechanism = None
            self.wavelength = None

        planet = self.inputs.geometry.planet
        self.unit = def_unit('R_' + planet.object, 'length', float(planet.radius.value) * 1e3)

    def _upload_weighting_tables(self, engine, setup):
        """Device tables for ``packet_weighting`` (ModelResult.py:140-170):
        column -> weight = frac; radiance -> frac * sunlit * sum_lambda g(v_r)/1e6."""
        if self.quantity in ('column', 'density'):
            return 0
        if self.quantity in ('radiance', 'difrad'):
            key = (id(setup), self.g, tuple(float(w) for w in self.wavelength))
            if getattr(engine, '_uploaded_gtables', None) == key:
                return 1
            if self.g is not None:
                # user-supplied constant g-value (ModelResult.py:158-159): a two-point
                # table whose clamped ends make np.interp return g everywhere
                import numpy as np
                g = float(np.asarray(self.g))
                engine.upload_gtables([(np.array([-1.0, 1.0]), np.array([g, g]))])
            else:
                engine.upload_gtables(setup.gtables(self.wavelength))
            engine._uploaded_gtables = key
            return 1
        raise InputError('ModelResults.packet_weighting', f'{self.quantity} is invalid.')
