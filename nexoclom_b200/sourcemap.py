"""Container for surface / speed source maps (reference
``initial_state/SourceMap.py:8-85``).  Accepts a dict or a pickle of a dict; the
IDL ``.sav`` reader of the reference is out of scope."""
import pickle

import numpy as np

from .units import value_of


class SourceMap:
    _fields = ('abundance', 'longitude', 'latitude', 'speed', 'speed_dist', 'azimuth',
               'azimuth_dist', 'altitude', 'altitude_dist', 'fraction_observed',
               'coordinate_system')

    def __init__(self, sourcemap=None):
        for f in self._fields:
            setattr(self, f, None)
        self.coordinate_system = 'solar-fixed'
        if isinstance(sourcemap, str):
            if sourcemap.endswith('.pkl'):
                with open(sourcemap, 'rb') as f:
                    sourcemap = pickle.load(f)
            else:
                raise NotImplementedError('only .pkl source maps are supported')
        if isinstance(sourcemap, dict):
            for k, v in sourcemap.items():
                if k in self._fields:
                    setattr(self, k, v if isinstance(v, str) or v is None
                            else np.asarray(value_of(v)))
        elif isinstance(sourcemap, SourceMap):
            self.__dict__.update(sourcemap.__dict__)
        elif sourcemap is not None:
            raise TypeError('sourcemap must be a dict, SourceMap or .pkl filename')
