"""Container for surface / speed source maps (reference
``initial_state/SourceMap.py:8-85``).  Accepts a dict or a pickle of a dict; the
IDL ``.sav`` reader of the reference is out of scope."""
import pickle

import numpy as np

from .units import value_of


class _BareQuantity(np.ndarray):
    """What an astropy Quantity inside a reference-written pickle becomes here: the bare
    array (the maps are in radians / km/s by construction, reference SourceMap.py:36-60)."""

    def __setstate__(self, state):
        super().__setstate__(state[0] if isinstance(state[0], tuple) else state)


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        pass

    def __call__(self, *a, **k):
        return _Anything()


class _CompatUnpickler(pickle.Unpickler):
    """Reads source-map pickles written by the reference (they hold astropy Quantities)
    without astropy; everything else resolves normally."""

    def find_class(self, module, name):
        if module.startswith('astropy'):
            return _BareQuantity if name == 'Quantity' else _Anything
        return super().find_class(module, name)


def load_pickle(filename):
    with open(filename, 'rb') as f:
        return _CompatUnpickler(f).load()


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
                sourcemap = load_pickle(sourcemap)
                if not isinstance(sourcemap, dict):
                    sourcemap = dict(vars(sourcemap))
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
