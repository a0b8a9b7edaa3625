"""Import the reference's pure-NumPy hot-path modules WITHOUT importing the
``nexoclom`` package itself (its ``__init__`` needs PostgreSQL, astropy, ...).

Only usable where ``/root/reference`` exists (the build container).  Used by
``tools/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``tests/test_oracle_vs_reference.py`` cross-check (skipped on the GPU box).

A namespace stub ``nexoclom`` with ``__path__`` pointing at the reference tree
lets ``from nexoclom.particle_tracking.rk5 import rk5`` etc. resolve to the
UNMODIFIED reference files; modules that need absent third-party packages are
replaced by empty stubs so that ``particle_tracking/Output.py`` (the drivers)
can be imported too.
"""
import os
import sys
import types

REF = os.environ.get('NEXOCLOM_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'nexoclom', 'particle_tracking'))


class _AnyUnit:
    """Absorbs every unit operation the drivers' epilogues perform
    (``self.aplanet *= u.au`` ...); numerical results are already in place."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _AnyUnit()

    def __call__(self, *a, **k):
        return _AnyUnit()

    def _same(self, *a, **k):
        return _AnyUnit()

    __mul__ = __rmul__ = __truediv__ = __rtruediv__ = __pow__ = _same
    __imul__ = _same
    __array_priority__ = 1e6


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Install the stubs (idempotent) and return the reference functions."""
    global _installed
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    if not _installed:
        pkg = _stub('nexoclom', engine=None, config=None)
        pkg.__path__ = [os.path.join(REF, 'nexoclom')]
        pkg.__file__ = os.path.join(REF, 'nexoclom', '__init__.py')
        for sub in ('particle_tracking', 'initial_state'):
            m = _stub(f'nexoclom.{sub}')
            m.__path__ = [os.path.join(REF, 'nexoclom', sub)]
        u = _AnyUnit()
        if 'astropy' not in sys.modules:
            ap = _stub('astropy')
            ap.__path__ = []
            ap.units = _stub('astropy.units', __getattr__=lambda name: _AnyUnit())
            ap.constants = _stub('astropy.constants', __getattr__=lambda name: _AnyUnit())
        if 'sqlalchemy' not in sys.modules:
            sa = _stub('sqlalchemy')
            sa.__path__ = []
            _stub('sqlalchemy.dialects').__path__ = []
            _stub('sqlalchemy.dialects.postgresql')
        _stub('nexoclom.solarsystem', planet_dist=None)
        _stub('nexoclom.atomicdata', RadPresConst=None, atomicmass=None)
        _stub('nexoclom.initial_state.satellite_initial_positions',
              satellite_initial_positions=None)
        _stub('nexoclom.initial_state.LossInfo', LossInfo=None)
        _stub('nexoclom.initial_state.source_distribution', surface_distribution=None,
              speed_distribution=None, angular_distribution=None)
        _stub('nexoclom.particle_tracking.SurfaceInteraction', SurfaceInteraction=None)
        _installed = True

    from nexoclom.particle_tracking.rk5 import rk5
    from nexoclom.particle_tracking.state import state
    from nexoclom.particle_tracking.bouncepackets import bouncepackets, rebound_direction
    from nexoclom.initial_state.surface_temperature import surface_temperature
    from nexoclom.particle_tracking.Output import Output
    return types.SimpleNamespace(rk5=rk5, state=state, bouncepackets=bouncepackets,
                                 rebound_direction=rebound_direction,
                                 surface_temperature=surface_temperature, Output=Output)


class TaaQuantity(__import__('numpy').ndarray):
    """``geometry.taa`` stand-in: ``np.cos(taa)`` and the arithmetic after it must
    yield something with ``.value`` (reference ``surface_temperature.py:9-10``)."""

    def __new__(cls, v):
        import numpy as np
        return np.asarray(float(v), dtype=np.float64).view(cls)

    @property
    def value(self):
        return float(self)


class Lifetime(float):
    """``options.lifetime`` stand-in: comparable to 0 and has ``.value``."""
    @property
    def value(self):
        return float(self)


def fake_output(*, GM, vrplanet=0.0, radpres_v=None, radpres_a=None, gravity=True,
                radpres=True, lifetime=0.0, photo=None, step_size=0.0,
                resolution=1e-4, outeredge=1e30, endtime=0.0, stickcoef=1.0,
                sticktype='constant', accomfactor=None, A=None, taa=0.0,
                planet_radius_km=2440.53, startpoint='Mercury', seed=0):
    """Duck-typed ``Output`` carrying exactly the attributes the reference's
    ``rk5/state/bouncepackets`` and the two drivers read."""
    import numpy as np
    ns = types.SimpleNamespace
    out = ns()
    surf = ns(sticktype=sticktype, accomfactor=accomfactor)
    if sticktype == 'constant':
        surf.stickcoef = stickcoef
    if A is not None:
        surf.A = A
    out.inputs = ns(
        forces=ns(gravity=gravity, radpres=radpres),
        options=ns(lifetime=Lifetime(lifetime), step_size=step_size,
                   resolution=resolution, outeredge=outeredge,
                   endtime=Lifetime(endtime)),
        surfaceinteraction=surf,
        geometry=ns(startpoint=startpoint, taa=TaaQuantity(taa),
                    planet=ns(radius=ns(value=planet_radius_km))))
    out.GM = GM
    out.vrplanet = vrplanet
    out.radpres = (ns(velocity=np.asarray(radpres_v), accel=np.asarray(radpres_a))
                   if radpres_v is not None else None)
    out.loss_info = ns(photo=photo)
    out.randgen = np.random.default_rng(seed)
    out.aplanet = 1.0
    out.unit = _AnyUnit()
    return out
