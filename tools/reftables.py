"""Execute the UNMODIFIED reference builders of the per-run host tables
(build container only; needs /root/reference):

  particle_tracking/SurfaceInteraction.py:10-61   sticking closure + accommodation
                                                  ``probgrid`` + RectBivariateSpline
  solarsystem/planet_dist.py:9-74                 distance / radial velocity at a TAA
  solarsystem/SSObject.py:27-71                   planetary constants

astropy is absent here: ``tools/refunits.py`` supplies the unit algebra, periodictable
is replaced by the five masses the reference's own docstring / golden pickle pin
(``make_golden_products.install``).  ``scipy.misc.derivative`` -- imported by
planet_dist.py:3 and never called -- no longer exists in SciPy and is stubbed.

The objects returned keep working after ``purge()`` (their closures hold the reference
functions), so ``tools/make_golden.py`` builds them first and then installs the lighter
``refimport`` stubs the drivers need.
"""
import sys
import types

import numpy as np


def install():
    import make_golden_products as mg
    mg.install()
    if 'scipy.misc' not in sys.modules or not hasattr(sys.modules['scipy.misc'], 'derivative'):
        m = types.ModuleType('scipy.misc')
        m.derivative = None
        sys.modules['scipy.misc'] = m
    sys.modules.pop('nexoclom.particle_tracking.SurfaceInteraction', None)
    import refunits as u
    return u


def purge():
    """Forget the stub packages so that another installer (refimport) starts clean."""
    for name in list(sys.modules):
        if name == 'nexoclom' or name.startswith('nexoclom.') or name == 'astropy' or \
                name.startswith('astropy.'):
            del sys.modules[name]
    import refimport
    refimport._installed = False


def reference_inputs(inputs):
    """The attributes SurfaceInteraction.__init__ reads, as the reference's own types."""
    import refunits as u
    ns = types.SimpleNamespace
    sint = inputs.surfaceinteraction
    surf = ns(sticktype=sint.sticktype, accomfactor=sint.accomfactor)
    if hasattr(sint, 'A'):
        surf.A = sint.A
    if hasattr(sint, 'stickcoef'):
        surf.stickcoef = sint.stickcoef
    return ns(surfaceinteraction=surf,
              geometry=ns(startpoint=inputs.geometry.startpoint,
                          taa=u.Quantity(float(np.asarray(inputs.geometry.taa)), u.rad)),
              options=ns(species=inputs.options.species))


def surface_interaction(inputs, **kwargs):
    """reference SurfaceInteraction(inputs) for one of this repo's parsed inputfiles."""
    install()
    from nexoclom.particle_tracking.SurfaceInteraction import SurfaceInteraction
    return SurfaceInteraction(reference_inputs(inputs), **kwargs)


def planet_dist(planet, taa):
    """(r [au], v_r [km/s]) from the reference's planet_dist."""
    install()
    from nexoclom.solarsystem.planet_dist import planet_dist as ref_planet_dist
    r, v_r = ref_planet_dist(planet, taa=float(taa))
    return float(np.asarray(r)), float(np.asarray(v_r))


def ssobject(name):
    install()
    from nexoclom.solarsystem.SSObject import SSObject
    return SSObject(name)
