"""Surface-interaction tables built on the host (once per run).

Follows the reference ``particle_tracking/SurfaceInteraction.py:10-61``,
``initial_state/surface_temperature.py:4-19`` and ``math/distributions.py:7-21``:

* temperature-dependent sticking ``A0*exp(A1*T)+A2`` clipped to [0,1];
* thermal-accommodation speed table ``probgrid[T, prob]`` (201 x 101): for each
  surface temperature the Maxwellian-flux CDF on 101 speeds 0..3 v_th is
  inverted onto 101 probabilities; wrapped in a bicubic
  ``scipy.interpolate.RectBivariateSpline``.  The spline's knots/coefficients
  (``tck``) are exported so the CUDA kernel evaluates the very same spline with
  de Boor's recurrence (``nx_tables_upload``).
"""
import numpy as np
from scipy import interpolate

from .atomicdata import atomicmass, K_BOLTZMANN, AMU
from .units import value_of


def surface_temperature(geometry, longitude, latitude, t0=100., t1=None, n=.25):
    """Mercury surface temperature [K] at (longitude, latitude) [rad]
    (reference ``surface_temperature.py:4-19``).  Returns None for any other
    start point, like the reference."""
    if geometry.startpoint == 'Mercury':
        if t1 is None:
            t1 = 600. + 125 * (np.cos(value_of(geometry.taa)) - 1) / 2.
        t_surf = np.zeros_like(longitude) + t0
        mask = (longitude <= np.pi / 2) | (longitude >= 3 * np.pi / 2)
        t_surf[mask] = t0 + t1 * np.abs(np.cos(longitude[mask]) *
                                        np.cos(latitude[mask]))**n
        return t_surf


def thermal_speed_kms(temperature, species):
    """sqrt(2 k T / m) in km/s."""
    mass = atomicmass(species).value
    return np.sqrt(2 * temperature * K_BOLTZMANN / mass) * (np.sqrt(1. / AMU) / 1e3)


def MaxwellianDist(velocity, temperature, species):
    """Maxwellian FLUX distribution v^3 exp(-v^2/v_th^2), peak-normalised
    (reference ``math/distributions.py:16-21``); velocity in km/s."""
    mass = atomicmass(species).value
    vth2 = 2 * temperature * K_BOLTZMANN / mass * (1. / AMU / 1e6)
    f_v = velocity**3 * np.exp(-velocity**2 / vth2)
    f_v /= np.max(f_v)
    return f_v


def sputdist(velocity, U_eV, alpha, beta, species):
    """Sputtering speed distribution v^(2b+1)/(v^2+v_b^2)^a, peak-normalised
    (reference ``math/distributions.py:7-13``); velocity in km/s, U in eV."""
    from .atomicdata import EV_J
    mass = atomicmass(species).value
    v_b = np.sqrt(2 * U_eV / mass) * (np.sqrt(EV_J / AMU) / 1e3)
    f_v = velocity**(2 * beta + 1) / (velocity**2 + v_b**2)**alpha
    f_v /= np.max(f_v)
    return f_v


class TemperatureDependentSticking:
    """stickcoef(lon, lat) = clip(A0 exp(A1 T_surf) + A2, 0, 1) (reference
    SurfaceInteraction.py:15-20).  A class rather than the reference's closure so that an
    Output carrying it can be pickled."""

    def __init__(self, geometry, A):
        self.geometry = geometry
        self.A = A

    def __call__(self, lon, lat):
        tsurf = surface_temperature(self.geometry, lon, lat)
        coef = self.A[0] * np.exp(self.A[1] * tsurf) + self.A[2]
        coef[coef > 1.] = 1.
        coef[coef < 0.] = 0.
        return coef


class SurfaceInteraction:
    def __init__(self, inputs, **kwargs):
        sint = inputs.surfaceinteraction
        if sint.sticktype == 'temperature dependent':
            self.stickcoef = TemperatureDependentSticking(inputs.geometry, sint.A)
        elif sint.sticktype == 'surface map':
            assert 0

        self.tck = None
        if sint.accomfactor == 0:
            return
        longitude = np.arange(361) * np.pi / 180.
        latitude = np.arange(181) * np.pi / 180. - np.pi / 2.
        longrid, latgrid = np.meshgrid(longitude, latitude)
        tsurf = surface_temperature(inputs.geometry, longrid.flatten(), latgrid.flatten())

        nt = kwargs.get('nt', 201)
        nv = kwargs.get('nv', 101)
        nprob = kwargs.get('nprob', 101)
        species = inputs.options.species

        temperature = np.linspace(min(tsurf), max(tsurf), nt)
        v_temp = thermal_speed_kms(temperature, species)
        probability = np.linspace(0, 1, nprob)
        probgrid = np.ndarray((nt, nprob))
        for i, t in enumerate(temperature):
            vrange = np.linspace(0, v_temp[i] * 3, nv)
            f_v = MaxwellianDist(vrange, t, species)
            cumdist = f_v.cumsum()
            cumdist -= cumdist.min()
            cumdist /= cumdist.max()
            probgrid[i, :] = np.interp(probability, cumdist, vrange)

        spline = interpolate.RectBivariateSpline(temperature, probability, probgrid)
        self.v_interp = spline.ev
        tx, ty = spline.get_knots()
        self.tck = (np.ascontiguousarray(tx), np.ascontiguousarray(ty),
                    np.ascontiguousarray(spline.get_coeffs()))
        self.probgrid = probgrid
        self.temperature = temperature
        self.probability = probability
