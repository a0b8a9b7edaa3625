"""``make_source_map()``: surface source maps from the initial states of a model run
(reference ``data_simulation/make_source_map.py:11-175``, same signature, same keys in the
returned dictionary).  The whole-planet histograms and the per-grid-point haversine-ball
gathers run in one CUDA kernel (K6, ``nx_source_map``); the host only prepares the bin
axes the way ``np.histogram2d`` defines them."""
import numpy as np

from ._lib import SourceMapParams
from .engine import get_engine
from .Output import Output
from .units import Quantity


def source_map_arrays(X0, R_planet_km, params, todo, device=0):
    """Numerical core on plain arrays: X0 needs the columns longitude, latitude, v,
    altitude, azimuth, frac.  Returns the dict of ndarrays of ``Engine.source_map`` plus
    the bin-centre axes."""
    params = params or {}
    sp = SourceMapParams()
    sp.smear_radius = float(params.get('smear_radius', np.radians(10)))
    sp.nlon = int(params.get('nlonbins', 180))
    sp.nlat = int(params.get('nlatbins', 90))
    sp.nvel = int(params.get('nvelbins', 100))
    sp.naz = int(params.get('nazbins', 45))
    sp.nalt = int(params.get('naltbins', 23))
    sp.weight_is_frac = 1 if todo == 'source' else 0
    v_kms = np.asarray(X0['v'], dtype=np.float64) * R_planet_km
    sp.vmax = float(np.ceil(np.asarray(X0['v'], dtype=np.float64).max() * R_planet_km))

    def centres(lo, hi, n):                       # math/histogram.py:34-39
        edges = np.linspace(lo, hi, n + 1)
        return edges[:-1] + (edges[1] - edges[0]) / 2
    lon_c = centres(0, 2 * np.pi, sp.nlon)
    lat_c = centres(-np.pi / 2, np.pi / 2, sp.nlat)
    radius = sp.smear_radius * np.cos(lat_c)      # make_source_map.py:115
    out = get_engine(device).source_map(
        sp, X0['longitude'], X0['latitude'], v_kms, X0['altitude'], X0['azimuth'], X0['frac'],
        lon_c, lat_c, radius)
    out['longitude'], out['latitude'] = lon_c, lat_c
    out['speed'] = centres(0, sp.vmax, sp.nvel)
    out['altitude'] = centres(0, np.pi / 2, sp.nalt)
    out['azimuth'] = centres(0, 2 * np.pi, sp.naz)
    return out


def make_source_map(outputfile, params, todo=None, device=0):
    if todo == 'source':
        print('Determining modeled source')
    elif todo == 'available':
        print('Determining available source')
    else:
        return None
    params = params or {}
    smear_abundance = params.get('smear_abundance', True)
    print(outputfile)
    output = Output.restore(outputfile)
    X0 = output.X0
    R_planet_km = float(output.inputs.geometry.planet.radius.to('km').value)
    del output

    r = source_map_arrays(X0, R_planet_km, params, todo, device=device)
    distribution = {
        'abundance_uncor': r['abundance'] if smear_abundance else r['abundance_hist'],
        'longitude': Quantity(r['longitude'], 'rad'),
        'latitude': Quantity(r['latitude'], 'rad'),
        'speed_dist': r['speed_dist'], 'speed': Quantity(r['speed'], 'km/s'),
        'altitude_dist': r['altitude_dist'], 'altitude': Quantity(r['altitude'], 'rad'),
        'azimuth_dist': r['azimuth_dist'], 'azimuth': Quantity(r['azimuth'], 'rad'),
        'n_included': r['n_included'].astype(np.float64),
        'n_total': r['n_total'].astype(np.float64),
        'speed_dist_map': r['speed_map'], 'altitude_dist_map': r['altitude_map'],
        'azimuth_dist_map': r['azimuth_map'],
    }
    return distribution
