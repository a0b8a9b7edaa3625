"""Oracle: source maps.

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

NumPy / scikit-learn restatement of reference
nexoclom/data_simulation/make_source_map.py:44-173 (the numerical body, astropy units
stripped), using the same library calls (np.histogram2d, np.histogram,
sklearn BallTree(metric='haversine').query_radius).

Pinned by: tests/golden/source_map.npz, produced by EXECUTING the unmodified reference
function (tools/make_golden_products.py); see tests/test_oracle_products_golden.py.
"""
import numpy as np


def make_source_map(X0, R_planet_km, params, todo):
    """X0: dict / DataFrame with longitude, latitude, v [R_p/s], altitude, azimuth, frac."""
    from sklearn.neighbors import BallTree
    smear_radius = params.get('smear_radius', np.radians(10))
    nlonbins = params.get('nlonbins', 180)
    nlatbins = params.get('nlatbins', 90)
    nvelbins = params.get('nvelbins', 100)
    nazbins = params.get('nazbins', 45)
    naltbins = params.get('naltbins', 23)
    cols = {k: np.asarray(X0[k], dtype=np.float64) for k in
            ('longitude', 'latitude', 'v', 'altitude', 'azimuth', 'frac')}
    vmax = np.ceil(cols['v'].max() * R_planet_km)
    included = cols['frac'] > 0
    weight = cols['frac'] if todo == 'source' else np.ones(len(included))

    H, xe, ye = np.histogram2d(cols['longitude'][included], cols['latitude'][included],
                               weights=weight[included],
                               range=[[0, 2 * np.pi], [-np.pi / 2, np.pi / 2]],
                               bins=(nlonbins, nlatbins))
    x = xe[:-1] + (xe[1] - xe[0]) / 2
    y = ye[:-1] + (ye[1] - ye[0]) / 2
    gridlatitude, gridlongitude = np.meshgrid(y, x)
    out = {'abundance_hist': H, 'longitude': x, 'latitude': y}

    def hist(a, bins, rng):
        h, e = np.histogram(a[included], bins=bins, range=rng, weights=weight[included])
        return h.astype(float), e[:-1] + (e[1] - e[0]) / 2
    out['speed_dist'], out['speed'] = hist(cols['v'] * R_planet_km, nvelbins, [0, vmax])
    out['altitude_dist'], out['altitude'] = hist(cols['altitude'], naltbins, [0, np.pi / 2])
    out['azimuth_dist'], out['azimuth'] = hist(cols['azimuth'], nazbins, [0, 2 * np.pi])

    points = np.array([gridlatitude.flatten(), gridlongitude.flatten()]).T
    tree = BallTree(np.stack([cols['latitude'], cols['longitude']], axis=1), metric='haversine')
    ind = tree.query_radius(points, smear_radius * np.cos(points[:, 0]))
    npts = points.shape[0]
    n_included, n_total, abundance = np.zeros(npts), np.zeros(npts), np.zeros(npts)
    v_point = np.zeros((npts, nvelbins))
    alt_point = np.zeros((npts, naltbins))
    az_point = np.zeros((npts, nazbins))
    for k in range(npts):
        sel = ind[k]
        if len(sel) > 0:
            inc = included[sel]
            w = weight[sel]
            n_included[k] = inc.sum()
            n_total[k] = len(sel)
            abundance[k] = w.sum()
            si = sel[inc]
            v_point[k], _ = np.histogram(cols['v'][si] * R_planet_km, bins=nvelbins,
                                         range=[0, vmax], weights=w[inc])
            alt_point[k], _ = np.histogram(cols['altitude'][si], bins=naltbins,
                                           range=[0, np.pi / 2], weights=w[inc])
            az_point[k], _ = np.histogram(cols['azimuth'][si], bins=nazbins,
                                          range=[0, 2 * np.pi], weights=w[inc])
    shp = gridlongitude.shape
    out.update(n_included=n_included.reshape(shp), n_total=n_total.reshape(shp),
               abundance=abundance.reshape(shp), speed_map=v_point.reshape(shp + (nvelbins,)),
               altitude_map=alt_point.reshape(shp + (naltbins,)),
               azimuth_map=az_point.reshape(shp + (nazbins,)))
    return out
