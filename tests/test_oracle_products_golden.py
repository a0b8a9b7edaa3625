"""The oracle's restatements of the PRODUCT stages -- initial-state transforms, image,
line-of-sight iteration -- against golden vectors produced by executing the unmodified
reference functions (tools/make_golden_products.py; reference
initial_state/source_distribution.py:37-283, data_simulation/ModelImage.py:229-274,
data_simulation/compute_iteration.py:90-240)."""
import os

import numpy as np
import pytest

from common import GOLDEN, source_case_input
from nexoclom_b200 import Input
from nexoclom_b200.runsetup import RunSetup
from oracle import initial_state, imaging

CASES = ['flat_iso', 'maxw_band', 'gauss_radial', 'sput_2d', 'spot_flat', 'lon1d_user']
COLS = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'v', 'longitude', 'latitude', 'local_time', 'altitude',
        'azimuth']
IDX = {'time': 0, 'x': 1, 'y': 2, 'z': 3, 'vx': 4, 'vy': 5, 'vz': 6, 'v': 8, 'longitude': 9, 'latitude': 10,
       'local_time': 11, 'altitude': 12, 'azimuth': 13}


def _replay(tag, g, device=None):
    """Map the reference's recorded draws (in its call order) onto the oracle's inputs
    (or, with ``device`` = the host build of the kernels' code, onto K1's transform)."""
    setup = RunSetup(source_case_input(tag))
    sp = setup.source_params(None)
    uni = list(g[f'{tag}_uniform'])
    legacy = list(g[f'{tag}_legacy'])
    normal = list(g[f'{tag}_normal'])
    n = len(g[f'{tag}_x'])
    u = {'time': uni.pop(0)}                       # Output.py:139
    lonlat = None
    if sp.spatial_type == 0:                       # source_distribution.py:52, 61
        u['sinlat'], u['lon'] = uni.pop(0), uni.pop(0)
    elif sp.spatial_type == 2:                     # random_deviates_1d: one legacy draw (:74)
        u['lon'] = legacy.pop(0)
    else:                                          # random_deviates_2d: pooled rounds of 3 draws
        fmap, xa, ya = setup.sourcemap
        rounds = [(legacy[k], legacy[k + 1], legacy[k + 2]) for k in range(0, len(legacy), 3)]
        legacy = []
        lonlat = initial_state.pooled_rejection(fmap, xa, ya, sp.map_fmax, rounds, n)
    if sp.speed_type == 0:                         # flat: randgen.random (:170)
        u['speed'] = uni.pop(0)
    elif sp.speed_type == 1:                       # gaussian: randgen.standard_normal (:145)
        u['normal'] = normal.pop(0)
    else:                                          # tabulated: module-level numpy.random.rand (Q12)
        u['speed'] = legacy.pop(0)
    if sp.angular_type == 1:                       # isotropic (:206, :212)
        u['alt'], u['az'] = uni.pop(0), uni.pop(0)
    elif sp.angular_type == 2:                     # 2d (:221)
        u['alt'] = uni.pop(0)
    assert not uni and not legacy and not normal, 'unconsumed reference draws'
    if device is not None:
        import ctypes as C
        from nexoclom_b200._lib import as_f64, dptr
        z = np.zeros(n)
        arr = {k: as_f64(u.get(k, z)) for k in ('time', 'sinlat', 'lon', 'speed', 'normal',
                                                 'alt', 'az')}
        ll = [as_f64(a) for a in lonlat] if lonlat is not None else None
        tab = getattr(setup, 'speed_table', None)
        cdf, vt = (as_f64(tab[0]), as_f64(tab[1])) if tab is not None else (None, None)
        out = np.zeros((n, 14))
        ltab = getattr(setup, 'lon_table', None)
        lcdf, lx = (as_f64(ltab[0]), as_f64(ltab[1])) if ltab is not None else (None, None)
        device.hc_init_from_deviates(
            C.c_long(n), C.byref(sp), dptr(cdf) if tab is not None else None,
            dptr(vt) if tab is not None else None, C.c_int(len(cdf) if tab is not None else 0),
            dptr(arr['time']), dptr(arr['sinlat']), dptr(arr['lon']),
            dptr(ll[0]) if ll else None, dptr(ll[1]) if ll else None, dptr(arr['speed']),
            dptr(arr['normal']), dptr(arr['alt']), dptr(arr['az']), dptr(out),
            dptr(lcdf) if ltab is not None else None, dptr(lx) if ltab is not None else None,
            C.c_int(len(lcdf) if ltab is not None else 0))
        return out
    return initial_state.transform(sp, u, getattr(setup, 'sourcemap', None),
                                   getattr(setup, 'speed_table', None), None, lonlat=lonlat,
                                   lon_table=getattr(setup, 'lon_table', None))


@pytest.mark.parametrize('tag', CASES)
def test_initial_state_transform_vs_reference(tag):
    g = np.load(os.path.join(GOLDEN, 'source_distribution.npz'))
    X0 = _replay(tag, g)
    for c in COLS:
        ref = g[f'{tag}_{c}']
        got = X0[:, IDX[c]]
        scale = np.maximum(np.abs(ref), 1e-300)
        err = np.max(np.abs(got - ref) / scale) if len(ref) else 0.0
        # identical operations on identical deviates: a few ulp at most (speed tables
        # go through unit conversions in the reference: km/s -> R_p/s)
        assert err < 5e-15, (tag, c, err)


IMAGE_CASES = ['col_pole', 'rad_pole', 'rad_side']


def _na_setup():
    from common import workload
    return RunSetup(workload('Na.maxwellian.radpres.input'))


@pytest.mark.parametrize('tag', IMAGE_CASES)
def test_create_image_vs_reference(tag):
    """oracle.imaging.create_image == the reference's ModelImage.create_image() (run
    unmodified, incl. its packet_weighting / interpu / Histogram2d) on the same packets."""
    g = np.load(os.path.join(GOLDEN, 'image.npz'))
    setup = _na_setup()
    assert float(g['vrplanet']) == setup.vrplanet
    X = g['X']
    view, dims = g[f'{tag}_view'], [int(d) for d in g[f'{tag}_dims']]
    M = imaging.image_rotation(*view)
    assert np.max(np.abs(np.asarray(M) - g[f'{tag}_M'])) < 1e-15
    quantity = 'column' if tag.startswith('col') else 'radiance'
    img, cnt, xe, ze = imaging.create_image(
        X[:, 1], X[:, 2], X[:, 3], X[:, 5], X[:, 7], vrplanet=setup.vrplanet, M=M, dims=dims,
        xrange=(-4, 4), zrange=(-4, 4), apix=float(g[f'{tag}_apix']), quantity=quantity,
        gtables=setup.gtables([5891, 5897]))
    assert np.array_equal(cnt, g[f'{tag}_packim'])            # bit-exact pixel indexing
    ref = g[f'{tag}_image']
    nz = ref > 0
    assert nz.sum() > 1000
    assert np.max(np.abs(img[nz] - ref[nz]) / ref[nz]) < 1e-12
    assert np.all(img[~nz] == 0)
    dx = xe[1] - xe[0]
    assert np.allclose(xe[:-1] + dx / 2, g[f'{tag}_xaxis'], rtol=0, atol=1e-15)


@pytest.mark.parametrize('tag', ['d1', 'd3'])
def test_los_iteration_vs_reference(tag):
    """oracle.imaging.los_iteration == the reference's compute_iteration() (run unmodified:
    KD-tree ladder, cone test, planet truncation, foot-point shadow, used / included sets)."""
    g = np.load(os.path.join(GOLDEN, 'los.npz'))
    setup = _na_setup()
    X = g['X']
    used = []
    rad, npk, inc, _ = imaging.los_iteration(
        X[:, 1], X[:, 2], X[:, 3], X[:, 5], X[:, 7], g['los'], vrplanet=setup.vrplanet,
        dphi=float(g[f'{tag}_dphi']), outeredge=float(g['outeredge']),
        rp_cm=setup.radius_km * 1e5, gtables=setup.gtables([5891, 5897]), used=used)
    assert np.array_equal(npk, g[f'{tag}_npackets'])          # bit-exact hit counts
    assert npk.sum() > 300
    assert np.array_equal(inc, g[f'{tag}_included'])
    ref = g[f'{tag}_radiance']
    nz = ref > 0
    assert np.max(np.abs(rad[nz] - ref[nz]) / ref[nz]) < 1e-12
    assert np.all(rad[~nz] == 0)
    off, idx = g[f'{tag}_used_off'], g[f'{tag}_used_idx']
    for i, s in enumerate(used):
        assert s == set(int(k) for k in idx[off[i]:off[i + 1]])


@pytest.fixture(scope='module')
def hc():
    import ctypes as C
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_hostcheck',
                        'libnexo_hostcheck.so')
    if not os.path.exists(path):
        import subprocess
        subprocess.run(['sh', os.path.join(os.path.dirname(path), 'build.sh')], check=True)
    return C.CDLL(path)


@pytest.mark.parametrize('tag', CASES)
def test_k1_transform_vs_reference(hc, tag):
    """The kernel's own initial-state transform (csrc/nx_init.cuh::init_packet_finish,
    compiled for the host) replayed on the reference's recorded deviates."""
    g = np.load(os.path.join(GOLDEN, 'source_distribution.npz'))
    X0 = _replay(tag, g, device=hc)
    for c in COLS:
        ref = g[f'{tag}_{c}']
        got = X0[:, IDX[c]]
        # libm vs NumPy SIMD sin/cos/asin differ by an ulp or two; vector components that
        # nearly cancel are judged against the size of the vector, not their own
        err = np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)
        assert err < 1e-14, (tag, c, err)


SRCMAP_KEYS = {'abundance_hist': None, 'speed_dist': 'speed_dist', 'altitude_dist': 'altitude_dist',
               'azimuth_dist': 'azimuth_dist', 'n_included': 'n_included', 'n_total': 'n_total',
               'speed_map': 'speed_dist_map', 'altitude_map': 'altitude_dist_map',
               'azimuth_map': 'azimuth_dist_map', 'longitude': 'longitude', 'latitude': 'latitude',
               'speed': 'speed', 'altitude': 'altitude', 'azimuth': 'azimuth'}


def check_source_map(result, g, todo, exact_counts=True, tol=1e-12):
    """Compare a source-map result (oracle or CUDA) with the reference's dictionaries."""
    for smear in ('smear', 'hist'):
        tag = f'{todo}_{smear}'
        ref_ab = g[f'{tag}_abundance_uncor']
        got_ab = result['abundance'] if smear == 'smear' else result['abundance_hist']
        assert np.max(np.abs(got_ab - ref_ab)) <= tol * max(np.max(np.abs(ref_ab)), 1.0)
        for key, refkey in SRCMAP_KEYS.items():
            if refkey is None:
                continue
            ref = g[f'{tag}_{refkey}']
            got = np.asarray(result[key], dtype=np.float64)
            assert got.shape == ref.shape, (key, got.shape, ref.shape)
            if key in ('n_included', 'n_total') and exact_counts:
                assert np.array_equal(got, ref), key            # ball membership is bit-exact
            else:
                assert np.max(np.abs(got - ref)) <= tol * max(np.max(np.abs(ref)), 1.0), key


def source_map_inputs(g):
    cols = ['longitude', 'latitude', 'v', 'altitude', 'azimuth', 'frac']
    X0 = {c: g['X0'][:, k] for k, c in enumerate(cols)}
    params = {k[len('param_'):]: (float(g[k]) if 'radius' in k else int(g[k]))
              for k in g.files if k.startswith('param_')}
    return X0, float(g['radius_km']), params


@pytest.mark.parametrize('todo', ['source', 'available'])
def test_source_map_vs_reference(todo):
    """oracle.source_map == the reference's make_source_map() (run unmodified)."""
    from oracle import source_map
    g = np.load(os.path.join(GOLDEN, 'source_map.npz'))
    X0, rkm, params = source_map_inputs(g)
    res = source_map.make_source_map(X0, rkm, params, todo)
    assert g[f'{todo}_smear_n_total'].sum() > 1e5
    check_source_map(res, g, todo)
