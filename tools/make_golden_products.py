#!/usr/bin/env python
"""Golden vectors for the PRODUCTS of the hot path, made by EXECUTING the unmodified
reference code (build container only; needs /root/reference):

  tests/golden/source_distribution.npz   reference surface_distribution(),
        speed_distribution(), angular_distribution() (source_distribution.py:37-283)
        with every random draw they make RECORDED, so that the oracle's pure transform
        (oracle/initial_state.transform) and the device transform (tests/_hostcheck,
        K1) can be replayed on exactly the same deviates;
  tests/golden/image.npz                 reference ModelImage.create_image()
        (ModelImage.py:229-274) incl. ModelResult.packet_weighting(), interpu(),
        Histogram2d();
  tests/golden/los.npz                   reference compute_iteration()
        (compute_iteration.py:90-240) incl. the KD-tree candidate ladder, cone test,
        planet truncation, foot-point shadow test and the used / included sets;
  tests/golden/losfit.npz                reference LOSResultFitted.determine_source_from_data()
        (LOSResultFitted.py:66-262) on the packets / lines of sight / used sets of los.npz
        (run `los` first), for use_weight in (None, 'dist', 'dist2', 'sigma');
  tests/golden/losresult_source_map.npz  reference LOSResult.make_source_map()
        (LOSResult.py:310-491) over two output files, normalised and raw.

astropy / periodictable / sqlalchemy are absent here: tools/refunits.py supplies a
functional miniature of astropy.units, the g-value tables come from this repo's host
port (nexoclom_b200.atomicdata.gValue, itself pinned to the reference's own golden
pickle to 0 ulp -- tests/test_host_tables.py), everything else on the path is the
reference's own source, imported from where it lies.
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
REF = os.environ.get('NEXOCLOM_REFERENCE', '/root/reference')
GOLD = os.path.join(REPO, 'tests', 'golden')

import refunits as u                                    # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Namespace stub for ``nexoclom`` (its __init__ needs PostgreSQL) + stand-ins for the
    absent third-party packages; the reference sub-packages are imported for real."""
    u.install()
    _stub('astropy.convolution', Gaussian2DKernel=None, convolve=None)
    _stub('astropy.time', Time=None)
    _stub('astropy.modeling', models=None, fitting=None)
    _stub('astropy.visualization', PercentileInterval=None)
    sa = _stub('sqlalchemy')
    sa.__path__ = []
    _stub('sqlalchemy.dialects').__path__ = []
    _stub('sqlalchemy.dialects.postgresql')
    masses = {'Na': 22.98977, 'Ca': 40.078, 'Mg': 24.305, 'K': 39.0983, 'O': 15.9994}
    els = [types.SimpleNamespace(symbol=k, mass=v) for k, v in masses.items()]
    _stub('periodictable', elements=els, **{e.symbol: e for e in els})
    u.u = u.Unit(1.66053906660e-27, (0, 0, 1, 0, 0), 'u')
    bk = _stub('bokeh')
    bk.__path__ = []
    for sub in ('plotting', 'palettes', 'models', 'io', 'themes'):
        _stub('bokeh.' + sub, Inferno256=None, HoverTool=None, ColumnDataSource=None,
              ColorBar=None, LogColorMapper=None, LogTicker=None, LinearColorMapper=None,
              curdoc=None, export_png=None, Theme=None)
    pkg = _stub('nexoclom', engine=None, config=None)
    pkg.__path__ = [os.path.join(REF, 'nexoclom')]
    pkg.__file__ = os.path.join(REF, 'nexoclom', '__init__.py')
    for sub in ('particle_tracking', 'initial_state', 'data_simulation', 'atomicdata',
                'utilities', 'solarsystem'):
        m = _stub(f'nexoclom.{sub}')
        m.__path__ = [os.path.join(REF, 'nexoclom', sub)]
    from nexoclom.utilities.exceptions import InputError
    sys.modules['nexoclom.utilities'].InputError = InputError
    # g-values: this repo's host port, wrapped in the reference's Quantity interface
    from nexoclom_b200 import atomicdata as ad

    def gValue(species, wavelength, aplanet):
        g = ad.gValue(species, float(np.asarray(wavelength)), float(np.asarray(aplanet)))
        return types.SimpleNamespace(velocity=u.Quantity(g.velocity, u.km / u.s),
                                     g=u.Quantity(g.g, 1 / u.s))
    sys.modules['nexoclom.atomicdata'].gValue = gValue
    sys.modules['nexoclom.atomicdata'].RadPresConst = None
    from nexoclom.atomicdata.atomicmass import atomicmass
    sys.modules['nexoclom.atomicdata'].atomicmass = atomicmass
    sys.modules['nexoclom.solarsystem'].planet_dist = None
    sys.modules['nexoclom.solarsystem'].SSObject = None
    _stub('nexoclom.initial_state.satellite_initial_positions', satellite_initial_positions=None)
    _stub('nexoclom.initial_state.LossInfo', LossInfo=None)
    _stub('nexoclom.initial_state.SourceMap', SourceMap=None)
    _stub('nexoclom.particle_tracking.SurfaceInteraction', SurfaceInteraction=None)
    _stub('nexoclom.initial_state.input_classes', InputError=InputError)
    _stub('nexoclom.math.smooth', smooth=None, smooth2d=None)


def write_source_case_tables(cdir):
    """The two tables of the `lon1d_user` source case, as unit-free pickled dicts."""
    import pickle
    lon = np.linspace(0, 2 * np.pi, 73)
    with open(os.path.join(cdir, 'lon1d_map.pkl'), 'wb') as f:
        pickle.dump({'longitude': lon, 'abundance': 1.0 + 0.8 * np.cos(lon - 1.0) ** 2,
                     'latitude': None, 'coordinate_system': 'solar-fixed'}, f, protocol=4)
    v = np.linspace(0.2, 8.0, 400)
    with open(os.path.join(cdir, 'user_speed.pkl'), 'wb') as f:
        pickle.dump({'speed': v, 'speed_dist': v ** 2 * np.exp(-v / 1.5)}, f, protocol=4)


class RecordingRNG:
    """numpy Generator that keeps every deviate it hands out, in call order."""

    def __init__(self, seed):
        self.g = np.random.default_rng(seed)
        self.uniform, self.normal = [], []

    def random(self, n):
        r = self.g.random(n)
        self.uniform.append(r)
        return r

    def standard_normal(self, n):
        r = self.g.standard_normal(n)
        self.normal.append(r)
        return r


def q(v, unit):
    return u.Quantity(v, unit)


# ---------------------------------------------------------------------------
def golden_source_distribution():
    import nexoclom.math.randomdeviates as rd
    from nexoclom.initial_state import source_distribution as sd
    ns = types.SimpleNamespace
    rp_km = 2440.53
    unit = u.def_unit('R_Mercury', q(rp_km, u.km))
    # the cases are inputfiles (tests/golden/source_cases/*.input) parsed by this repo's
    # Input class; the reference side receives the same numbers as astropy-style Quantities
    from nexoclom_b200 import Input
    from nexoclom_b200.units import Quantity as PQ
    umap = {'rad': u.rad, 'km/s': u.km / u.s, 'K': u.K, 'eV': u.eV, '': u.dimensionless}

    def conv(v):
        if isinstance(v, PQ):
            return q(float(v.value), umap[v.unit])
        if isinstance(v, tuple):
            return tuple(conv(x) for x in v)
        return v

    def group(obj):
        return ns(**{k: conv(v) for k, v in vars(obj).items()})
    cases = {}
    cdir = os.path.join(GOLD, 'source_cases')
    write_source_case_tables(cdir)
    from common import source_case_input
    from nexoclom_b200.sourcemap import load_pickle

    def RefSourceMap(filename):
        """What the reference's SourceMap(filename) holds for a pickled dict (SourceMap.py:
        19-30, 73-85): the dict's entries as astropy Quantities.  The committed pickles are
        unit-free so that the product reads them without astropy."""
        d = load_pickle(filename)
        units = {'longitude': u.rad, 'latitude': u.rad, 'speed': u.km / u.s,
                 'abundance': u.dimensionless, 'speed_dist': u.dimensionless}
        m = ns(coordinate_system=d.get('coordinate_system', 'solar-fixed'))
        for key, unit_ in units.items():
            v = d.get(key, None)
            setattr(m, key, None if v is None else q(np.asarray(v, dtype=float), unit_))
        return m
    sd.SourceMap = RefSourceMap
    for fn in sorted(f for f in os.listdir(cdir) if f.endswith('.input')):
        inp = source_case_input(fn[:-len('.input')])
        cases[fn[:-len('.input')]] = dict(spatial=group(inp.spatialdist),
                                          speed=group(inp.speeddist),
                                          angular=group(inp.angulardist),
                                          species=inp.options.species,
                                          endtime=float(inp.options.endtime.value))
    out = {}
    n = 4000
    for tag, c in cases.items():
        rng = RecordingRNG(77)
        legacy = []                       # draws of the module-level numpy.random.rand (Q12)
        glob = np.random.RandomState(1234)

        def rand(k, _l=legacy, _g=glob):
            r = _g.rand(k)
            _l.append(r)
            return r
        rd.random = types.SimpleNamespace(rand=rand)
        # Output.py:136-147: the time draw comes first; 'time' and 'frac' are the first two
        # columns, which angular_distribution() relies on (it reads x,y,z by POSITION 2:5)
        X0 = pd.DataFrame()
        X0['time'] = rng.random(n) * c['endtime']
        X0['frac'] = np.ones(n)
        outputs = ns(npackets=n, randgen=rng, unit=unit, X0=X0,
                     inputs=ns(spatialdist=c['spatial'], speeddist=c['speed'],
                               angulardist=c['angular'],
                               options=ns(species=c['species']),
                               geometry=ns(planet=ns(type='Planet'))))
        sd.surface_distribution(outputs)
        sd.speed_distribution(outputs)
        sd.angular_distribution(outputs)
        cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'v', 'longitude', 'latitude', 'local_time',
                'altitude', 'azimuth']
        for cname in cols:
            out[f'{tag}_{cname}'] = np.asarray(outputs.X0[cname].values, dtype=np.float64)
        out[f'{tag}_uniform'] = (np.stack(rng.uniform) if rng.uniform else np.zeros((0, n)))
        out[f'{tag}_normal'] = (np.stack(rng.normal) if rng.normal else np.zeros((0, n)))
        out[f'{tag}_legacy'] = (np.stack(legacy) if legacy else np.zeros((0, n)))
        print(tag, 'uniform draws', len(rng.uniform), 'normal', len(rng.normal),
              'legacy rand', len(legacy))
    np.savez_compressed(os.path.join(GOLD, 'source_distribution.npz'), **out)
    print('source_distribution.npz')


# ---------------------------------------------------------------------------
def _final_packets(n, seed, f32):
    """A plausible cloud of packets (no GPU here): oracle initial state pushed outward."""
    from common import workload
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    rng = np.random.default_rng(seed)
    X = initial_state.draw_x0(setup, n, seed)[:, :8].copy()
    X[:, 1:4] *= (1.0 + 5.0 * rng.random(n) ** 2)[:, None]
    X[:, 4:7] *= rng.normal(1.0, 0.5, n)[:, None]
    X[:, 7] = rng.random(n) * (rng.random(n) > 0.2)
    if f32:
        X = X.astype(np.float32).astype(np.float64)      # Output.save / restore (Q14)
    return setup, X


def golden_image():
    from nexoclom.data_simulation.ModelImage import ModelImage
    from nexoclom.data_simulation import ModelImage as mi_mod
    ns = types.SimpleNamespace
    setup, X = _final_packets(60000, 5, True)
    rp_km = setup.radius_km
    unit = u.def_unit('R_Mercury', q(rp_km, u.km))
    out = {'X': X, 'vrplanet': setup.vrplanet, 'aplanet': setup.aplanet}
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']
    for tag, quantity, view, dims in (('col_pole', 'column', (0.0, np.pi / 2), (200, 200)),
                                      ('rad_pole', 'radiance', (0.0, np.pi / 2), (200, 200)),
                                      ('rad_side', 'radiance', (0.7, 0.3), (160, 120))):
        packets = pd.DataFrame(X, columns=cols)
        fake_out = ns(X=packets, vrplanet=q(setup.vrplanet, unit / u.s), aplanet=setup.aplanet,
                      filename='none')
        mi_mod.Output = ns(restore=lambda fname, _o=fake_out: _o)
        planet = ns(object='Mercury', radius=q(rp_km, u.km))
        width = (q(8., unit), q(8., unit))
        center = (q(0., unit), q(0., unit))
        self = ns(inputs=ns(geometry=ns(planet=planet), options=ns(species='Na')),
                  origin=planet, unit=unit, quantity=quantity, g=None,
                  mechanism=['resonant scattering'] if quantity == 'radiance' else None,
                  wavelength=(q(5891, u.AA), q(5897, u.AA)) if quantity == 'radiance' else None,
                  subobslongitude=q(view[0], u.rad), subobslatitude=q(view[1], u.rad),
                  dims=dims,
                  xrange=[center[0] - width[0] / 2, center[0] + width[0] / 2],
                  zrange=[center[1] - width[1] / 2, center[1] + width[1] / 2])
        scale = (width[0] / dims[0], width[1] / dims[1])
        self.Apix = (scale[0] * scale[1]).to(u.cm ** 2)              # ModelImage.py:76-77
        self.image_rotation = types.MethodType(ModelImage.image_rotation, self)
        from nexoclom.data_simulation.ModelResult import ModelResult
        self.packet_weighting = types.MethodType(ModelResult.packet_weighting, self)
        self.save = lambda *a, **k: None
        image, packim = ModelImage.create_image(self, 'none')
        out[f'{tag}_image'] = np.asarray(image.histogram, dtype=np.float64)
        out[f'{tag}_packim'] = np.asarray(packim.histogram, dtype=np.float64)
        out[f'{tag}_xaxis'] = np.asarray(image.x, dtype=np.float64)
        out[f'{tag}_zaxis'] = np.asarray(image.y, dtype=np.float64)
        out[f'{tag}_M'] = np.asarray(self.image_rotation(), dtype=np.float64)
        out[f'{tag}_apix'] = float(np.asarray(self.Apix))
        out[f'{tag}_view'] = np.asarray(view)
        out[f'{tag}_dims'] = np.asarray(dims)
        print(tag, 'sum', float(image.histogram.sum()), 'packets in frame',
              int(packim.histogram.sum()))
    np.savez_compressed(os.path.join(GOLD, 'image.npz'), **out)
    print('image.npz')


def golden_los():
    from nexoclom.data_simulation import compute_iteration as ci
    from nexoclom.data_simulation.ModelResult import ModelResult
    ns = types.SimpleNamespace
    setup, X = _final_packets(40000, 8, True)
    rp_km = setup.radius_km
    unit = u.def_unit('R_Mercury', q(rp_km, u.km))
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']
    rng = np.random.default_rng(3)
    nlos = 120
    th = rng.random(nlos) * 2 * np.pi
    rr = 1.15 + 3.0 * rng.random(nlos)
    x_sc = np.stack([0.3 * rr * np.cos(th), 0.2 * rr * np.cos(th) - 0.6, rr * np.sin(th)], axis=1)
    x_sc *= (np.maximum(np.linalg.norm(x_sc, axis=1), 1.1) / np.linalg.norm(x_sc, axis=1))[:, None]
    tgt = rng.normal(size=(nlos, 3))
    tgt *= ((0.5 + 3 * rng.random(nlos)) / np.linalg.norm(tgt, axis=1))[:, None]
    tgt[::3] *= 0.2                                   # every third boresight hits the planet
    bore = tgt - x_sc
    bore /= np.linalg.norm(bore, axis=1)[:, None]
    data = pd.DataFrame({'x': x_sc[:, 0], 'y': x_sc[:, 1], 'z': x_sc[:, 2],
                         'xbore': bore[:, 0], 'ybore': bore[:, 1], 'zbore': bore[:, 2]})
    out = {'X': X, 'vrplanet': setup.vrplanet, 'aplanet': setup.aplanet,
           'los': np.concatenate([x_sc, bore], axis=1), 'outeredge': 25.0}
    captured = {}

    class Capture:
        def __init__(self, iteration, losresult):
            captured.update(iteration)

        def save_iteration(self):
            pass
    ci.IterationResult = Capture
    for tag, dphi_deg in (('d1', 1.0), ('d3', 3.0)):
        packets = pd.DataFrame(X, columns=cols)
        fake_out = ns(X=packets, X0=pd.DataFrame(index=packets.index),
                      vrplanet=q(setup.vrplanet, unit / u.s), aplanet=setup.aplanet,
                      totalsource=float(len(X)), idnum=1, unit=unit)
        ci.Output = ns(restore=lambda fname, _o=fake_out: _o)
        self = ns(inputs=ns(options=ns(outeredge=25.0, species='Na')), unit=unit,
                  quantity='radiance', g=None, mechanism=['resonant scattering'],
                  wavelength=(q(5891, u.AA), q(5897, u.AA)), dphi=np.radians(dphi_deg),
                  query='golden', fitted=False)
        self.packet_weighting = types.MethodType(ModelResult.packet_weighting, self)
        scdata = ns(data=data.copy(), query='golden')
        ci.compute_iteration(self, 'none', scdata)
        out[f'{tag}_radiance'] = np.asarray(captured['radiance'].values, dtype=np.float64)
        out[f'{tag}_npackets'] = np.asarray(captured['npackets'].values, dtype=np.int64)
        out[f'{tag}_included'] = np.asarray(captured['included'].values, dtype=bool)
        used = captured['used0']
        off = np.zeros(nlos + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s_) for s_ in used])
        out[f'{tag}_used_off'] = off
        out[f'{tag}_used_idx'] = np.concatenate(
            [np.sort(np.fromiter(s_, dtype=np.int64, count=len(s_))) for s_ in used]
            + [np.zeros(0, dtype=np.int64)])
        out[f'{tag}_dphi'] = np.radians(dphi_deg)
        print(tag, 'hits', int(out[f'{tag}_npackets'].sum()), 'lines with signal',
              int((out[f'{tag}_radiance'] > 0).sum()), 'included', int(out[f'{tag}_included'].sum()))
    np.savez_compressed(os.path.join(GOLD, 'los.npz'), **out)
    print('los.npz')


def golden_losfit():
    """Execute the UNMODIFIED LOSResultFitted.determine_source_from_data (reference
    LOSResultFitted.py:66-262) on the packets / lines of sight / `used` sets of los.npz:
    PostgreSQL search -> "no saved result", Output.restore / the unfitted iteration pickle /
    IterationResultFitted replaced by in-memory stand-ins that capture what the method
    computes."""
    import pickle
    import tempfile
    from nexoclom.data_simulation import LOSResultFitted as lf
    from nexoclom.data_simulation.ModelResult import ModelResult
    ns = types.SimpleNamespace
    g = np.load(os.path.join(GOLD, 'los.npz'))
    X = g['X']
    n = len(X)
    los = g['los']
    nlos = len(los)
    setup, _ = _final_packets(8, 8, True)
    unit = u.def_unit('R_Mercury', q(setup.radius_km, u.km))
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']
    off, idx = g['d3_used_off'], g['d3_used_idx']
    used = pd.Series([set(int(k) for k in idx[off[i]:off[i + 1]]) for i in range(nlos)])
    model_rad = g['d3_radiance'] * 3.0e-7                 # any positive scale: a ratio enters
    rng = np.random.default_rng(12)
    truth = model_rad * (1.0 + 0.8 * np.sin(np.arange(nlos) * 0.37)) + 0.02 * rng.random(nlos)
    sigma = 0.05 + 0.1 * rng.random(nlos)
    mask = (model_rad > 0) & (rng.random(nlos) > 0.15)
    data = pd.DataFrame({'x': los[:, 0], 'y': los[:, 1], 'z': los[:, 2], 'xbore': los[:, 3],
                         'ybore': los[:, 4], 'zbore': los[:, 5], 'radiance': truth,
                         'sigma': sigma, 'mask_unfit': mask})
    out = {'model_radiance': model_rad, 'data_radiance': truth, 'sigma': sigma, 'mask': mask,
           'dphi': g['d3_dphi'], 'endtime': 50000.0}
    tmp = tempfile.mkdtemp()
    modelfile = os.path.join(tmp, 'unfit_iteration.pkl')
    with open(modelfile, 'wb') as f:
        pickle.dump(ns(used_packets=used), f)
    captured = {}

    class Capture:
        def __init__(self, iteration, losresult):
            captured.update(iteration)
            self.radiance = pd.Series(iteration['radiance'])
            self.totalsource = iteration['totalsource']
            self.outputfile = iteration['outputfile']
            self.modelfile = 'fitted_iteration.pkl'

        def save_iteration(self):
            pass
    lf.IterationResultFitted = Capture
    for tag, use_weight in (('none', None), ('dist', 'dist'), ('dist2', 'dist2'),
                            ('sigma', 'sigma')):
        packets = pd.DataFrame(X, columns=cols)
        saved = {}
        fake_out = ns(X=packets, X0=pd.DataFrame({'frac': np.ones(n)}), npackets=n, nsteps=1,
                      vrplanet=q(setup.vrplanet, unit / u.s), aplanet=setup.aplanet,
                      totalsource=float(n), idnum=2, filename='fitted_output.pkl', unit=unit,
                      inputs=None)
        fake_out.save = lambda _o=fake_out: saved.update(frac=_o.X['frac'].values.copy(),
                                                         frac0=_o.X0['frac'].values.copy(),
                                                         totalsource=float(_o.totalsource))
        lf.Output = ns(restore=lambda fname, _o=fake_out: _o)
        unfit = ns(outid=[1], outputfiles=['unfit_output.pkl'],
                   modelfiles={'unfit_output.pkl': modelfile}, radiance=pd.Series(model_rad))
        self = ns(unfitted_label='unfit', unit=unit, quantity='radiance', g=None,
                  mechanism=['resonant scattering'], dphi=float(g['d3_dphi']), query='golden',
                  wavelength=(q(5891, u.AA), q(5897, u.AA)), radiance=pd.Series(np.zeros(nlos)),
                  totalsource=0., fitted=True,
                  inputs=ns(delete_files=lambda: None,
                            options=ns(endtime=q(50000., u.s), species='Na', outeredge=25.0)),
                  fitted_iteration_search=lambda ufit_id: None)
        self.packet_weighting = types.MethodType(ModelResult.packet_weighting, self)
        self.determine_source_rate = lambda scdata, use_weight=False: setattr(
            self, 'sourcerate', q(1.0, 1 / u.s))
        scdata = ns(data=data.copy(), model_result={'unfit': unfit}, query='golden')
        captured.clear()
        try:
            lf.LOSResultFitted.determine_source_from_data(self, scdata, use_weight=use_weight)
        except Exception as exc:                      # unit bookkeeping of the epilogue (:253-259)
            print('  epilogue stopped at:', type(exc).__name__, exc)
        assert 'radiance' in captured and saved, 'the per-file body did not complete'
        out[f'{tag}_radiance'] = np.asarray(captured['radiance'], dtype=np.float64)
        out[f'{tag}_frac'] = saved['frac']
        out[f'{tag}_frac0'] = saved['frac0']
        out[f'{tag}_totalsource'] = saved['totalsource']
        print(tag, 'fitted radiance sum', float(out[f'{tag}_radiance'].sum()), 'packets used',
              int((saved['frac0'] > 0).sum()))
    # use_selected=True (:95-113): a constant-step-like output (several rows per packet on a
    # common time grid), one random step per trajectory drawn with the output's own generator
    from oracle import imaging
    n0, nst = 4000, 4
    _, R = _final_packets(n0 * nst, 21, True)
    R[:, 0] = np.tile(np.array([90., 60., 30., 0.]), n0)
    rows = pd.DataFrame(R, columns=cols)
    rows['Index'] = np.repeat(np.arange(n0), nst)
    keep = np.random.default_rng(4).random(len(rows)) > 0.25          # packets die on the way
    rows = rows[keep]
    gt = setup.gtables([5891, 5897])
    used_rows = []
    rad_sel, _, _, _ = imaging.los_iteration(
        rows.x.values, rows.y.values, rows.z.values, rows.vy.values, rows.frac.values, los,
        vrplanet=setup.vrplanet, dphi=float(g['d3_dphi']), outeredge=25.,
        rp_cm=setup.radius_km * 1e5, gtables=gt, used=used_rows)
    labels = rows.index.values
    used_s = pd.Series([set(int(labels[k]) for k in u_) for u_ in used_rows])
    model_s = rad_sel * 3.0e-7
    truth_s = model_s * (1.0 + 0.8 * np.cos(np.arange(nlos) * 0.21)) + 1e-22
    mask_s = model_s > 0
    data_s = data.copy()
    data_s['radiance'] = truth_s
    data_s['mask_unfit'] = mask_s
    with open(modelfile, 'wb') as f:
        pickle.dump(ns(used_packets=used_s), f)
    saved = {}
    fake_out = ns(X=rows.copy(), X0=pd.DataFrame({'frac': np.ones(n0)}), npackets=n0, nsteps=nst,
                  vrplanet=q(setup.vrplanet, unit / u.s), aplanet=setup.aplanet,
                  totalsource=float(n0), idnum=3, filename='fitted_output.pkl', unit=unit,
                  inputs=None, randgen=np.random.default_rng(5))
    fake_out.save = lambda _o=fake_out: saved.update(
        labels=_o.X.index.values.copy(), frac=_o.X['frac'].values.copy(),
        frac0=_o.X0['frac'].values.copy(), totalsource=float(_o.totalsource))
    lf.Output = ns(restore=lambda fname, _o=fake_out: _o)
    unfit = ns(outid=[1], outputfiles=['unfit_output.pkl'],
               modelfiles={'unfit_output.pkl': modelfile}, radiance=pd.Series(model_s))
    self = ns(unfitted_label='unfit', unit=unit, quantity='radiance', g=None,
              mechanism=['resonant scattering'], dphi=float(g['d3_dphi']), query='golden',
              wavelength=(q(5891, u.AA), q(5897, u.AA)), radiance=pd.Series(np.zeros(nlos)),
              totalsource=0., fitted=True,
              inputs=ns(delete_files=lambda: None,
                        options=ns(endtime=q(50000., u.s), species='Na', outeredge=25.0)),
              fitted_iteration_search=lambda ufit_id: None)
    self.packet_weighting = types.MethodType(ModelResult.packet_weighting, self)
    self.determine_source_rate = lambda scdata, use_weight=False: setattr(
        self, 'sourcerate', q(1.0, 1 / u.s))
    captured.clear()
    lf.LOSResultFitted.determine_source_from_data(
        self, ns(data=data_s, model_result={'unfit': unfit}, query='golden'),
        use_selected=True, use_weight='dist2')
    order = np.argsort(saved['labels'])
    out.update(sel_rows=rows[cols + ['Index']].values, sel_row_labels=labels,
               sel_used_off=np.concatenate([[0], np.cumsum([len(u_) for u_ in used_rows])]),
               sel_used_idx=np.concatenate([np.sort(labels[np.asarray(u_, dtype=np.int64)])
                                            if len(u_) else np.zeros(0, dtype=np.int64)
                                            for u_ in (list(x) for x in used_rows)]),
               sel_model_radiance=model_s, sel_data_radiance=truth_s, sel_mask=mask_s,
               sel_kept_labels=saved['labels'][order], sel_frac=saved['frac'][order],
               sel_frac0=saved['frac0'], sel_totalsource=saved['totalsource'],
               sel_radiance=np.asarray(captured['radiance'], dtype=np.float64),
               sel_npackets0=n0, sel_seed=5)
    print('selected: rows kept', len(saved['labels']), 'of', len(rows), 'fitted radiance sum',
          float(out['sel_radiance'].sum()), 'lines with signal', int((out['sel_radiance'] > 0).sum()))
    np.savez_compressed(os.path.join(GOLD, 'losfit.npz'), **out)
    print('losfit.npz')


def golden_source_map():
    """reference data_simulation/make_source_map.py, unmodified, on oracle-drawn X0."""
    from nexoclom.data_simulation import make_source_map as msm
    from nexoclom_b200 import Input
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state
    ns = types.SimpleNamespace
    setup = RunSetup(Input(os.path.join(GOLD, 'source_cases', 'maxw_band.input')))
    n = 30000
    X0a = initial_state.draw_x0(setup, n, 17)
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac', 'v', 'longitude', 'latitude',
            'local_time', 'altitude', 'azimuth']
    X0 = pd.DataFrame(X0a, columns=cols)
    rng = np.random.default_rng(4)
    X0['frac'] = rng.random(n) * (rng.random(n) > 0.3)          # some packets not seen (frac = 0)
    X0 = X0.astype(np.float32).astype(np.float64)                # Output.save / restore (Q14)
    rp_km = setup.radius_km
    fake = ns(X0=X0, inputs=ns(geometry=ns(planet=ns(radius=q(rp_km, u.km)))))
    msm.Output = ns(restore=lambda fname, _o=fake: _o)
    params = {'smear_radius': np.radians(12.), 'nlonbins': 36, 'nlatbins': 18, 'nvelbins': 25,
              'nazbins': 12, 'naltbins': 9}
    out = {'X0': X0[['longitude', 'latitude', 'v', 'altitude', 'azimuth', 'frac']].values,
           'radius_km': rp_km}
    for k, v in params.items():
        out['param_' + k] = v
    for todo in ('source', 'available'):
        for smear in (True, False):
            d = msm.make_source_map('none', dict(params, smear_abundance=smear), todo=todo)
            tag = f"{todo}_{'smear' if smear else 'hist'}"
            for key, val in d.items():
                out[f'{tag}_{key}'] = np.asarray(val, dtype=np.float64)
            print(tag, 'n_total', d['n_total'].sum(), 'n_included', d['n_included'].sum())
    np.savez_compressed(os.path.join(GOLD, 'source_map.npz'), **out)
    print('source_map.npz')


def golden_losresult_source_map():
    """reference LOSResult.make_source_map (LOSResult.py:310-491), unmodified, summing the
    reference's own make_source_map() of two output files with different speed ranges
    (the np.interp branch) and converting to fluxes with astropy-style units."""
    from nexoclom.data_simulation import LOSResult as lr
    from nexoclom.data_simulation import make_source_map as msm
    from nexoclom_b200 import Input
    from nexoclom_b200.runsetup import RunSetup
    from oracle import initial_state
    ns = types.SimpleNamespace
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac', 'v', 'longitude', 'latitude',
            'local_time', 'altitude', 'azimuth']
    keep = ['longitude', 'latitude', 'v', 'altitude', 'azimuth', 'frac']
    files, out = {}, {}
    setup = RunSetup(Input(os.path.join(GOLD, 'source_cases', 'maxw_band.input')))
    for name, seed, n in (('f1', 17, 20000), ('f2', 23, 12000)):
        X0 = pd.DataFrame(initial_state.draw_x0(setup, n, seed), columns=cols)
        rng = np.random.default_rng(seed)
        X0['frac'] = rng.random(n) * (rng.random(n) > 0.3)
        if name == 'f2':
            X0['v'] *= 0.8                  # smaller speed range -> the np.interp branch (:361-369)
        X0 = X0.astype(np.float32).astype(np.float64)
        files[name] = ns(X0=X0, inputs=ns(geometry=ns(planet=ns(radius=q(setup.radius_km, u.km)))))
        out[f'{name}_X0'] = X0[keep].values
    msm.Output = ns(restore=lambda fname: files[fname])
    lr.SourceMap = lambda d: ns(**d)
    params = {'smear_radius': np.radians(12.), 'nlonbins': 24, 'nlatbins': 12, 'nvelbins': 15,
              'nazbins': 8, 'naltbins': 6}
    for k, v in params.items():
        out['param_' + k] = v
    out['radius_km'] = setup.radius_km
    out['sourcerate_1e23'] = 2.5
    rate_unit = u.def_unit('10**23 atoms/s', 1e23 / u.s)
    for tag, normalize in (('norm', True), ('raw', False)):
        self = ns(modelfiles={'f1': 'm1', 'f2': 'm2'}, sourcerate=q(2.5, rate_unit),
                  inputs=ns(geometry=ns(planet=ns(radius=q(setup.radius_km, u.km)))))
        src, avail = lr.LOSResult.make_source_map(self, params, normalize=normalize)
        for which, m in (('source', src), ('available', avail)):
            for key, val in vars(m).items():
                out[f'{tag}_{which}_{key}'] = np.asarray(val, dtype=np.float64)
        print(tag, 'abundance sum', float(np.asarray(src.abundance).sum()))
    np.savez_compressed(os.path.join(GOLD, 'losresult_source_map.npz'), **out)
    print('losresult_source_map.npz')


class _PercentileInterval:
    """astropy.visualization.PercentileInterval as astropy 5.3 defines it (the version the
    reference pins, poetry.lock:33-34; astropy itself is absent here): get_limits() ravels
    WHATEVER it is given, drops the non-finite values and returns
    np.percentile(values, ((100 - p) / 2, 100 - (100 - p) / 2))."""

    def __init__(self, percentile, n_samples=None):
        self.lower = (100 - percentile) * 0.5
        self.upper = 100 - self.lower

    def get_limits(self, values):
        values = np.asarray(values).ravel()
        values = values[np.isfinite(values)]
        vmin, vmax = np.percentile(values, (self.lower, self.upper))
        return vmin, vmax


class _Multiply:
    """astropy.modeling.models.Multiply: y = factor * x, one linear parameter."""

    def __init__(self, factor=1.0):
        self.factor = u.Quantity(factor, u.dimensionless_unscaled) \
            if hasattr(u, 'dimensionless_unscaled') else types.SimpleNamespace(value=factor)


class _LinearLSQFitter:
    """astropy.modeling.fitting.LinearLSQFitter on a one-parameter linear model, as astropy
    5.3 does it: the design matrix column (d model / d factor = x) and the data are BOTH
    multiplied by `weights` (a length mismatch raises, as numpy broadcasting does there),
    the columns are scaled to unit norm and np.linalg.lstsq solves the system."""

    def __call__(self, model, x, y, weights=None):
        lhs = np.asarray(x, dtype=float)[:, np.newaxis].copy()
        rhs = np.asarray(y, dtype=float).copy()
        if weights is not None:
            weights = np.asarray(weights, dtype=float)
            lhs *= weights[:, np.newaxis]
            rhs = rhs * weights
        scl = (lhs * lhs).sum(0)
        lhs /= scl
        lacoef, *_ = np.linalg.lstsq(lhs, rhs, rcond=len(x) * np.finfo(float).eps)
        factor = float((lacoef.T / scl).T[0])
        fitted = _Multiply()
        fitted.factor = types.SimpleNamespace(value=factor, __float__=lambda: factor)
        fitted.factor = _Factor(factor)
        return fitted


class _Factor(float):
    """best_fit.factor: usable as a number and through .value (LOSResult.py:293, 301)."""
    @property
    def value(self):
        return float(self)


def golden_source_rate():
    """Execute the UNMODIFIED reference LOSResult.make_mask / determine_source_rate
    (LOSResult.py:171-200, 278-308) on synthetic spectra for every masking keyword, with
    astropy's PercentileInterval / LinearLSQFitter restated above -> tests/golden/source_rate.npz
    (masks, scale factors, scaled radiances; 'raises' where the reference itself fails)."""
    from nexoclom.data_simulation import LOSResult as lr
    lr.PercentileInterval = _PercentileInterval
    lr.models = types.SimpleNamespace(Multiply=_Multiply)
    lr.fitting = types.SimpleNamespace(LinearLSQFitter=_LinearLSQFitter)
    ns = types.SimpleNamespace
    rng = np.random.default_rng(12)
    n = 240
    model = rng.random(n) * (rng.random(n) > 0.1)
    data = pd.DataFrame({'radiance': 2.5 * model + rng.normal(0, 0.05, n),
                         'sigma': 0.02 + 0.2 * rng.random(n),
                         'alttan': rng.random(n) * 2,
                         'x': rng.normal(0, 2, n), 'xbore': rng.normal(0, 1, n)})
    out = {'model': model}
    for c in data.columns:
        out['data_' + c] = data[c].values
    cases = [None, 'middle90', 'middle50; minalt0.3', 'minsnr3', 'minalt0.5; minsnr2',
             'siglimit2', 'siglimit50', 'minsnr3; siglimit3']
    out['cases'] = np.array([str(c) for c in cases])
    for ic, masking in enumerate(cases):
        for use_weight in (False, True):
            tag = f'c{ic}_w{int(use_weight)}'
            me = ns(masking=masking, radiance=pd.Series(model.copy()))
            me.make_mask = types.MethodType(lr.LOSResult.make_mask, me)
            mask0, siglimit = me.make_mask(data)
            out[tag + '_mask0'] = np.asarray(mask0, dtype=bool)
            try:
                lr.LOSResult.determine_source_rate(me, ns(data=data), use_weight=use_weight)
            except ValueError as err:            # the refit with the first mask's weights
                out[tag + '_raises'] = np.array(str(err)[:60])
                print(tag, masking, 'raises', str(err)[:60])
                continue
            out[tag + '_mask'] = np.asarray(me.mask, dtype=bool)
            out[tag + '_factor'] = float(me.sourcerate.value)
            out[tag + '_radiance'] = me.radiance.values
            print(tag, masking, float(me.sourcerate.value), int(np.sum(me.mask)))
    np.savez_compressed(os.path.join(GOLD, 'source_rate.npz'), **out)
    print('source_rate.npz')


if __name__ == '__main__':
    install()
    which = sys.argv[1:] or ['source', 'image', 'los', 'map']
    if 'map' in which:
        golden_source_map()
    if 'source' in which:
        golden_source_distribution()
    if 'image' in which:
        golden_image()
    if 'los' in which:
        golden_los()
    if 'losfit' in which:
        golden_losfit()
    if 'losmap' in which:
        golden_losresult_source_map()
    if 'rate' in which:
        golden_source_rate()
