"""LOSResultFitted: the vectorised CSR reductions against the loop restatement of the
reference (oracle/losfit.py), and the public class on the GPU."""
import numpy as np
import pytest

from nexoclom_b200.LOSResultFitted import fit_packet_weights, fitted_radiance
from oracle import losfit


@pytest.mark.parametrize('use_weight', [None, 'dist2', 'dist', 'sigma'])
def test_reweighting_matches_loop_restatement(use_weight):
    rng = np.random.default_rng(8)
    npk, n0, nspec = 4000, 5000, 60
    index0 = rng.choice(n0, npk, replace=False)
    xyz = rng.normal(size=(npk, 3)) * 3
    frac = rng.random(npk)
    gsum = rng.random(npk) * 5
    sc = rng.normal(size=(nspec, 3)) * 6
    used = [sorted(rng.choice(npk, rng.integers(0, 200), replace=False).tolist())
            for _ in range(nspec)]
    used[7] = []
    data_rad = rng.random(nspec) * 10
    model_rad = rng.random(nspec) * 10
    model_rad[3] = 0.0                                  # inf ratio -> reference keeps inf
    data_rad[3] = 0.0                                   # 0/0 -> NaN -> 0
    mask = rng.random(nspec) > 0.2
    sigma = 0.1 + rng.random(nspec)
    w_ref, frac_ref, rad_ref = losfit.fit(used, index0, xyz, frac, gsum, n0, sc, data_rad,
                                          model_rad, mask, sigma, use_weight,
                                          np.radians(1.0), 2.44e8)
    off = np.zeros(nspec + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(u) for u in used])
    rows = np.concatenate([np.asarray(u, dtype=np.int64) for u in used])
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = data_rad / model_rad
    ratio[np.isnan(ratio)] = 0
    w = fit_packet_weights(off, rows, index0, n0, sc, xyz, ratio, mask, sigma, use_weight)
    assert np.allclose(w, w_ref, rtol=1e-12, atol=0)
    new_frac = frac * w[index0]
    rad = fitted_radiance(off, rows, sc, xyz, new_frac * gsum / 1e6, np.radians(1.0), 2.44e8)
    assert np.allclose(new_frac, frac_ref, rtol=1e-12)
    assert np.allclose(rad, rad_ref, rtol=1e-12)
    assert rad[7] == 0.0


@pytest.mark.parametrize('tag, use_weight', [('none', None), ('dist', 'dist'),
                                             ('dist2', 'dist2'), ('sigma', 'sigma')])
def test_reweighting_vs_reference_golden(tag, use_weight):
    """tests/golden/losfit.npz: outputs of the UNMODIFIED reference method
    LOSResultFitted.determine_source_from_data (tools/make_golden_products.py losfit) on the
    packets, lines of sight and `used` sets of los.npz.  The oracle restatement and the
    vectorised product code both reproduce the re-weighted packets and the fitted radiance."""
    import os
    from common import GOLDEN, workload
    from nexoclom_b200.LOSResultFitted import fit_packet_weights, fitted_radiance
    from nexoclom_b200.runsetup import RunSetup
    g = np.load(os.path.join(GOLDEN, 'los.npz'))
    f = np.load(os.path.join(GOLDEN, 'losfit.npz'))
    X, los = g['X'], g['los']
    n, nlos = len(X), len(los)
    off, idx = g['d3_used_off'], g['d3_used_idx']
    used = [idx[off[i]:off[i + 1]].tolist() for i in range(nlos)]
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    gsum = np.zeros(n)
    for v, gv in setup.gtables([5891, 5897]):
        gsum += np.interp(X[:, 5] + setup.vrplanet, v, gv)
    index0 = np.arange(n)
    dphi, rp_cm = float(f['dphi']), setup.radius_km * 1e5
    w_o, frac_o, rad_o = losfit.fit(used, index0, X[:, 1:4], X[:, 7], gsum, n, los[:, :3],
                                    f['data_radiance'], f['model_radiance'], f['mask'],
                                    f['sigma'], use_weight, dphi, rp_cm)
    assert np.allclose(w_o, f[f'{tag}_frac0'], rtol=1e-12, atol=0)      # X0.frac was 1
    assert np.allclose(frac_o, f[f'{tag}_frac'], rtol=1e-12, atol=0)
    ref = f[f'{tag}_radiance']
    assert (ref > 0).sum() > 50
    assert np.allclose(rad_o, ref, rtol=1e-11, atol=0)
    assert abs(frac_o.sum() * 0 + w_o.sum() - float(f[f'{tag}_totalsource'])) < 1e-9 * n

    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.nan_to_num(f['data_radiance'] / f['model_radiance'], nan=0.0, posinf=np.inf)
    w = fit_packet_weights(off, idx, index0, n, los[:, :3], X[:, 1:4], ratio, f['mask'],
                           f['sigma'], use_weight)
    assert np.allclose(w, f[f'{tag}_frac0'], rtol=1e-12, atol=0)
    rad = fitted_radiance(off, idx, los[:, :3], X[:, 1:4], X[:, 7] * w * gsum / 1e6, dphi, rp_cm)
    assert np.allclose(rad, ref, rtol=1e-11, atol=0)


@pytest.mark.parametrize('use_weight', [False, True])
@pytest.mark.parametrize('masking', [None, 'minsnr3; minalt0.5', 'siglimit2'])
def test_determine_source_rate_is_astropy_linear_lsq(use_weight, masking):
    """reference LOSResult.py:171-200, 278-308 with astropy's LinearLSQFitter restated as what
    it does: scale both sides by the weights, np.linalg.lstsq."""
    import types
    import pandas as pd
    from nexoclom_b200.LOSResult import LOSResult
    rng = np.random.default_rng(2)
    n = 300
    model = rng.random(n) * (rng.random(n) > 0.1)
    data = pd.DataFrame({'radiance': 2.5 * model + rng.normal(0, 0.05, n),
                         'sigma': 0.02 + 0.2 * rng.random(n), 'alttan': rng.random(n) * 2})
    me = types.SimpleNamespace(masking=masking, radiance=pd.Series(model.copy()),
                               reference_exact=False)      # weights follow the clipped mask
    me.make_mask = types.MethodType(LOSResult.make_mask, me)
    LOSResult.determine_source_rate(me, types.SimpleNamespace(data=data), use_weight=use_weight)

    mask = np.ones(n, dtype=bool)
    if masking and 'minsnr' in masking:
        mask = (data.radiance / data.sigma > 3).values & (data.alttan >= 0.5).values

    def lsq(msk):
        w = 1 / data.sigma.values[msk]**2 if use_weight else np.ones(msk.sum())
        lhs = (model[msk] * w)[:, None]
        return np.linalg.lstsq(lhs, data.radiance.values[msk] * w, rcond=None)[0][0]
    f = lsq(mask)
    if masking == 'siglimit2':
        mask = mask & (np.abs((data.radiance.values - f * model) / data.sigma.values) < 2)
        assert mask.sum() < n
        f = lsq(mask)
    assert float(me.sourcerate) == pytest.approx(f, rel=1e-12)
    assert np.array_equal(np.asarray(me.mask), mask)
    assert np.allclose(me.radiance.values, model * f, rtol=1e-12)


def test_make_mask_and_source_rate_vs_reference_golden():
    """tests/golden/source_rate.npz: the UNMODIFIED reference LOSResult.make_mask /
    determine_source_rate (LOSResult.py:171-200, 278-308) for every masking keyword --
    `middleNN` over the whole data frame, the refit after `siglimit` that raises whenever a
    point was clipped -- against the product methods in their default (reference_exact) mode."""
    import os
    import types
    import pandas as pd
    from common import GOLDEN
    from nexoclom_b200.LOSResult import LOSResult
    g = np.load(os.path.join(GOLDEN, 'source_rate.npz'))
    data = pd.DataFrame({c[5:]: g[c] for c in g.files if c.startswith('data_')})
    data = data[['radiance', 'sigma', 'alttan', 'x', 'xbore']]        # the generator's column order
    nraise = 0
    for ic, masking in enumerate(g['cases']):
        masking = None if masking == 'None' else str(masking)
        for w in (0, 1):
            tag = f'c{ic}_w{w}'
            me = types.SimpleNamespace(masking=masking, radiance=pd.Series(g['model'].copy()))
            me.make_mask = types.MethodType(LOSResult.make_mask, me)
            mask0, _ = me.make_mask(data)
            assert np.array_equal(mask0, g[tag + '_mask0']), tag
            if tag + '_raises' in g.files:
                with pytest.raises(ValueError, match='could not be broadcast'):
                    LOSResult.determine_source_rate(me, types.SimpleNamespace(data=data),
                                                    use_weight=bool(w))
                nraise += 1
                continue
            LOSResult.determine_source_rate(me, types.SimpleNamespace(data=data), use_weight=bool(w))
            assert np.array_equal(np.asarray(me.mask), g[tag + '_mask']), tag
            assert float(me.sourcerate) == pytest.approx(float(g[tag + '_factor']), rel=1e-12)
            assert np.allclose(me.radiance.values, g[tag + '_radiance'], rtol=1e-12, atol=0)
    assert nraise == 4
    # the non-default variant: percentiles of the radiances only
    me = types.SimpleNamespace(masking='middle50', radiance=None, reference_exact=False)
    mask, _ = LOSResult.make_mask(me, data)
    assert 0.45 < mask.mean() < 0.55


@pytest.mark.parametrize('tag, normalize', [('norm', True), ('raw', False)])
def test_losresult_make_source_map_vs_reference_golden(monkeypatch, tag, normalize):
    """tests/golden/losresult_source_map.npz: the UNMODIFIED reference
    LOSResult.make_source_map (LOSResult.py:310-491) over two output files with different
    speed ranges.  The per-file maps (K6 in the product) come from the oracle here, so the
    host arithmetic of the port -- sums, the interp branch, observed-fraction correction,
    flux normalisation, the reference's habits included -- is pinned on the CPU."""
    import os
    import types
    from common import GOLDEN
    from nexoclom_b200 import make_source_map as msm_mod
    from nexoclom_b200.LOSResult import LOSResult
    from nexoclom_b200.units import Quantity
    from oracle import source_map as osm
    g = np.load(os.path.join(GOLDEN, 'losresult_source_map.npz'))
    keep = ['longitude', 'latitude', 'v', 'altitude', 'azimuth', 'frac']
    files = {f: {k: g[f + '_X0'][:, i] for i, k in enumerate(keep)} for f in ('f1', 'f2')}
    params = {k[6:]: (float(g[k]) if k == 'param_smear_radius' else int(g[k]))
              for k in g.files if k.startswith('param_')}
    rp = float(g['radius_km'])

    def arrays(X0, R_planet_km, grid_params, todo, device=0):       # stands in for K6
        return osm.make_source_map(X0, R_planet_km, grid_params, todo)
    monkeypatch.setattr(msm_mod, 'source_map_arrays', arrays)
    radius = Quantity(rp, 'km')
    monkeypatch.setattr(msm_mod, 'Output', types.SimpleNamespace(
        restore=lambda fname: types.SimpleNamespace(X0=files[fname], inputs=types.SimpleNamespace(
            geometry=types.SimpleNamespace(planet=types.SimpleNamespace(radius=radius))))))
    me = types.SimpleNamespace(
        modelfiles={'f1': 'm1', 'f2': 'm2'}, sourcerate=Quantity(float(g['sourcerate_1e23']), ''),
        _device=0, inputs=types.SimpleNamespace(geometry=types.SimpleNamespace(
            planet=types.SimpleNamespace(radius=Quantity(rp, 'km')))))
    src, avail = LOSResult.make_source_map(me, params, normalize=normalize)
    checked = 0
    for which, m in (('source', src), ('available', avail)):
        for key in ('abundance', 'abundance_uncor', 'longitude', 'latitude', 'speed',
                    'speed_dist', 'altitude', 'altitude_dist', 'azimuth', 'azimuth_dist',
                    'n_included', 'n_total', 'fraction_observed', 'speed_dist_map',
                    'altitude_dist_map', 'azimuth_dist_map'):
            ref = g[f'{tag}_{which}_{key}']
            got = np.asarray(getattr(m, key), dtype=float)
            assert got.shape == ref.shape, key
            assert np.array_equal(np.isnan(got), np.isnan(ref)), key
            ok = ~np.isnan(ref)
            assert np.allclose(got[ok], ref[ok], rtol=1e-10, atol=0), (which, key)
            checked += 1
    assert checked == 32 and np.nansum(g[f'{tag}_source_abundance']) > 0


def test_use_selected_vs_reference_golden():
    """`use_selected=True` of the unmodified reference method on a constant-step-like output
    (several rows per packet, its own generator seeded): same rows kept, same re-weighted
    packets, same fitted radiance."""
    import os
    import pandas as pd
    from common import GOLDEN, workload
    from nexoclom_b200.LOSResultFitted import (fit_packet_weights, fitted_radiance,
                                               restrict_csr, select_one_step)
    from nexoclom_b200.runsetup import RunSetup
    g = np.load(os.path.join(GOLDEN, 'los.npz'))
    f = np.load(os.path.join(GOLDEN, 'losfit.npz'))
    los = g['los']
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac', 'Index']
    X = pd.DataFrame(f['sel_rows'], columns=cols, index=f['sel_row_labels'])
    n0 = int(f['sel_npackets0'])
    sel = select_one_step(X, n0, np.random.default_rng(int(f['sel_seed'])))
    assert np.array_equal(np.sort(sel.index.values), f['sel_kept_labels'])
    sel = sel.sort_index()
    off, rows = restrict_csr(f['sel_used_off'], sel.index.get_indexer(f['sel_used_idx']))
    assert 0 < len(rows) < len(f['sel_used_idx'])
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    gsum = np.zeros(len(sel))
    for v, gv in setup.gtables([5891, 5897]):
        gsum += np.interp(sel.vy.values + setup.vrplanet, v, gv)
    with np.errstate(divide='ignore', invalid='ignore'):
        ratio = np.nan_to_num(f['sel_data_radiance'] / f['sel_model_radiance'], nan=0.0,
                              posinf=np.inf)
    ind0 = sel['Index'].values.astype(np.int64)
    xyz = sel[['x', 'y', 'z']].values
    w = fit_packet_weights(off, rows, ind0, n0, los[:, :3], xyz, ratio, f['sel_mask'],
                           f['sigma'], 'dist2')
    assert np.allclose(w, f['sel_frac0'], rtol=1e-12, atol=0)
    frac = sel.frac.values * w[ind0]
    assert np.allclose(frac, f['sel_frac'], rtol=1e-12, atol=0)
    assert abs(w.sum() * 4 - float(f['sel_totalsource'])) < 1e-9 * n0      # x nsteps (:186)
    rad = fitted_radiance(off, rows, los[:, :3], xyz, frac * gsum / 1e6, float(f['dphi']),
                          setup.radius_km * 1e5)
    assert (f['sel_radiance'] > 0).sum() > 30
    assert np.allclose(rad, f['sel_radiance'], rtol=1e-11, atol=0)


def test_use_selected_restatement():
    """`use_selected` keeps one step per trajectory; literal restatement of the reference's
    set / MultiIndex selection (LOSResultFitted.py:95-113) and of its `to_use` filter."""
    import pandas as pd
    from nexoclom_b200.LOSResultFitted import restrict_csr, select_one_step
    npk, nst = 40, 7
    rng0 = np.random.default_rng(3)
    X = pd.DataFrame({'Index': np.repeat(np.arange(npk), nst),
                      'time': np.tile(np.arange(nst, 0, -1) * 30.0, npk).astype(np.float32),
                      'x': rng0.normal(size=npk * nst)})
    X = X[rng0.random(len(X)) > 0.3]                       # some steps are missing
    sel = select_one_step(X, npk, np.random.default_rng(11))
    # the reference's way
    Xr = X.copy()
    Xr['ind_'] = Xr.index
    times = Xr.time.unique()
    Xr.set_index(['Index', 'time'], inplace=True)
    steps = set(zip(np.arange(npk), np.random.default_rng(11).choice(times, npk)))
    steps = steps.intersection(set(Xr.index))
    ref = Xr.loc[pd.MultiIndex.from_tuples(sorted(steps), names=['Index', 'time'])]
    assert sorted(ref.ind_.values) == sorted(sel.index.values) and 0 < len(sel) <= npk
    assert sel['Index'].is_unique

    labels = rng0.choice(X.index.values, 60)
    off = np.array([0, 10, 10, 35, 60])
    rows = sel.index.get_indexer(labels)
    off2, rows2 = restrict_csr(off, rows)
    for i in range(4):
        to_use = [x for x in labels[off[i]:off[i + 1]] if x in sel.index]
        assert list(sel.index[rows2[off2[i]:off2[i + 1]]]) == to_use


@pytest.mark.gpu
def test_losresultfitted_public_api(engine):
    """LOSResult -> LOSResultFitted through the reference-facing classes: the re-weighted
    packets reproduce the loop restatement on the K5 `used` sets, the fitted output is
    catalogued under the fitted Input, and the fit moves the model towards the data."""
    from common import workload
    from nexoclom_b200 import Output, LOSResult, LOSResultFitted
    from nexoclom_b200.runsetup import RunSetup
    from nexoclom_b200.units import Quantity
    from test_gpu_parity import _FakeSCData, _synthetic_los
    inputs = workload('Ca.isotropic.flat.input')
    inputs.delete_files()
    out = Output(inputs, 40000, seed=5)
    los = _synthetic_los(150, seed=9)
    truth = 1.0 + np.abs(np.sin(np.arange(150) * 0.3)) * 4
    sc = _FakeSCData(los, truth)
    unfit = LOSResult(sc, inputs, dphi=Quantity(2.0, 'deg'), label='unfit')
    unfit.simulate_data_from_inputs(sc)
    sc.model_result = {'unfit': unfit}
    sc.data['mask_unfit'] = unfit.mask

    fitted = LOSResultFitted(sc, 'unfit', dphi=Quantity(2.0, 'deg'), label='fitted')
    assert fitted.inputs.options.fitted and not unfit.inputs.options.fitted
    fitted.determine_source_from_data(sc, use_weight='dist2')
    assert len(fitted.outputfiles) == 1 and fitted.outputfiles[0] != out.filename
    # loop restatement on the same `used` sets
    P = Output.restore(out.filename)
    it = unfit._iterations[out.filename]
    off, idx0, labels = it.used_csr
    rows = P.X.index.get_indexer(labels)
    used = [rows[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]
    setup = RunSetup(inputs)
    gsum = np.zeros(len(P.X))
    for v, g in setup.gtables([4227]):
        gsum += np.interp(P.X.vy.values + setup.vrplanet, v, g)
    w_ref, frac_ref, rad_ref = losfit.fit(
        used, P.X['Index'].values, P.X[['x', 'y', 'z']].values, P.X.frac.values, gsum, len(P.X0),
        los[:, :3], truth, unfit.radiance.values, unfit.mask, sc.data.sigma.values, 'dist2',
        np.radians(2.0), setup.radius_km * 1e5)
    res = fitted._iterations[fitted.outputfiles[0]]
    assert np.allclose(res.weighting, w_ref, rtol=1e-10)
    Pf = Output.restore(fitted.outputfiles[0])
    assert np.allclose(Pf.X0.frac.values, (P.X0.frac.values * w_ref).astype(np.float32), rtol=1e-6)
    assert np.allclose(res.radiance.values, rad_ref, rtol=1e-10)
    # better agreement with the data than the unfitted model (both scaled by their source rate)
    m = unfit.mask & (unfit.radiance.values > 0)
    err0 = np.mean((unfit.radiance.values[m] - truth[m])**2)
    err1 = np.mean((fitted.radiance.values[m] - truth[m])**2)
    assert err1 < err0


@pytest.mark.gpu
def test_losresult_make_source_map(engine):
    """LOSResult.make_source_map (reference LOSResult.py:310-491) on top of K6: sums over the
    output files, observed-fraction correction, flux normalisation."""
    from common import workload
    from nexoclom_b200 import Output, LOSResult
    from nexoclom_b200.units import Quantity
    from test_gpu_parity import _FakeSCData, _synthetic_los
    inputs = workload('Ca.isotropic.flat.input')
    inputs.delete_files()
    Output(inputs, 20000, seed=1)
    Output(inputs, 20000, seed=2)
    los = _synthetic_los(80, seed=4)
    sc = _FakeSCData(los, np.linspace(1.0, 2.0, 80))
    res = LOSResult(sc, inputs, dphi=Quantity(2.0, 'deg'))
    res.simulate_data_from_inputs(sc)
    assert len(res.modelfiles) == 2
    grid = {'nlonbins': 36, 'nlatbins': 18, 'nvelbins': 20, 'nazbins': 8, 'naltbins': 6}
    raw_s, raw_a = res.make_source_map(grid, normalize=False)
    assert raw_a.n_total.sum() > 0 and np.array_equal(raw_a.n_total, raw_s.n_total)
    assert raw_a.abundance_uncor.shape == (36, 18)
    # 'available' weights every packet by 1: smeared abundance == number of packets in the ball
    assert np.allclose(raw_a.abundance_uncor, raw_a.n_total)
    assert np.all(raw_a.fraction_observed <= 1) and np.all(raw_a.fraction_observed >= 0)
    src, avail = res.make_source_map(grid, normalize=True)
    # normalised abundance integrates to the source rate over the sphere
    lon, lat = np.asarray(src.longitude), np.asarray(src.latitude)
    dx, dy = lon[1] - lon[0], lat[1] - lat[0]
    r_cm = 2440.53e5
    area = r_cm**2 * np.abs(dx * (np.sin(lat + dy / 2) - np.sin(lat - dy / 2)))[np.newaxis, :]
    total = np.nansum(np.asarray(src.abundance) * area)
    assert total == pytest.approx(float(res.sourcerate) * 1e23, rel=1e-9)
    inputs.delete_files()
