"""GPU parity: the CUDA path (through the C ABI) against the oracle and the
golden vectors of the reference.  Gates (BASELINE.json north_star): packet state
1e-8 relative (compared before the f32 down-cast), images / LOS radiance 1e-6
relative, pixel indices and per-LOS hit counts bit-exact."""
import os

import numpy as np
import pytest

from common import GOLDEN, workload, oracle_constants, state_parity
from nexoclom_b200._lib import ImageParams, LosParams
from nexoclom_b200.ModelImage import image_rotation
from nexoclom_b200.runsetup import RunSetup
from nexoclom_b200.units import Quantity
from oracle import tracking, initial_state, imaging

pytestmark = pytest.mark.gpu
STATE_TOL = 1e-8
IMAGE_TOL = 1e-6


def _image_params(setup, quantity, view=(0.0, np.pi / 2), dims=(800, 800), round_f32=0):
    M = image_rotation(*view)
    ip = ImageParams()
    for k in range(9):
        ip.M[k] = float(M.flat[k])
    ip.x0, ip.x1, ip.z0, ip.z1 = -4, 4, -4, 4
    ip.nx, ip.nz = dims
    rcm = setup.radius_km * 1e5
    ip.apix = (8 / dims[0] * rcm) * (8 / dims[1] * rcm)
    ip.vrplanet = setup.vrplanet
    ip.quantity = quantity
    ip.round_f32 = round_f32
    ip.skip_dead = 0
    return ip


@pytest.mark.parametrize('wl', ['Na.maxwellian.radpres.input', 'Ca.isotropic.flat.input'])
@pytest.mark.parametrize('strict', [False, True])
def test_adaptive_driver_vs_oracle(engine, wl, strict):
    setup = RunSetup(workload(wl), strict_math=strict)
    setup.upload(engine)
    n = 3000
    X0 = initial_state.draw_x0(setup, n, 21)[:, :8]
    engine.import_state(X0)
    att, acc = engine.integrate_adaptive()
    Xg = engine.export_state().T
    a_g, c_g = engine.export_stats()
    Xo, a_o, c_o = tracking.integrate_adaptive(X0, oracle_constants(setup))
    par = state_parity(Xg, Xo)
    assert par['alive_mismatch'] == 0, par
    assert max(par['pos'], par['vel'], par['frac']) < STATE_TOL, par
    assert att == int(a_o.sum()) and acc == int(c_o.sum())      # same step sequences
    assert np.array_equal(a_g, a_o) and np.array_equal(c_g, c_o)
    dead = Xo[:, 7] == 0
    assert np.all(Xg[dead, 0] == 0)                              # dead => time = 0 (Q8)


def test_config0_at_its_named_size_vs_golden(engine):
    """BASELINE configs[0] at the size it names (1e5 Ca packets): attempted / accepted step
    counts of EVERY packet, the final state of every 16th packet and the column sums of all
    final states against tests/golden/config0_1e5.npz (the oracle's port of the reference
    driver, 3 min of CPU: tools/make_golden_config0.py)."""
    g = np.load(os.path.join(GOLDEN, 'config0_1e5.npz'))
    n, seed, stride = int(g['n']), int(g['seed']), int(g['stride'])
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    setup.upload(engine)
    X0 = initial_state.draw_x0(setup, n, seed)[:, :8].astype(np.float32).astype(np.float64)
    if not np.array_equal(X0.sum(axis=0), g['x0_sums']):          # the same initial state?
        pytest.skip('initial state not bit-reproducible on this host (NumPy SIMD sin / cos)')
    engine.import_state(X0)
    att, acc = engine.integrate_adaptive()
    Xg = engine.export_state().T
    a_g, c_g = engine.export_stats()
    assert np.array_equal(a_g, g['attempted']) and np.array_equal(c_g, g['accepted'])
    assert att == int(g['attempted'].astype(np.int64).sum())
    par = state_parity(Xg[::stride], g['final_subset'])
    assert par['alive_mismatch'] == 0, par
    assert max(par['pos'], par['vel'], par['frac']) < STATE_TOL, par
    assert np.allclose(Xg.sum(axis=0), g['final_sums'], rtol=0, atol=1e-9 * g['final_abs_sums'].max())
    assert np.allclose(np.abs(Xg).sum(axis=0), g['final_abs_sums'], rtol=1e-10)


def test_adaptive_import_mode_vs_reference_golden(engine):
    """Reference-generated initial states -> final state of the reference's own
    driver (tests/golden/adaptive_driver.npz)."""
    g = np.load(os.path.join(GOLDEN, 'adaptive_driver.npz'))
    for tag, wl in (('na', 'Na.maxwellian.radpres.input'), ('ca', 'Ca.isotropic.flat.input')):
        setup = RunSetup(workload(wl))
        setup.upload(engine)
        engine.import_state(g[f'{tag}_x0'])
        engine.integrate_adaptive()
        par = state_parity(engine.export_state().T, g[f'{tag}_final'])
        assert par['alive_mismatch'] == 0, par
        assert max(par['pos'], par['vel'], par['frac']) < STATE_TOL, par
        step = engine.export_step()
        assert np.max(np.abs(step - g[f'{tag}_step']) / g[f'{tag}_step']) < STATE_TOL


def test_energy_conservation_gravity_only(engine):
    """The reference's only hot-path test (test_gravity.py:46-55): specific
    energy v^2/2 + GM/r is constant along each trajectory -- at full chunk size."""
    inputs = workload('Gravity.input')
    inputs.options.step_size = 0.
    inputs.options.resolution = 1e-4
    inputs.options.lifetime = Quantity(1e30, 's')        # no loss: isolate the orbit
    setup = RunSetup(inputs)
    setup.upload(engine)
    n = 1_000_000
    engine.init_state(setup.source_params(engine), 3, 0, n)
    x0 = engine.export_x0()
    engine.integrate_adaptive()
    x = engine.export_state()
    alive = x[7] > 0
    assert alive.sum() > 1000

    def energy(a):
        return 0.5 * (a[4]**2 + a[5]**2 + a[6]**2) + setup.GM / np.sqrt(a[1]**2 + a[2]**2 + a[3]**2)
    e0, e1 = energy(x0[:8])[alive], energy(x)[alive]
    assert np.max(np.abs(e1 - e0) / np.abs(e0)) < 1e-6
    assert np.all(x[0][alive] <= 1e-4)                   # everybody reached the image time


@pytest.mark.parametrize('wl', ['Na.maxwellian.radpres.input', 'Ca.isotropic.flat.input',
                                'Na.bounce.stick05.input'])
def test_init_state_vs_oracle(engine, wl):
    setup = RunSetup(workload(wl))
    setup.upload(engine)
    n = 20000
    engine.init_state(setup.source_params(engine), 42, 1000, n)
    got = engine.export_x0().T
    ref = initial_state.draw_x0(setup, n, 42, first_id=1000)
    assert np.max(np.abs(got - ref)) < 1e-12
    assert np.array_equal(engine.export_state().T, got[:, :8])
    assert np.all(engine.export_step() == 1000.)


def test_init_state_sharding_invariance(engine):
    """Global-id Philox counters: two shards == one run (GPU-count invariance)."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    sp = setup.source_params(engine)
    engine.init_state(sp, 9, 0, 4096)
    whole = engine.export_x0()
    engine.init_state(sp, 9, 0, 2048)
    a = engine.export_x0()
    engine.init_state(sp, 9, 2048, 2048)
    b = engine.export_x0()
    assert np.array_equal(whole, np.concatenate([a, b], axis=1))


@pytest.mark.parametrize('quantity', [0, 1])
@pytest.mark.parametrize('view', [(0.0, np.pi / 2), (0.7, 0.3)])
def test_image_vs_oracle(engine, quantity, view):
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    gt = setup.gtables([5891, 5897])
    engine.upload_gtables(gt)
    rng = np.random.default_rng(5)
    n = 300001
    X = np.zeros((n, 8))
    X[:, 1:4] = rng.normal(size=(n, 3)) * 1.7
    X[:, 5] = rng.normal(size=n) * 2 / setup.radius_km
    X[:, 7] = rng.random(n)
    X[:64, 1] = 4.0
    X[64:128, 3] = -4.0
    X = X.astype(np.float32).astype(np.float64)        # what Output.restore delivers (Q14)
    engine.import_state(X)
    ip = _image_params(setup, quantity, view, dims=(800, 640))
    img, cnt = engine.image_accumulate(ip)
    oi, oc, _, _ = imaging.create_image(X[:, 1], X[:, 2], X[:, 3], X[:, 5], X[:, 7],
                                        vrplanet=setup.vrplanet, M=imaging.image_rotation(*view),
                                        dims=[800, 640], xrange=(-4, 4), zrange=(-4, 4),
                                        apix=ip.apix, quantity='radiance' if quantity else 'column',
                                        gtables=gt)
    assert np.array_equal(cnt, oc.astype(np.int64))     # bit-exact pixel indexing
    assert cnt.sum() > 0.9 * n * 0.9
    nz = oi > 0
    assert np.max(np.abs(img[nz] - oi[nz]) / oi[nz]) < IMAGE_TOL
    assert np.all(img[~nz] == 0)


@pytest.mark.parametrize('strict', [False, True])
@pytest.mark.parametrize('tag, wl', [('tdep', 'Na.bounce.input'),
                                     ('c05', 'Na.bounce.stick05.input'),
                                     ('grav', 'Gravity.input')])
def test_constant_driver_vs_oracle(engine, tag, wl, strict):
    inputs = workload(wl)
    inputs.options.endtime = Quantity(1500., 's')
    setup = RunSetup(inputs, strict_math=strict)
    setup.upload(engine)
    n, seed, first = 2000, 99, 7
    X0 = initial_state.draw_x0(setup, n, 9)[:, :8]
    ref, nsteps, natt = tracking.integrate_constant(
        X0, oracle_constants(setup), uniforms=initial_state.bounce_uniforms(seed, first))
    engine.import_state(X0)
    traj, ns, steps = engine.integrate_constant(seed=seed, first_id=first, trajectory=True)
    assert ns == nsteps and traj.shape == ref.shape
    assert np.array_equal(traj[:, 7, :] > 0, ref[:, 7, :] > 0)
    scale = np.maximum(np.abs(ref), 1e-3)
    assert np.max(np.abs(traj - ref) / scale) < STATE_TOL
    assert steps == int(natt.sum())


def test_constant_driver_vs_reference_golden(engine):
    """Gravity-only constant-step run of the reference driver itself."""
    g = np.load(os.path.join(GOLDEN, 'constant_driver.npz'))
    setup = RunSetup(workload('Gravity.input'))
    setup.upload(engine)
    engine.import_state(g['grav_x0'])
    traj, ns, _ = engine.integrate_constant(trajectory=True)
    ref = g['grav_traj']
    assert traj.shape == ref.shape
    assert np.array_equal(traj[:, 7, :] > 0, ref[:, 7, :] > 0)
    assert np.max(np.abs(traj - ref) / np.maximum(np.abs(ref), 1e-3)) < STATE_TOL


def test_fused_image_equals_separate(engine):
    """K3 with fused per-step accumulation == K4 over the dense trajectory."""
    import torch
    inputs = workload('Na.bounce.stick05.input')
    inputs.options.endtime = Quantity(600., 's')
    setup = RunSetup(inputs)
    setup.upload(engine)
    gt = setup.gtables([5891, 5897])
    engine.upload_gtables(gt)
    n = 5000
    X0 = initial_state.draw_x0(setup, n, 3)[:, :8]
    ip = _image_params(setup, 1, dims=(200, 200))
    ip.skip_dead = 1
    img = torch.zeros((200, 200), dtype=torch.float64, device='cuda')
    cnt = torch.zeros((200, 200), dtype=torch.int64, device='cuda')
    engine.import_state(X0)
    traj, ns, _ = engine.integrate_constant(seed=1, image_params=ip, image_dev=img.data_ptr(),
                                            counts_dev=cnt.data_ptr(), trajectory=True)
    torch.cuda.synchronize()
    rows = traj.transpose(0, 2, 1).reshape(-1, 8)
    rows = rows[rows[:, 7] > 0]
    oi, oc, _, _ = imaging.create_image(rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 5], rows[:, 7],
                                        vrplanet=setup.vrplanet, M=imaging.image_rotation(0, np.pi / 2),
                                        dims=[200, 200], xrange=(-4, 4), zrange=(-4, 4),
                                        apix=ip.apix, quantity='radiance', gtables=gt)
    assert np.array_equal(cnt.cpu().numpy(), oc.astype(np.int64))
    nz = oi > 0
    assert np.max(np.abs(img.cpu().numpy()[nz] - oi[nz]) / oi[nz]) < IMAGE_TOL


def test_constant_step_kernel_repeatable(engine):
    """K3's slot pool, bounce batching and fused image are schedule-dependent machinery (shared-
    memory slots, warp ballots, global atomics): six back-to-back runs of the same 60 000 packets
    (configs[2] physics, device-drawn) must give bit-identical final states, step totals and
    per-pixel packet counts; the float image may differ in the order of its atomic additions."""
    setup = RunSetup(workload('Na.bounce.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    sp = setup.source_params(engine)
    ip = _image_params(setup, 1, dims=(200, 200))
    ip.skip_dead, ip.round_f32 = 1, 1
    n = 60_000
    ref = None
    for rep in range(6):
        engine.init_state(sp, 7, 0, n)
        engine.image_begin(200, 200)
        a, b = engine.image_device_ptrs()
        _, _, steps = engine.integrate_constant(seed=11, image_params=ip, image_dev=a,
                                                counts_dev=b, n=n)
        x = engine.export_state()
        img, cnt = engine.image_fetch(200, 200)
        if ref is None:
            ref = (x, steps, cnt, img)
            assert steps > 100 * n and cnt.sum() > 1e6
        assert steps == ref[1] and np.array_equal(x, ref[0]) and np.array_equal(cnt, ref[2])
        nz = ref[3] > 0
        assert np.array_equal(nz, img > 0)
        assert np.max(np.abs(img[nz] - ref[3][nz]) / ref[3][nz]) < 1e-12


def _synthetic_los(nlos, seed=1):
    """MESSENGER-UVVS-like sweep: spacecraft on a polar ellipse 1.1-6 R_p,
    boresights sweeping limb tangent altitudes 0-3 R_p."""
    rng = np.random.default_rng(seed)
    th = rng.random(nlos) * 2 * np.pi
    r = 1.1 + 4.9 * rng.random(nlos)
    x_sc = np.stack([0.3 * r * np.cos(th), r * np.cos(th) * 0.2 - 0.5, r * np.sin(th)], axis=1)
    x_sc *= (np.maximum(np.linalg.norm(x_sc, axis=1), 1.1) / np.linalg.norm(x_sc, axis=1))[:, None]
    target = rng.normal(size=(nlos, 3))
    target *= ((1 + 3 * rng.random(nlos)) / np.linalg.norm(target, axis=1))[:, None]
    bore = target - x_sc
    bore /= np.linalg.norm(bore, axis=1)[:, None]
    return np.concatenate([x_sc, bore], axis=1)


@pytest.mark.parametrize('los_mode', [1, 2])
def test_los_vs_oracle(engine, los_mode):
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    gt = setup.gtables([5891, 5897])
    engine.upload_gtables(gt)
    rng = np.random.default_rng(11)
    n = 200000
    X = np.zeros((n, 8))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    X[:, 1:4] = d * (1 + 9 * rng.random(n)**2)[:, None]
    X[:, 5] = rng.normal(size=n) * 2 / setup.radius_km
    X[:, 7] = rng.random(n)
    X = X.astype(np.float32).astype(np.float64)
    los = _synthetic_los(300)
    dphi = np.radians(1.0)
    used_o = []
    rad_o, np_o, inc_o, dist = imaging.los_iteration(
        X[:, 1], X[:, 2], X[:, 3], X[:, 5], X[:, 7], los, vrplanet=setup.vrplanet, dphi=dphi,
        outeredge=25., rp_cm=setup.radius_km * 1e5, gtables=gt, used=used_o)
    engine.import_state(X)
    lp = LosParams()
    lp.dphi, lp.outeredge, lp.vrplanet, lp.rp_cm = dphi, 25., setup.vrplanet, setup.radius_km * 1e5
    lp.quantity = 1
    engine.set_option('los_mode', los_mode)        # 1 brute force, 2 cell-grid culling
    try:
        rad_g, np_g, inc_g = engine.los_accumulate(los.T.copy(), dist, lp)
    finally:
        engine.set_option('los_mode', 0)
    assert np_o.sum() > 1000
    assert np.array_equal(np_g, np_o)                   # bit-exact hit counts
    assert np.array_equal(inc_g, inc_o)
    nz = rad_o > 0
    assert np.max(np.abs(rad_g[nz] - rad_o[nz]) / rad_o[nz]) < IMAGE_TOL
    assert np.all(rad_g[~nz] == 0)
    if los_mode == 2:
        # `used` packet sets as CSR
        off, idx = engine.los_used(los.T.copy(), dist, lp)
        assert off[-1] == sum(len(u) for u in used_o) > 0
        for i, u in enumerate(used_o):
            assert set(int(k) for k in idx[off[i]:off[i + 1]]) == u


@pytest.mark.parametrize('tag', ['col_pole', 'rad_pole', 'rad_side'])
def test_image_vs_reference_golden(engine, tag):
    """K4 through the C ABI against images made by the reference's own
    ModelImage.create_image() (tests/golden/image.npz, tools/make_golden_products.py)."""
    g = np.load(os.path.join(GOLDEN, 'image.npz'))
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    engine.import_state(g['X'])
    view, dims = tuple(g[f'{tag}_view']), tuple(int(d) for d in g[f'{tag}_dims'])
    ip = _image_params(setup, 0 if tag.startswith('col') else 1, view, dims=dims)
    ip.apix = float(g[f'{tag}_apix'])
    img, cnt = engine.image_accumulate(ip)
    assert np.array_equal(cnt, g[f'{tag}_packim'].astype(np.int64))    # bit-exact pixel indexing
    ref = g[f'{tag}_image']
    nz = ref > 0
    assert nz.sum() > 1000
    assert np.max(np.abs(img[nz] - ref[nz]) / ref[nz]) < IMAGE_TOL
    assert np.all(img[~nz] == 0)


def test_los_batches_when_the_pair_buffer_fills(engine):
    """A pair buffer far too small for the sweep: the lines of sight go through in batches (an
    overflowing batch is repeated with half as many lines) -- radiance, hit counts, `included`
    and `used` sets as with one batch."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    n = 300_000
    engine.init_state(setup.source_params(engine), 3, 0, n)
    from nexoclom_b200.LOSResult import dist_from_planet_cut
    import pandas as pd
    los6 = _synthetic_los(3000)
    dist = np.asarray(dist_from_planet_cut(pd.DataFrame(
        los6, columns=['x', 'y', 'z', 'xbore', 'ybore', 'zbore'])), dtype=np.float64)
    los = los6.T.copy()
    lp = LosParams()
    lp.dphi, lp.outeredge = np.radians(3.0), 25.
    lp.vrplanet, lp.rp_cm, lp.quantity = setup.vrplanet, setup.radius_km * 1e5, 1
    engine.set_option('los_mode', 2)
    try:
        rad0, npk0, inc0, cnt0 = engine.los_accumulate(los, dist, lp, count_used=True)
        off0, idx0 = engine.los_used_fill(los, dist, lp, cnt0)
        assert npk0.sum() > 5 * (n + 1024)              # several batches with the small buffer
        engine.set_option('los_pair_cap', n + 1024)
        rad1, npk1, inc1, cnt1 = engine.los_accumulate(los, dist, lp, count_used=True)
        off1, idx1 = engine.los_used_fill(los, dist, lp, cnt1)
        off2, idx2 = engine.los_used(los, dist, lp)
    finally:
        engine.set_option('los_pair_cap', 0)
        engine.set_option('los_mode', 0)
    assert np.array_equal(npk1, npk0) and np.array_equal(inc1, inc0) and np.array_equal(cnt1, cnt0)
    assert np.allclose(rad1, rad0, rtol=1e-12, atol=0)
    assert np.array_equal(off1, off0) and np.array_equal(off2, off0)
    for i in range(0, len(off0) - 1, 37):
        ref = np.sort(idx0[off0[i]:off0[i + 1]])
        assert np.array_equal(np.sort(idx1[off1[i]:off1[i + 1]]), ref)
        assert np.array_equal(np.sort(idx2[off2[i]:off2[i + 1]]), ref)


def test_los_counted_pass_and_pair_reuse(engine):
    """nx_los_accumulate_counted + nx_los_used_fill == nx_los_accumulate + the two-pass
    nx_los_used, on the reference's own golden lines of sight: once re-resolving the candidate
    pairs the counted pass left on the device, once after a call in between dropped them (the
    fill then searches again), once with a pair buffer too small for one batch."""
    from nexoclom_b200.LOSResult import dist_from_planet_cut
    import pandas as pd
    g = np.load(os.path.join(GOLDEN, 'los.npz'))
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    engine.import_state(g['X'])
    los = g['los'].T.copy()
    data = pd.DataFrame(g['los'], columns=['x', 'y', 'z', 'xbore', 'ybore', 'zbore'])
    dist = np.asarray(dist_from_planet_cut(data), dtype=np.float64)
    lp = LosParams()
    lp.dphi, lp.outeredge = float(g['d3_dphi']), float(g['outeredge'])
    lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
    lp.quantity = 1
    rad0, npk0, inc0 = engine.los_accumulate(los, dist, lp)
    off0, idx0 = engine.los_used(los, dist, lp)
    assert off0[-1] > 1000

    def same_sets(off, idx):
        assert np.array_equal(off, off0)
        for i in range(len(off) - 1):
            assert np.array_equal(np.sort(idx[off[i]:off[i + 1]]), np.sort(idx0[off0[i]:off0[i + 1]]))

    for between in (False, True):
        rad, npk, inc, cnt = engine.los_accumulate(los, dist, lp, count_used=True)
        assert np.array_equal(npk, npk0) and np.array_equal(inc, inc0)
        assert np.array_equal(cnt, np.diff(off0))
        assert np.allclose(rad, rad0, rtol=1e-12, atol=0)
        if between:
            engine.set_option('los_mode', 0)           # any state-changing call drops the pairs
        same_sets(*engine.los_used_fill(los, dist, lp, cnt))


@pytest.mark.parametrize('los_mode', [1, 2])
@pytest.mark.parametrize('tag', ['d1', 'd3'])
def test_los_vs_reference_golden(engine, tag, los_mode):
    """K5 through the C ABI against the reference's own compute_iteration()
    (tests/golden/los.npz): radiance, hit counts, included mask, used sets."""
    from nexoclom_b200.LOSResult import dist_from_planet_cut
    import pandas as pd
    g = np.load(os.path.join(GOLDEN, 'los.npz'))
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    engine.import_state(g['X'])
    los = g['los']
    data = pd.DataFrame(los, columns=['x', 'y', 'z', 'xbore', 'ybore', 'zbore'])
    dist = np.asarray(dist_from_planet_cut(data), dtype=np.float64)
    lp = LosParams()
    lp.dphi, lp.outeredge = float(g[f'{tag}_dphi']), float(g['outeredge'])
    lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
    lp.quantity = 1
    engine.set_option('los_mode', los_mode)
    try:
        rad, npk, inc = engine.los_accumulate(los.T.copy(), dist, lp)
        off, idx = engine.los_used(los.T.copy(), dist, lp)
    finally:
        engine.set_option('los_mode', 0)
    assert np.array_equal(npk, g[f'{tag}_npackets'])                  # bit-exact hit counts
    assert np.array_equal(inc, g[f'{tag}_included'])
    ref = g[f'{tag}_radiance']
    nz = ref > 0
    assert np.max(np.abs(rad[nz] - ref[nz]) / ref[nz]) < IMAGE_TOL
    assert np.all(rad[~nz] == 0)
    roff, ridx = g[f'{tag}_used_off'], g[f'{tag}_used_idx']
    assert np.array_equal(off, roff)
    for i in range(len(los)):
        assert np.array_equal(np.sort(idx[off[i]:off[i + 1]]), ridx[roff[i]:roff[i + 1]])


@pytest.mark.parametrize('todo', ['source', 'available'])
def test_source_map_vs_reference_golden(engine, todo):
    """K6 through the C ABI against the reference's own make_source_map()
    (tests/golden/source_map.npz): ball membership counts bit-exact, histograms 1e-6."""
    from nexoclom_b200.make_source_map import source_map_arrays
    from test_oracle_products_golden import check_source_map, source_map_inputs
    g = np.load(os.path.join(GOLDEN, 'source_map.npz'))
    X0, rkm, params = source_map_inputs(g)
    res = source_map_arrays(X0, rkm, params, todo)
    check_source_map(res, g, todo, tol=IMAGE_TOL)


def test_source_map_large_vs_oracle(engine):
    """Default 180 x 90 grid, 2e5 packets over the whole sphere (longitude wrap, poles)."""
    from nexoclom_b200.make_source_map import source_map_arrays
    from oracle import source_map
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    X0a = initial_state.draw_x0(setup, 200_000, 23)
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac', 'v', 'longitude', 'latitude',
            'local_time', 'altitude', 'azimuth']
    X0 = {c: X0a[:, k].astype(np.float32).astype(np.float64) for k, c in enumerate(cols)}
    rng = np.random.default_rng(1)
    X0['frac'] = rng.random(len(X0['v'])) * (rng.random(len(X0['v'])) > 0.5)
    params = {'smear_radius': np.radians(6.0), 'nlonbins': 90, 'nlatbins': 45}
    got = source_map_arrays(X0, setup.radius_km, params, 'source')
    ref = source_map.make_source_map(X0, setup.radius_km, params, 'source')
    assert np.array_equal(got['n_total'], ref['n_total'].astype(np.int64))
    assert np.array_equal(got['n_included'], ref['n_included'].astype(np.int64))
    for k in ('abundance_hist', 'abundance', 'speed_dist', 'altitude_dist', 'azimuth_dist',
              'speed_map', 'altitude_map', 'azimuth_map'):
        assert np.max(np.abs(got[k] - ref[k])) <= IMAGE_TOL * max(np.max(np.abs(ref[k])), 1.0), k


def test_maxwellian_speeds_histogram(engine, tmp_path):
    """The reference's own sampler test (tests/unit_tests/math/test_randomdeviates.py:34-43),
    on K1: 1e7 Maxwellian speeds (Na, 1500 K) histogrammed against MaxwellianDist, both
    normalised to mean 1; mean and std of the squared deviation below 1e-3."""
    from nexoclom_b200 import Input
    from nexoclom_b200.surfaceinteraction import MaxwellianDist, thermal_speed_kms
    src = open(os.path.join(os.path.dirname(__file__), '..', 'nexoclom_b200', 'workloads',
                            'Na.maxwellian.radpres.input')).read()
    f = tmp_path / 'maxw1500.input'
    f.write_text(src.replace('SpeedDist.temperature = 1200.', 'SpeedDist.temperature = 1500.'))
    setup = RunSetup(Input(str(f)))
    setup.upload(engine)
    n = 10_000_000
    engine.init_state(setup.source_params(engine), 3, 0, n)
    v = engine.export_x0()[8] * setup.radius_km                     # km/s
    vth = thermal_speed_kms(1500., 'Na')
    edges = np.linspace(0.1, 5 * vth, 201)                          # the sampler's own range
    hist, _ = np.histogram(v, bins=edges)
    assert hist.sum() == n
    x = edges[:-1] + (edges[1] - edges[0]) / 2
    h = hist / hist.mean()
    f1 = MaxwellianDist(x, 1500., 'Na')
    f1 = f1 / f1.mean()
    d2 = (h - f1)**2
    assert d2.mean() < 1e-3 and d2.std() < 1e-3, (d2.mean(), d2.std())


@pytest.mark.parametrize('smear_deg, nlon, nlat', [(100.0, 12, 6), (0.5, 24, 12), (30.0, 7, 5)])
def test_source_map_edge_cases(engine, smear_deg, nlon, nlat):
    """Smear radius larger than a hemisphere (every point reaches every packet), smaller than
    a grid cell, odd grids; packets at the poles and on the longitude seam; empty input."""
    from nexoclom_b200.make_source_map import source_map_arrays
    from oracle import source_map
    rng = np.random.default_rng(2)
    n = 5000
    X0 = {'longitude': rng.random(n) * 2 * np.pi, 'latitude': np.arcsin(rng.random(n) * 2 - 1),
          'v': rng.random(n) * 1e-3 + 1e-4, 'altitude': rng.random(n) * np.pi / 2,
          'azimuth': rng.random(n) * 2 * np.pi, 'frac': rng.random(n) * (rng.random(n) > 0.3)}
    X0['latitude'][:4] = [np.pi / 2, -np.pi / 2, 0.0, 1e-9]
    X0['longitude'][:4] = [0.0, np.pi, 0.0, 2 * np.pi - 1e-12]
    params = {'smear_radius': np.radians(smear_deg), 'nlonbins': nlon, 'nlatbins': nlat,
              'nvelbins': 10, 'nazbins': 8, 'naltbins': 5}
    got = source_map_arrays(X0, 2440.53, params, 'available')
    ref = source_map.make_source_map(X0, 2440.53, params, 'available')
    assert np.array_equal(got['n_total'], ref['n_total'].astype(np.int64))
    assert np.array_equal(got['n_included'], ref['n_included'].astype(np.int64))
    for k in ('abundance_hist', 'abundance', 'speed_dist', 'altitude_dist', 'azimuth_dist',
              'speed_map', 'altitude_map', 'azimuth_map'):
        assert np.max(np.abs(got[k] - ref[k])) <= IMAGE_TOL * max(np.max(np.abs(ref[k])), 1.0), k
    if smear_deg >= 100:
        assert np.all(got['n_total'][:, nlat // 2] > 0.5 * n)


def test_surface_map_source_on_gpu(engine, tmp_path):
    """`SpatialDist.type = surface map` (source_distribution.py:63-95) with a map pickle in the
    layout the reference ships (longitude / latitude edges, abundance one shorter per axis):
    K1 follows the oracle draw for draw."""
    import pickle
    from nexoclom_b200 import Input
    lon = np.linspace(0, 2 * np.pi, 73)
    lat = np.linspace(-np.pi / 2, np.pi / 2, 37)
    lc, bc = 0.5 * (lon[1:] + lon[:-1]), 0.5 * (lat[1:] + lat[:-1])
    ab = 1.0 + 40 * np.exp(-((lc[:, None] - 1.0)**2 + (bc[None, :] - 0.3)**2) / 0.2)
    mapfile = tmp_path / 'map.pkl'
    with open(mapfile, 'wb') as f:
        pickle.dump({'longitude': lon, 'latitude': lat, 'abundance': ab,
                     'coordinate_system': 'solar-fixed'}, f)
    src = open(os.path.join(os.path.dirname(__file__), '..', 'nexoclom_b200', 'workloads',
                            'Ca.isotropic.flat.input')).read()
    inp = tmp_path / 'map.input'
    inp.write_text(src.replace('SpatialDist.type = uniform',
                               f'SpatialDist.type = surface map\nSpatialDist.mapfile = {mapfile}'))
    setup = RunSetup(Input(str(inp)))
    setup.upload(engine)
    n = 50000
    engine.init_state(setup.source_params(engine), 8, 0, n)
    got = engine.export_x0().T
    ref = initial_state.draw_x0(setup, n, 8)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-12
    near = (np.abs(ref[:, 9] - 1.0) < 0.5) & (np.abs(ref[:, 10] - 0.3) < 0.5)
    assert near.mean() > 0.25                      # the spot of the map attracts the packets


def test_model_image_of_large_constant_step_run_is_regenerated(engine):
    """Public API, BASELINE configs[2] pattern: an Output whose dense trajectory tensor was
    not kept gives the same ModelImage as one that kept every row."""
    from nexoclom_b200 import Output, ModelImage
    params = {'quantity': 'radiance', 'dims': '200,200'}
    inputs = workload('Na.bounce.input')
    inputs.delete_files()
    Output(inputs, 3000, seed=21, keep_trajectory=True)
    dense = ModelImage(inputs, params)
    inputs.delete_files()
    out = Output(inputs, 3000, seed=21, keep_trajectory=False)
    assert not out.trajectory_kept and out._table is None      # recipe only: nothing resident
    fused = ModelImage(inputs, params)
    assert out._X is None                                      # ... and no rows were rebuilt
    inputs.delete_files()
    assert np.array_equal(fused.packet_image, dense.packet_image) and dense.packet_image.sum() > 1e5
    nz = dense.image > 0
    assert np.max(np.abs(fused.image[nz] - dense.image[nz]) / dense.image[nz]) < IMAGE_TOL
    assert np.all(fused.image[~nz] == 0)
    assert fused.atoms_per_packet == pytest.approx(dense.atoms_per_packet, rel=1e-12)


def test_public_api_end_to_end(engine):
    """Input -> Output (device-drawn packets) -> ModelImage through the
    reference-facing classes; the image equals the oracle's create_image on the
    Output's own (f32-saved) packets."""
    from nexoclom_b200 import Output, ModelImage
    inputs = workload('Ca.isotropic.flat.input')
    inputs.delete_files()
    out = Output(inputs, 20000, seed=5)
    assert out.X0.shape == (20000, 14) and str(out.X0['x'].dtype) == 'float32'
    assert (out.X.frac > 0).all() and out.totalsource == 20000.
    ids, files, npack, tot = inputs.search()
    assert npack == 20000 and files == [out.filename]
    im = ModelImage(inputs, {'quantity': 'radiance', 'dims': '300,300'})
    restored = Output.restore(out.filename)
    P = restored.X
    setup = RunSetup(inputs)
    gt = setup.gtables([4227])
    oi, oc, _, _ = imaging.create_image(
        P.x.values, P.y.values, P.z.values, P.vy.values, P.frac.values, vrplanet=setup.vrplanet,
        M=imaging.image_rotation(0, np.pi / 2), dims=[300, 300], xrange=(-4, 4), zrange=(-4, 4),
        apix=float(im.Apix), quantity='radiance', gtables=gt)
    assert np.array_equal(im.packet_image, oc)
    oi *= im.atoms_per_packet
    nz = oi > 0
    assert nz.sum() > 100
    assert np.max(np.abs(im.image[nz] - oi[nz]) / oi[nz]) < IMAGE_TOL
    assert im.atoms_per_packet == 1e23 / (20000 / 10800.)


def test_packet_order_does_not_change_results(engine):
    """Longest-first scheduling only permutes the work queue."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    X0 = initial_state.draw_x0(setup, 20000, 4)[:, :8]
    out = []
    for order in (0, 1, 2, 3):
        engine.set_option('order_packets', order)
        engine.import_state(X0)
        att, acc = engine.integrate_adaptive()
        out.append((engine.export_state(), att, acc))
    engine.set_option('order_packets', 3)
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1:] == out[0][1:]
    with pytest.raises(Exception):
        engine.set_option('no_such_option', 1)


def test_adaptive_rejects_bounce_inputs(engine):
    """Q6: the adaptive driver only supports stickcoef == 1 (reference Output.py:309-315)."""
    inputs = workload('Na.bounce.stick05.input')
    inputs.options.step_size = 0.
    inputs.options.resolution = 1e-4
    setup = RunSetup(inputs)
    setup.upload(engine)
    engine.import_state(initial_state.draw_x0(setup, 256, 1)[:, :8])
    with pytest.raises(Exception, match='Not set up'):
        engine.integrate_adaptive()


@pytest.mark.parametrize('n', [0, 1, 31, 33, 4095, 4097])
def test_ragged_sizes(engine, n):
    """Empty and ragged packet counts through K2 / K4 / K5 (warp- and
    batch-boundary cases of the feeder, odd counts of the vector loads)."""
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    setup.upload(engine)
    gt = setup.gtables([4227])
    engine.upload_gtables(gt)
    X0 = initial_state.draw_x0(setup, max(n, 1), 8)[:n, :8]
    engine.import_state(X0 if n else np.zeros((0, 8)))
    att, acc = engine.integrate_adaptive()
    Xg = engine.export_state().T
    if n == 0:
        assert att == 0 and Xg.shape == (0, 8)
    else:
        Xo, a_o, _ = tracking.integrate_adaptive(X0, oracle_constants(setup))
        par = state_parity(Xg, Xo)
        assert par['alive_mismatch'] == 0 and att == int(a_o.sum())
        if par['n_both']:
            assert max(par['pos'], par['vel'], par['frac']) < STATE_TOL
    ip = _image_params(setup, 1, dims=(64, 64))
    img, cnt = engine.image_accumulate(ip)
    if n:
        oi, oc, _, _ = imaging.create_image(Xg[:, 1], Xg[:, 2], Xg[:, 3], Xg[:, 5], Xg[:, 7],
                                            vrplanet=setup.vrplanet, M=imaging.image_rotation(0, np.pi / 2),
                                            dims=[64, 64], xrange=(-4, 4), zrange=(-4, 4), apix=ip.apix,
                                            quantity='radiance', gtables=gt)
        assert np.array_equal(cnt, oc.astype(np.int64))
        assert np.allclose(img, oi, rtol=IMAGE_TOL, atol=0)
    else:
        assert cnt.sum() == 0 and img.sum() == 0
    los = _synthetic_los(3)
    lp = LosParams()
    lp.dphi, lp.outeredge, lp.vrplanet, lp.rp_cm = np.radians(2.0), 15., setup.vrplanet, setup.radius_km * 1e5
    lp.quantity = 1
    if n:
        ro, no_, io, dist = imaging.los_iteration(Xg[:, 1], Xg[:, 2], Xg[:, 3], Xg[:, 5], Xg[:, 7], los,
                                                  vrplanet=setup.vrplanet, dphi=lp.dphi, outeredge=15.,
                                                  rp_cm=lp.rp_cm, gtables=gt)
    else:
        dist = np.full(3, 1e30)
    for mode in (1, 2):
        engine.set_option('los_mode', mode)
        try:
            rad, npk, inc = engine.los_accumulate(los.T.copy(), dist, lp)
        finally:
            engine.set_option('los_mode', 0)
        assert rad.shape == (3,) and npk.shape == (3,) and inc.shape == (n,)
        if n:
            assert np.array_equal(npk, no_) and np.array_equal(inc, io)
            assert np.allclose(rad, ro, rtol=IMAGE_TOL, atol=0)
        else:
            assert npk.sum() == 0 and rad.sum() == 0


class _FakeSCData:
    """Duck type of MESSENGERuvvs' data object as LOSResult uses it (reference
    LOSResult.py:82-94, 211; compute_iteration.py:104-109)."""

    def __init__(self, los, radiance):
        import pandas as pd
        self.data = pd.DataFrame(los, columns=['x', 'y', 'z', 'xbore', 'ybore', 'zbore'])
        self.data['radiance'] = radiance
        self.data['sigma'] = 0.1 * radiance + 1e-3
        self.data['alttan'] = 1.0
        self.species = 'Ca'
        self.query = 'synthetic'
        self.subslong = self.data.x * 0
        self.frame = None

    def set_frame(self, frame):
        self.frame = frame

    def __len__(self):
        return len(self.data)


def test_losresult_public_api(engine):
    from nexoclom_b200 import Output, LOSResult
    inputs = workload('Ca.isotropic.flat.input')
    inputs.delete_files()
    out = Output(inputs, 30000, seed=11)
    los = _synthetic_los(200, seed=3)
    truth = np.linspace(1.0, 3.0, 200)
    sc = _FakeSCData(los, truth)
    res = LOSResult(sc, inputs, dphi=Quantity(2.0, 'deg'))
    assert sc.frame == 'Model'
    res.simulate_data_from_inputs(sc)
    P = Output.restore(out.filename).X
    setup = RunSetup(inputs)
    gt = setup.gtables([4227])
    rad_o, np_o, inc_o, _ = imaging.los_iteration(
        P.x.values, P.y.values, P.z.values, P.vy.values, P.frac.values, los,
        vrplanet=setup.vrplanet, dphi=np.radians(2.0), outeredge=15.,
        rp_cm=setup.radius_km * 1e5, gtables=gt)
    assert np.array_equal(res.npackets_los.values, np_o) and np_o.sum() > 100
    kR = rad_o * res.atoms_per_packet / 1e3
    m = kR
    factor = np.sum(m * truth) / np.sum(m * m)          # determine_source_rate, unweighted
    nz = kR > 0
    assert np.max(np.abs(res.radiance.values[nz] - (kR * factor)[nz]) / (kR * factor)[nz]) < IMAGE_TOL
    assert float(res.sourcerate) == pytest.approx(factor, rel=1e-9)
    it = res._iterations[out.filename]
    used, used0 = it.used_sets()
    assert sum(len(u) for u in used) > 0 and (used.apply(len) <= np_o).all()


def test_losresult_on_constant_step_output(engine):
    """Lines of sight through a constant-step (bounce) Output: every step of every packet is
    a row (Output.py:434-449), several rows share one 'Index'."""
    from nexoclom_b200 import Output, LOSResult
    inputs = workload('Na.bounce.input')
    inputs.delete_files()
    out = Output(inputs, 1500, seed=2, keep_trajectory=True)
    assert out.trajectory_kept
    los = _synthetic_los(120, seed=6)
    sc = _FakeSCData(los, np.linspace(1.0, 2.0, 120))
    sc.species = 'Na'
    res = LOSResult(sc, inputs, dphi=Quantity(3.0, 'deg'))
    res.simulate_data_from_inputs(sc)
    P = Output.restore(out.filename).X
    assert len(P) > 1500 and P['Index'].nunique() <= 1500
    setup = RunSetup(inputs)
    rad_o, np_o, inc_o, _ = imaging.los_iteration(
        P.x.values, P.y.values, P.z.values, P.vy.values, P.frac.values, los,
        vrplanet=setup.vrplanet, dphi=np.radians(3.0), outeredge=float(inputs.options.outeredge),
        rp_cm=setup.radius_km * 1e5, gtables=setup.gtables([5891, 5897]))
    assert np.array_equal(res.npackets_los.values, np_o) and np_o.sum() > 500
    it = res._iterations[out.filename]
    assert it.included.sum() == len(np.unique(P['Index'].values[inc_o]))
    kR = rad_o * res.atoms_per_packet / 1e3 * float(res.sourcerate)
    nz = kR > 0
    assert np.max(np.abs(res.radiance.values[nz] - kR[nz]) / kR[nz]) < IMAGE_TOL
    inputs.delete_files()


def test_full_size_config1_properties(engine):
    """BASELINE configs[1] at its FULL size (1e7 packets, Na, radiation pressure + photo-loss,
    adaptive RK5(4)), checked through size-independent properties: two shards == one run
    bit for bit (state, step counts, image counts; the Philox counter is the global packet
    id and packets do not interact), the image is additive over shards, physical invariants
    hold for every packet, and a sample of the very same run matches the oracle."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    gt = setup.gtables([5891, 5897])
    engine.upload_gtables(gt)
    sp = setup.source_params(engine)
    n, half = 10_000_000, 5_000_000
    ip = _image_params(setup, 1, round_f32=1)

    def run(first, count):
        engine.init_state(sp, 0, first, count)
        X0 = engine.export_state()
        att, acc = engine.integrate_adaptive()
        X = engine.export_state()
        a, c = engine.export_stats()
        img, cnt = engine.image_accumulate(ip)
        assert att == int(a.sum()) and acc == int(c.sum())
        return X0, X, a, c, img, cnt

    X0, X, a, c, img, cnt = run(0, n)
    parts = [run(0, half), run(half, half)]
    assert np.array_equal(X0, np.concatenate([q[0] for q in parts], axis=1))
    assert np.array_equal(X, np.concatenate([q[1] for q in parts], axis=1))
    assert np.array_equal(a, np.concatenate([q[2] for q in parts]))
    assert np.array_equal(c, np.concatenate([q[3] for q in parts]))
    assert np.array_equal(cnt, parts[0][5] + parts[1][5])           # hit counts: exact
    s = parts[0][4] + parts[1][4]
    nz = s > 0
    assert np.array_equal(nz, img > 0)
    assert np.max(np.abs(img[nz] - s[nz]) / s[nz]) < 1e-12          # f64 sums, other order

    # every packet
    assert np.all(np.isfinite(X))
    assert np.all((c >= 0) & (c <= a)) and a.min() >= 0 and a.max() < 100000
    alive = X[7] > 0
    assert 0.001 < alive.mean() < 0.5
    # photo-loss only removes -- up to the step tolerance: a step across the shadow edge mixes
    # stages with and without loss, one tableau weight is negative (Q9 accepts such steps)
    print('largest frac increase', float((X[7] - X0[7]).max()))
    assert np.all(X[7] <= X0[7] * (1 + 10 * float(setup.params.resolution)))
    assert np.all(X[0][~alive] == 0)                                # Q8
    assert np.all((X[0] >= 0) & (X[0] <= X0[0]))
    r2 = X[1]**2 + X[2]**2 + X[3]**2
    assert np.all(r2[alive] > 1.0) and np.all(r2[alive] <= float(setup.params.outeredge))  # Q7
    assert 0.5 * n < cnt.sum() <= n            # packet_image counts dead packets too
    assert (img > 0).sum() <= alive.sum() and img.sum() > 0

    # a sample of the same run against the oracle (first / last packets and a stride)
    sel = np.unique(np.concatenate([np.arange(700), np.arange(n - 700, n),
                                    np.arange(0, n, 16661)]))
    Xo, a_o, c_o = tracking.integrate_adaptive(np.ascontiguousarray(X0[:, sel].T),
                                               oracle_constants(setup))
    par = state_parity(np.ascontiguousarray(X[:, sel].T), Xo)
    assert par['alive_mismatch'] == 0, par
    assert max(par['pos'], par['vel'], par['frac']) < STATE_TOL, par
    assert np.array_equal(a[sel], a_o) and np.array_equal(c[sel], c_o)


def test_full_size_config3_properties(engine):
    """BASELINE configs[2] (Na, temperature-dependent sticking + bounce + accommodation,
    constant 30 s step, 361 steps, image accumulated inside the integrator) on the LAST
    1.25e7-packet shard of the 1e8-packet run: the fused image is additive over sub-shards
    with exact counts (initial states and bounce deviates are keyed by the global packet
    id), and a slice at the top of the id range matches the oracle step by step."""
    import torch
    setup = RunSetup(workload('Na.bounce.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    sp = setup.source_params(engine)
    ip = _image_params(setup, 1, round_f32=1)
    ip.skip_dead = 1
    first, n = 87_500_000, 12_500_000

    def run(f, count):
        img = torch.zeros((800, 800), dtype=torch.float64, device='cuda')
        cnt = torch.zeros((800, 800), dtype=torch.int64, device='cuda')
        engine.init_state(sp, 0, f, count)
        _, ns, steps = engine.integrate_constant(seed=1, first_id=f, image_params=ip,
                                                 image_dev=img.data_ptr(),
                                                 counts_dev=cnt.data_ptr())
        torch.cuda.synchronize()
        return img.cpu().numpy(), cnt.cpu().numpy(), ns, steps

    img, cnt, ns, steps = run(first, n)
    a = run(first, 5_000_000)
    b = run(first + 5_000_000, n - 5_000_000)
    assert ns == 361 and steps == a[3] + b[3] and 0.05 * n * ns < steps <= n * ns
    assert np.array_equal(cnt, a[1] + b[1]) and cnt.sum() > 0
    ssum = a[0] + b[0]
    nz = ssum > 0
    assert np.array_equal(nz, img > 0)
    assert np.max(np.abs(img[nz] - ssum[nz]) / ssum[nz]) < 1e-10

    m, f = 400, 99_999_000
    engine.init_state(sp, 0, f, m)
    X0 = np.ascontiguousarray(engine.export_state().T)
    traj, ns2, st2 = engine.integrate_constant(seed=1, first_id=f, trajectory=True)
    ref, nsteps, natt = tracking.integrate_constant(
        X0, oracle_constants(setup), uniforms=initial_state.bounce_uniforms(1, f))
    assert ns2 == nsteps == 361 and st2 == int(natt.sum())
    assert np.array_equal(traj[:, 7, :] > 0, ref[:, 7, :] > 0)
    assert np.max(np.abs(traj - ref) / np.maximum(np.abs(ref), 1e-3)) < STATE_TOL


def test_full_size_los_properties(engine):
    """1e5 lines of sight x 1e7 packets (BASELINE's line-of-sight size) through the cell-grid
    path: additive over packet shards (each shard builds its own grid) with exact hit counts
    and `included` masks, and equal to the brute-force kernel on a subset of the lines."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    rng = np.random.default_rng(17)
    n, half, nlos = 10_000_000, 4_000_000, 100_000
    X = np.zeros((n, 8))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    X[:, 1:4] = d * (1 + 9 * rng.random(n)**2)[:, None]
    X[:, 5] = rng.normal(size=n) * 2 / setup.radius_km
    X[:, 7] = rng.random(n) * (rng.random(n) > 0.2)            # 20 % dead
    X = X.astype(np.float32).astype(np.float64)
    los = _synthetic_los(nlos, seed=4)
    sx = los[:, :3]
    dist = np.linalg.norm(sx, axis=1)
    ang = np.arccos(-(sx * los[:, 3:]).sum(axis=1) / dist)
    dist[ang > np.arcsin(1. / dist)] = 1e30
    lp = LosParams()
    lp.dphi, lp.outeredge = np.radians(1.0), 25.
    lp.vrplanet, lp.rp_cm, lp.quantity = setup.vrplanet, setup.radius_km * 1e5, 1
    lt = los.T.copy()

    def run(P, mode, lines=slice(None)):
        engine.import_state(P)
        engine.set_option('los_mode', mode)
        try:
            return engine.los_accumulate(np.ascontiguousarray(lt[:, lines]), dist[lines], lp)
        finally:
            engine.set_option('los_mode', 0)

    rad, npk, inc = run(X, 2)
    ra, na, ia = run(X[:half], 2)
    rb, nb, ib = run(X[half:], 2)
    assert npk.sum() > 1e7
    assert np.array_equal(npk, na + nb)
    assert np.array_equal(inc, np.concatenate([ia, ib]))
    ssum = ra + rb
    nz = ssum > 0
    assert np.array_equal(nz, rad > 0)
    assert np.max(np.abs(rad[nz] - ssum[nz]) / ssum[nz]) < 1e-10
    sub = slice(0, nlos, 211)
    r1, n1, i1 = run(X, 1, sub)                                   # brute force
    assert np.array_equal(n1, npk[sub])
    assert not np.any(i1 & ~inc)
    nz = r1 > 0
    assert np.max(np.abs(rad[sub][nz] - r1[nz]) / r1[nz]) < 1e-10


def test_pipelined_host_path_equals_resident_path(engine):
    """nx_integrate_adaptive_host (chunked H2D/compute pipeline) == import + integrate."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    X0 = initial_state.draw_x0(setup, 300_000, 6)[:, :8]
    engine.import_state(X0)
    att0, acc0 = engine.integrate_adaptive()
    ref = engine.export_state()
    a0, c0 = engine.export_stats()
    step0 = engine.export_step()
    for schedule in (1, 0):          # 1: one streaming class-ordered kernel; 0: per-chunk sort + kernel
        engine.set_option('schedule', schedule)
        for nchunks in (1, 3, 32):
            att, acc = engine.integrate_adaptive_host(X0, nchunks=nchunks)
            assert (att, acc) == (att0, acc0)
            assert np.array_equal(engine.export_state(), ref)
            a1, c1 = engine.export_stats()
            assert np.array_equal(a0, a1) and np.array_equal(c0, c1)
            assert np.array_equal(engine.export_step(), step0)
    engine.set_option('schedule', 1)


@pytest.mark.parametrize('n', [1, 127, 129, 4097, 50_001])
def test_streamed_host_path_ragged_and_dead_packets(engine, n):
    """Ragged segment sizes, packets that are dead or finished on arrival (they must
    pass through untouched with zero step counts), more segments than groups."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    X0 = initial_state.draw_x0(setup, n, 9)[:, :8].copy()
    X0[::7, 7] = 0.0                  # dead on arrival
    X0[3::11, 0] = 0.0                # no time left
    engine.import_state(X0)
    att0, acc0 = engine.integrate_adaptive()
    ref = engine.export_state()
    a0, c0 = engine.export_stats()
    engine.set_option('schedule', 1)
    for nchunks in (1, 5, 32):
        att, acc = engine.integrate_adaptive_host(X0, nchunks=nchunks)
        assert (att, acc) == (att0, acc0)
        got = engine.export_state()
        assert np.array_equal(got, ref)
        a1, c1 = engine.export_stats()
        assert np.array_equal(a0, a1) and np.array_equal(c0, c1)
    assert np.array_equal(got[:, ::7].T, X0[::7])             # untouched
    assert np.all(a1[::7] == 0)


def test_streamed_host_path_repeatable(engine):
    """The class-ordered schedule must hand out every packet exactly once: ten
    back-to-back runs give bit-identical states and step totals."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    X0 = initial_state.draw_x0(setup, 400_000, 12)[:, :8]
    ref, tot = None, None
    for rep in range(10):
        t = engine.integrate_adaptive_host(X0, nchunks=16)
        x = engine.export_state()
        if ref is None:
            ref, tot = x, t
        assert t == tot and np.array_equal(x, ref)


@pytest.mark.parametrize('sparams', [{'type': 'uniform'},
                                     {'type': 'uniform', 'exobase': '1.0', 'longitude': '1, 2',
                                      'latitude': '0, 1'},
                                     {'type': 'uniform', 'exobase': '1.0', 'longitude': '5, 1',
                                      'latitude': '-0.5, 0.25'},
                                     {'type': 'uniform', 'exobase': '1.0', 'longitude': '0, 0',
                                      'latitude': '0, 0'}])
def test_initial_state_distributions_ks(engine, sparams):
    """Statistical parity of the on-device sampler (the reference's own test of
    this function is a KS test: tests/unit_tests/Initial_state/test_spatial_distribution.py:95-143):
    longitude uniform on its (possibly wrapping) range, sin(latitude) uniform, sin(altitude)
    uniform, azimuth uniform, Maxwellian flux speeds, times uniform on [0, endtime]."""
    from scipy import stats
    from nexoclom_b200.input_classes import SpatialDist
    from nexoclom_b200.surfaceinteraction import thermal_speed_kms
    inputs = workload('Na.maxwellian.radpres.input')
    inputs.spatialdist = SpatialDist(sparams)
    setup = RunSetup(inputs)
    setup.upload(engine)
    n = 100_000
    engine.init_state(setup.source_params(engine), 2024, 0, n)
    X0 = engine.export_x0()
    t, x, y, z, vx, vy, vz, f, v, lon, lat, loct, alt, az = X0
    lon0, lon1 = (float(a) for a in inputs.spatialdist.longitude)
    lat0, lat1 = (float(a) for a in inputs.spatialdist.latitude)
    pmin = 1e-4
    if lon0 == lon1:
        assert np.all(lon == lon0) and np.all(lat == lat0)
    else:
        span = lon1 - lon0 if lon1 > lon0 else lon1 + 2 * np.pi - lon0
        u = ((lon - lon0) % (2 * np.pi)) / span
        assert u.max() <= 1 + 1e-12
        assert stats.kstest(u, 'uniform').pvalue > pmin
        s0, s1 = np.sin(lat0), np.sin(lat1)
        assert stats.kstest((np.sin(lat) - s0) / (s1 - s0), 'uniform').pvalue > pmin
    assert stats.kstest(np.sin(alt), 'uniform').pvalue > pmin
    assert stats.kstest(az / (2 * np.pi), 'uniform').pvalue > pmin
    assert stats.kstest(t / 50000., 'uniform').pvalue > pmin
    assert np.all(f == 1.0)
    assert np.allclose(np.sqrt(x**2 + y**2 + z**2), 1.0, rtol=1e-14)
    assert np.allclose(np.sqrt(vx**2 + vy**2 + vz**2), v, rtol=1e-13)
    assert np.allclose(loct, (lon * 12 / np.pi + 12) % 24, rtol=1e-14)
    # Maxwellian flux distribution f(v) ~ v^3 exp(-v^2/vth^2) on [0.1, 5 vth] km/s
    vth = thermal_speed_kms(1200., 'Na')
    grid = np.linspace(0.1, 5 * vth, 20001)
    pdf = grid**3 * np.exp(-grid**2 / vth**2)
    cdf = np.cumsum(pdf)
    cdf = (cdf - cdf[0]) / (cdf[-1] - cdf[0])
    vk = v * setup.radius_km
    assert stats.kstest(vk, lambda q: np.interp(q, grid, cdf)).pvalue > pmin
    # geometry: packets leave the surface outward, |v_radial| = v sin(alt)
    vr = (x * vx + y * vy + z * vz)
    assert np.allclose(vr, v * np.sin(alt), rtol=1e-9, atol=1e-18)


# ---------------------------------------------------------------------------
# K1 on the device against the reference's own draws: every source
# variant -- gaussian / sputtering / Maxwellian / user-table speeds, radial / isotropic / 2d
# directions, uniform band with wrapped longitudes, surface spot, longitude-only map -- runs
# through the sm_100a build of the transform on the deviates the unmodified reference
# functions drew (tests/golden/source_distribution.npz).
# ---------------------------------------------------------------------------
SOURCE_CASES = ['flat_iso', 'maxw_band', 'gauss_radial', 'sput_2d', 'spot_flat', 'lon1d_user']
SOURCE_COLS = {'time': 0, 'x': 1, 'y': 2, 'z': 3, 'vx': 4, 'vy': 5, 'vz': 6, 'v': 8,
               'longitude': 9, 'latitude': 10, 'local_time': 11, 'altitude': 12, 'azimuth': 13}


def _reference_deviates(tag, g, setup, sp):
    """The reference's recorded draws mapped onto K1's deviate columns (same bookkeeping as
    tests/test_oracle_products_golden.py::_replay)."""
    uni, legacy = list(g[f'{tag}_uniform']), list(g[f'{tag}_legacy'])
    normal = list(g[f'{tag}_normal'])
    n = len(g[f'{tag}_x'])
    d = {'u_time': uni.pop(0)}
    if sp.spatial_type == 0:
        d['u_sinlat'], d['u_lon'] = uni.pop(0), uni.pop(0)
    elif sp.spatial_type == 2:
        d['u_lon'] = legacy.pop(0)
    else:
        fmap, xa, ya = setup.sourcemap
        rounds = [(legacy[k], legacy[k + 1], legacy[k + 2]) for k in range(0, len(legacy), 3)]
        legacy = []
        d['lon'], d['lat'] = initial_state.pooled_rejection(fmap, xa, ya, sp.map_fmax, rounds, n)
        if sp.map_lat_is_sin:
            d['lat'] = np.arcsin(d['lat'])
    if sp.speed_type == 0:
        d['u_speed'] = uni.pop(0)
    elif sp.speed_type == 1:
        d['z_normal'] = normal.pop(0)
    else:
        d['u_speed'] = legacy.pop(0)
    if sp.angular_type == 1:
        d['u_alt'], d['u_az'] = uni.pop(0), uni.pop(0)
    elif sp.angular_type == 2:
        d['u_alt'] = uni.pop(0)
    assert not uni and not legacy and not normal
    return n, d


@pytest.mark.parametrize('tag', SOURCE_CASES)
def test_k1_transform_on_device_vs_reference(engine, tag):
    from common import source_case_input
    g = np.load(os.path.join(GOLDEN, 'source_distribution.npz'))
    setup = RunSetup(source_case_input(tag))
    setup.upload(engine)
    sp = setup.source_params(engine)           # uploads the speed / longitude tables, the map
    n, dev = _reference_deviates(tag, g, setup, sp)
    engine.init_state_deviates(sp, n, **dev)
    X0 = engine.export_x0().T
    for c, k in SOURCE_COLS.items():
        ref = g[f'{tag}_{c}']
        err = np.max(np.abs(X0[:, k] - ref)) / max(np.max(np.abs(ref)), 1e-300)
        assert err < 1e-14, (tag, c, err)
    assert np.all(X0[:, 7] == 1.0)
    # the freshly drawn packets are the current state (no second copy was written)
    assert np.array_equal(engine.export_state().T, X0[:, :8])


@pytest.mark.parametrize('tag', ['gauss_radial', 'sput_2d', 'lon1d_user', 'spot_flat'])
def test_k1_sampler_on_device_statistics(engine, tag):
    """The device SAMPLER (Philox draws + table / map lookups) for the variants the oracle
    comparison of test_init_state_vs_oracle does not reach: identical to the oracle driven by
    the same counters, odd packet counts included (paired 16-byte stores + scalar tail)."""
    from common import source_case_input
    setup = RunSetup(source_case_input(tag))
    setup.upload(engine)
    sp = setup.source_params(engine)
    for n in (30001, 4096):
        engine.init_state(sp, 17, 555, n)
        got = engine.export_x0().T
        ref = initial_state.draw_x0(setup, n, 17, first_id=555)
        scale = np.maximum(np.max(np.abs(ref), axis=0), 1e-300)
        assert np.max(np.abs(got - ref) / scale) < 1e-12, tag


def test_rewind_state_reruns_the_same_packets(engine):
    """nx_rewind_state: the integrators read X0 and write the state slab, so a run can be
    repeated on the resident initial state without any copy -- bit-identical results."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    n = 50001
    engine.init_state(setup.source_params(engine), 5, 0, n)
    x0 = engine.export_x0()[:8]
    a1 = engine.integrate_adaptive()
    s1 = engine.export_state()
    assert np.array_equal(engine.export_x0()[:8], x0)          # the input is immutable
    engine.rewind_state()
    assert np.array_equal(engine.export_state(), x0)
    a2 = engine.integrate_adaptive()
    assert a1 == a2 and np.array_equal(engine.export_state(), s1)
    # packets that need no integration (time <= resolution / frac == 0) are passed through
    X0 = x0.T.copy()
    X0[::7, 0] = 0.0
    X0[::11, 7] = 0.0
    engine.import_state(X0)
    engine.integrate_adaptive()
    got = engine.export_state().T
    idle = (X0[:, 0] <= 1e-4) | (X0[:, 7] <= 0)
    assert np.array_equal(got[idle], X0[idle])


# ---------------------------------------------------------------------------
# resident packet tables (device-side Output.save), the K3 row sink, privatised K4
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('skip_dead', [True, False])
@pytest.mark.parametrize('n', [1, 2047, 2049, 300_001])
def test_compact_state_equals_host_save(engine, n, skip_dead):
    """nx_compact_state == what Output.save does on the host (Output.py:522-543): rows with
    f64 frac > 0 in packet order, every column cast to float32, packet index kept."""
    rng = np.random.default_rng(n)
    X = rng.normal(size=(n, 8)) * 3
    X[:, 7] = np.where(rng.random(n) < 0.6, 0.0, rng.random(n))
    X[n // 2, 7] = 1e-40                          # > 0 in f64, underflows in f32: still kept
    engine.import_state(X)
    tab = engine.compact_state(skip_dead=skip_dead, round_f32=True)
    keep = (X[:, 7] > 0) if skip_dead else np.ones(n, dtype=bool)
    assert tab.n == int(keep.sum())
    cols, index = tab.export()
    assert np.array_equal(index, np.nonzero(keep)[0].astype(np.int32))
    for k, c in enumerate(('time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac')):
        assert cols[c].dtype == np.float32
        assert np.array_equal(cols[c], X[keep, k].astype(np.float32)), c
    # the bound table is what K4 reads: same image as importing the host-saved rows
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    engine.upload_gtables(setup.gtables([5891, 5897]))
    ip = _image_params(setup, 1, dims=(64, 64))
    engine.bind_packets(tab)
    img_a, cnt_a = engine.image_accumulate(ip, n=tab.n)
    engine.import_state(X[keep].astype(np.float32).astype(np.float64))
    img_b, cnt_b = engine.image_accumulate(ip)
    assert np.array_equal(cnt_a, cnt_b)
    assert np.allclose(img_a, img_b, rtol=1e-12, atol=0)
    tab.free()


def test_row_table_equals_dense_trajectory(engine):
    """K3's row sink == the rows of the reference's results[N, 8, nsteps] tensor that
    Output.save keeps (frac > 0, float32), identified by (packet, step)."""
    inputs = workload('Na.bounce.input')
    inputs.options.endtime = Quantity(900., 's')
    setup = RunSetup(inputs)
    setup.upload(engine)
    n, seed, first = 3001, 5, 1000
    sp = setup.source_params(engine)
    engine.init_state(sp, seed, first, n)
    traj, nsteps, steps = engine.integrate_constant(seed=seed + 1, first_id=first, trajectory=True)
    for skip_dead in (True, False):
        engine.init_state(sp, seed, first, n)
        tab, ns, steps2 = engine.integrate_constant_rows(seed=seed + 1, first_id=first,
                                                         skip_dead=skip_dead)
        assert ns == nsteps and steps2 == steps
        cols, index, step = tab.export(with_step=True)
        order = np.lexsort((step, index))
        dense = traj.transpose(0, 2, 1)                       # (n, nsteps, 8)
        # rows the reference writes: up to and including the step a packet ends at; the
        # rest of the dense tensor stays zero (Output.py:376 allocates, :392-421 fill)
        written = np.zeros((n, nsteps), dtype=bool)
        written[:, 0] = True
        alive = dense[:, :, 7] > 0
        for k in range(1, nsteps):
            written[:, k] = written[:, k - 1] & alive[:, k - 1] & (dense[:, k, 0] != 0)
        keep = alive if skip_dead else None
        if skip_dead:
            pk, st = np.nonzero(keep)
            assert tab.n == len(pk)
            assert np.array_equal(index[order], pk) and np.array_equal(step[order], st)
            for k, c in enumerate(('time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac')):
                assert np.array_equal(cols[c][order], dense[pk, st, k].astype(np.float32)), c
        else:
            assert tab.n >= int(alive.sum())
            live_rows = cols['frac'] > 0
            assert int(live_rows.sum()) == int(alive.sum())
        tab.free()


def test_los_over_row_table_equals_dense_rows(engine):
    """VERDICT r1 missing #2: lines of sight over a constant-step run WITHOUT its dense
    trajectory -- K5 over the bound row table == K5 / the oracle over the dense rows."""
    inputs = workload('Na.bounce.input')
    inputs.options.endtime = Quantity(1200., 's')
    setup = RunSetup(inputs)
    setup.upload(engine)
    gt = setup.gtables([5891, 5897])
    engine.upload_gtables(gt)
    n, seed = 4000, 8
    sp = setup.source_params(engine)
    engine.init_state(sp, seed, 0, n)
    traj, nsteps, _ = engine.integrate_constant(seed=seed, trajectory=True)
    rows = traj.transpose(0, 2, 1).reshape(-1, 8)
    rows = rows[rows[:, 7] > 0].astype(np.float32).astype(np.float64)
    los = _synthetic_los(150, seed=4)
    lp = LosParams()
    lp.dphi, lp.outeredge = np.radians(3.0), 25.0
    lp.vrplanet, lp.rp_cm = setup.vrplanet, setup.radius_km * 1e5
    lp.quantity, lp.round_f32, lp.skip_dead = 1, 0, 0
    rad_o, np_o, _, dist = imaging.los_iteration(
        rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 5], rows[:, 7], los, vrplanet=setup.vrplanet,
        dphi=np.radians(3.0), outeredge=25.0, rp_cm=setup.radius_km * 1e5, gtables=gt)
    engine.init_state(sp, seed, 0, n)
    tab, _, _ = engine.integrate_constant_rows(seed=seed, skip_dead=True)
    assert tab.n == len(rows)
    engine.bind_packets(tab)
    for mode in (1, 2):
        engine.set_option('los_mode', mode)
        rad, npk, inc = engine.los_accumulate(los.T.copy(), dist, lp, n=tab.n)
        assert np.array_equal(npk, np_o) and np_o.sum() > 300
        nz = rad_o > 0
        assert np.max(np.abs(rad[nz] - rad_o[nz]) / rad_o[nz]) < IMAGE_TOL
    engine.set_option('los_mode', 0)
    engine.bind_packets(None)
    tab.free()


@pytest.mark.parametrize('quantity', [0, 1])
def test_privatised_image_counts_equal_global(engine, quantity):
    """K4 with the shared-memory count tile (image_mode 2) == the all-global kernel: counts
    bit-exact (also for packets outside the tile and off the image), sums to rounding."""
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    setup.upload(engine)
    engine.upload_gtables(setup.gtables([5891, 5897]))
    rng = np.random.default_rng(12)
    n = 700_001
    X = np.zeros((n, 8))
    X[:, 1:4] = rng.normal(size=(n, 3)) * np.where(rng.random(n) < 0.7, 1.0, 3.5)[:, None]
    X[:, 5] = rng.normal(size=n) * 2e-4
    X[:, 7] = np.where(rng.random(n) < 0.2, 0.0, rng.random(n))
    engine.import_state(X)
    for dims, view in (((800, 800), (0.0, np.pi / 2)), ((160, 120), (0.7, 0.3))):
        ip = _image_params(setup, quantity, view=view, dims=dims, round_f32=1)
        ip.skip_dead = 1
        engine.set_option('image_mode', 1)
        img_g, cnt_g = engine.image_accumulate(ip)
        engine.set_option('image_mode', 2)
        img_t, cnt_t = engine.image_accumulate(ip)
        engine.set_option('image_mode', 0)
        assert cnt_g.sum() > 1e5 and np.array_equal(cnt_g, cnt_t)
        nz = img_g > 0
        assert np.array_equal(nz, img_t > 0)
        assert np.max(np.abs(img_t[nz] - img_g[nz]) / img_g[nz]) < 1e-10


def test_pinned_result_buffers_are_recycled(engine):
    """ModelImage's planes and the line-of-sight results live in pooled page-locked arrays: a
    buffer returns to the pool when the last array viewing it dies and is handed out again."""
    import gc
    from nexoclom_b200.engine import _pinned_pool
    pool = _pinned_pool(engine.lib)
    a = pool.array((800, 800))
    addr = a.ctypes.data
    v = a[10:20]                       # a view keeps the buffer alive
    del a
    gc.collect()
    b = pool.array((800, 800))
    assert b.ctypes.data != addr
    del v, b
    gc.collect()
    c = pool.array((800, 800))
    d = pool.array((800, 800))
    assert addr in (c.ctypes.data, d.ctypes.data)
    c[:] = 1.0                          # writable, ordinary ndarray semantics
    assert float(c.sum()) == 640000.0
