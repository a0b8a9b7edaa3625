"""Live cross-check of oracle/ against the UNMODIFIED reference functions
(build container only: needs /root/reference; skipped on the GPU box, where the
committed tests/golden fixtures stand in)."""
import numpy as np
import pandas as pd
import pytest

import refimport
from common import workload, oracle_constants
from nexoclom_b200.runsetup import RunSetup
from nexoclom_b200.units import Quantity
from oracle import tracking, initial_state

pytestmark = pytest.mark.skipif(not refimport.available(), reason='reference tree not present')
COLS = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']


@pytest.fixture(scope='module')
def ref():
    return refimport.install()


def _fake(setup, seed=0):
    import make_golden
    return make_golden.fake_from_setup(setup, seed)


def test_rk5_bit_exact(ref):
    setup = RunSetup(workload('Na.maxwellian.radpres.input'))
    fo = _fake(setup)
    rng = np.random.default_rng(0)
    x0 = initial_state.draw_x0(setup, 3000, 1)[:, :8]
    x0[:, 1:4] *= (1 + 2 * rng.random(3000))[:, None]
    h = np.minimum(x0[:, 0] + 1, 10**rng.uniform(0, 3, 3000))
    a, b = ref.rk5(fo, x0.copy(), h)
    c, d = tracking.dp_step(x0, h, oracle_constants(setup))
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_adaptive_driver_bit_exact(ref):
    setup = RunSetup(workload('Ca.isotropic.flat.input'))
    fo = _fake(setup)
    x0 = initial_state.draw_x0(setup, 60, 2)[:, :8]
    fo.X = pd.DataFrame(x0.copy(), columns=COLS)
    fo.X['lossfrac'] = 0.
    fo.npackets = len(x0)
    ref.Output.variable_step_size_driver(fo)
    X, _, _ = tracking.integrate_adaptive(x0, oracle_constants(setup))
    assert np.array_equal(fo.X[COLS].values, X)


def test_constant_driver_bounce_bit_exact(ref):
    inputs = workload('Na.bounce.input')
    inputs.options.endtime = Quantity(900., 's')
    setup = RunSetup(inputs)
    fo = _fake(setup, seed=4)
    x0 = initial_state.draw_x0(setup, 300, 3)[:, :8]
    fo.X0 = pd.DataFrame(x0.copy(), columns=COLS)
    fo.npackets, fo.totalsource = len(x0), float(len(x0))
    ref.Output.constant_step_size_driver(fo)
    gen = np.random.default_rng(4)
    traj, nsteps, _ = tracking.integrate_constant(
        x0, oracle_constants(setup),
        uniforms=lambda ct, idx: (gen.random(len(idx)), gen.random(len(idx)),
                                  gen.random(len(idx))))
    got = fo.X[COLS].values.reshape(len(x0), nsteps, 8).transpose(0, 2, 1)
    assert np.array_equal(got, traj)


def test_host_tables_live_vs_reference():
    """The port's SurfaceInteraction table and planet_dist against the reference classes
    executed here (tools/reftables.py), beyond the committed sample points."""
    import reftables
    from nexoclom_b200.solarsystem import planet_dist
    from nexoclom_b200.surfaceinteraction import SurfaceInteraction
    try:
        inputs = workload('Na.bounce.input')
        inputs.geometry.taa = Quantity(0.7, 'rad')
        theirs = reftables.surface_interaction(inputs)
        mine = SurfaceInteraction(inputs)
        assert np.array_equal(mine.probgrid, np.asarray(theirs.probgrid))
        rng = np.random.default_rng(5)
        T = rng.uniform(100., float(np.max(mine.temperature)), 500)
        P = rng.random(500)
        assert np.array_equal(mine.v_interp(T, P), theirs.v_interp(T, P))
        lon, lat = rng.random(500) * 2 * np.pi, np.arcsin(rng.random(500) * 2 - 1)
        assert np.array_equal(mine.stickcoef(lon, lat), theirs.stickcoef(lon, lat))
        for taa in rng.random(20) * 2 * np.pi:
            r, v = planet_dist('Mercury', float(taa))
            assert (float(r.value), float(v.value)) == reftables.planet_dist('Mercury', taa)
    finally:
        reftables.purge()
