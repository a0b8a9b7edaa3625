"""oracle/ replayed against golden vectors produced by the UNMODIFIED reference
functions (tools/make_golden.py).  Bit-exact on the machine that generated them;
elsewhere NumPy's SIMD pow/exp/log may differ in the last ulp, so the gate is
1e-12 with step-count equality for the drivers."""
import os

import numpy as np
import pytest

from common import GOLDEN, workload, oracle_constants
from nexoclom_b200.runsetup import RunSetup
from nexoclom_b200.units import Quantity
from oracle import tracking, initial_state

WL = {'na': 'Na.maxwellian.radpres.input', 'ca': 'Ca.isotropic.flat.input',
      'grav': 'Gravity.input'}


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize('tag', ['na', 'ca', 'grav'])
def test_rk5_and_state(tag):
    g = np.load(os.path.join(GOLDEN, 'rk5_steps.npz'))
    rc = oracle_constants(RunSetup(workload(WL[tag])))
    acc, rate = tracking.rhs(g[f'{tag}_x0'], rc)
    assert rel(acc, g[f'{tag}_accel']) < 1e-13
    assert np.array_equal(rate, g[f'{tag}_rate'])
    res, delta = tracking.dp_step(g[f'{tag}_x0'], g[f'{tag}_h'], rc)
    scale = np.abs(g[f'{tag}_result']).max(axis=0)
    assert np.max(np.abs(res - g[f'{tag}_result']) / scale) < 1e-13
    assert np.max(np.abs(delta - g[f'{tag}_delta'])) < 1e-13 * np.abs(g[f'{tag}_delta']).max()


@pytest.mark.parametrize('tag', ['na', 'ca'])
def test_adaptive_driver(tag):
    g = np.load(os.path.join(GOLDEN, 'adaptive_driver.npz'))
    rc = oracle_constants(RunSetup(workload(WL[tag])))
    X, att, acc, step = tracking.integrate_adaptive(g[f'{tag}_x0'], rc, return_step=True)
    ref = g[f'{tag}_final']
    assert np.array_equal(X[:, 7] > 0, ref[:, 7] > 0)
    alive = ref[:, 7] > 0
    assert rel(X[alive, 1:8], ref[alive, 1:8]) < 1e-10
    assert np.max(np.abs(step - g[f'{tag}_step']) / g[f'{tag}_step']) < 1e-10


@pytest.mark.parametrize('tag, wl', [('tdep', 'Na.bounce.input'),
                                     ('c05', 'Na.bounce.stick05.input'),
                                     ('grav', 'Gravity.input')])
def test_constant_driver_with_bounce(tag, wl):
    g = np.load(os.path.join(GOLDEN, 'constant_driver.npz'))
    inputs = workload(wl)
    inputs.options.endtime = Quantity(float(g[f'{tag}_endtime']), 's')
    setup = RunSetup(inputs)
    rc = oracle_constants(setup)
    gen = np.random.default_rng(int(g[f'{tag}_seed']))

    def uniforms(ct, idx):      # the reference's draw order: sinalt, az, probability
        k = len(idx)
        return gen.random(k), gen.random(k), (gen.random(k) if rc.accomfactor != 0 else None)

    traj, nsteps, _ = tracking.integrate_constant(g[f'{tag}_x0'], rc, uniforms=uniforms)
    ref = g[f'{tag}_traj']
    assert traj.shape == ref.shape
    assert np.array_equal(traj[:, 7, :] > 0, ref[:, 7, :] > 0)
    assert np.max(np.abs(traj - ref)) < 1e-9


def test_surface_temperature_and_rebound():
    g = np.load(os.path.join(GOLDEN, 'surface.npz'))
    ts = tracking.surface_temperature(float(g['taa']), g['lon'], g['lat'])
    assert rel(ts, g['tsurf']) < 1e-14
    gen = np.random.default_rng(int(g['seed']))
    n = len(g['lon'])
    d = tracking.local_frame_direction(g['pos'], gen.random(n), 2 * np.pi * gen.random(n))
    assert np.max(np.abs(d - g['direction'])) < 1e-14


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors file of the
    Random123 distribution)."""
    f = initial_state.philox4x32_10
    out = f(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    out = f(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = f(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_config0_named_size_fixture_is_the_oracles():
    """tests/golden/config0_1e5.npz (tools/make_golden_config0.py): the oracle replayed on a
    slice of the 1e5 packets reproduces the stored step counts and final states."""
    g = np.load(os.path.join(GOLDEN, 'config0_1e5.npz'))
    n, seed, stride = int(g['n']), int(g['seed']), int(g['stride'])
    setup = RunSetup(workload(WL['ca']))
    X0 = initial_state.draw_x0(setup, n, seed)[:, :8].astype(np.float32).astype(np.float64)
    if not np.array_equal(X0.sum(axis=0), g['x0_sums']):
        pytest.skip('initial state not bit-reproducible on this host (NumPy SIMD sin / cos)')
    pick = np.arange(0, n, stride)[:240]
    Xo, att, acc = tracking.integrate_adaptive(X0[pick], oracle_constants(setup))
    assert np.array_equal(att, g['attempted'][pick]) and np.array_equal(acc, g['accepted'][pick])
    ref = g['final_subset'][:240]
    assert np.max(np.abs(Xo - ref)) < 1e-12 * max(1.0, np.abs(ref).max())
