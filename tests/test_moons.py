"""Moons (BASELINE configs[3], Na from Io at Jupiter) -- an EXTENSION: the reference asserts for
planets with moons (Output.py:153-155; state.py:12 "does not do moons yet"), so there is no
reference behaviour to match.  The specification is SURVEY.md section 8 a-note / DESIGN.md
section 8: circular prograde equatorial orbits at the phases geometry.phi
(docs/nexoclom/inputfiles.rst:62-77), gravity sum_obj GM (x - x_obj)/r_obj^3 (state.py:5-10).
Checked here: (1) the physics itself through the Jacobi integral of the restricted three-body
problem, (2) the kernels' code (host build) against the NumPy restatement, (3) on the GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from common import workload, oracle_constants, state_parity
from nexoclom_b200.runsetup import RunSetup
from oracle import initial_state, tracking


def jacobi(X, tau0, moon, GM):
    """J = v^2/2 + U - omega L_z in the planet-centred frame, U = GM/r + GM_m/|r - r_m|
    - GM_m (r . r_m)/a^3 (the last term is the potential of the frame's own acceleration)."""
    mx, my = tracking.moon_xy(moon, X[:, 0])
    x, y, z, vx, vy, vz = (X[:, k] for k in range(1, 7))
    r = np.sqrt(x * x + y * y + z * z)
    d = np.sqrt((x - mx)**2 + (y - my)**2 + z * z)
    U = GM / r + moon['GM'] / d - moon['GM'] * (x * mx + y * my) / moon['a']**3
    return 0.5 * (vx * vx + vy * vy + vz * vz) + U - moon['omega'] * (x * vy - y * vx)


def _gravity_only_setup():
    setup = RunSetup(workload('Na.Io.Jupiter.input'))
    rc = oracle_constants(setup)
    rc.radpres = False
    rc.photo = None
    rc.outeredge = 1e30
    return setup, rc


def test_io_start_geometry():
    """Packets start on Io's surface, moving with Io."""
    setup = RunSetup(workload('Na.Io.Jupiter.input'))
    m = setup.moons[0]
    X0 = initial_state.draw_x0(setup, 2000, 5)
    mx, my = tracking.moon_xy(m, X0[:, 0])
    d = np.sqrt((X0[:, 1] - mx)**2 + (X0[:, 2] - my)**2 + X0[:, 3]**2)
    assert np.allclose(d, m['radius'], rtol=1e-12)
    phi = m['phi'] - m['omega'] * X0[:, 0]
    vorb = m['a'] * m['omega']
    vrel = np.stack([X0[:, 4] + vorb * np.cos(phi), X0[:, 5] + vorb * np.sin(phi), X0[:, 6]], axis=1)
    assert np.allclose(np.linalg.norm(vrel, axis=1), X0[:, 8], rtol=1e-12)
    # ejected outward: relative velocity has a positive component along the local normal
    nrm = np.stack([X0[:, 1] - mx, X0[:, 2] - my, X0[:, 3]], axis=1) / m['radius']
    assert np.all(np.sum(nrm * vrel, axis=1) > -1e-18)
    # the sub-planet point (longitude 0) faces Jupiter
    k = np.argmin(np.abs(X0[:, 9]) + np.abs(X0[:, 10]))
    to_planet = -np.array([mx[k], my[k], 0.0]) / m['a']
    assert np.dot(nrm[k], to_planet) > 0.9


def test_jacobi_integral_is_conserved():
    """Gravity of Jupiter + Io only: the Jacobi integral of every packet that neither hits
    nor escapes stays constant to the integrator's tolerance (reference analogue:
    tests/unit_tests/particle_tracking/test_gravity.py:46-55, energy conservation)."""
    setup, rc = _gravity_only_setup()
    m = setup.moons[0]
    X0 = initial_state.draw_x0(setup, 48, 11)[:, :8]
    X0[:, 0] = 12000.0
    J0 = jacobi(X0, None, m, rc.GM)
    X, att, acc = tracking.integrate_adaptive(X0, rc)
    ok = (X[:, 7] > 0)
    assert ok.sum() > 30
    J1 = jacobi(X, None, m, rc.GM)                # X[:, 0] is the time remaining at the end
    # the step is sized by a FIRST-order error estimate (quirk Q1), far smaller than a
    # fifth-order method needs: the integral is conserved to rounding
    drift = np.max(np.abs(J1[ok] - J0[ok]) / np.abs(J0[ok]))
    assert drift < 1e-10, drift
    # a wrong sign or frame in the moon terms breaks it: drop the indirect term only
    bad = dict(m)
    mx, my = tracking.moon_xy(m, X[:, 0])
    Jbad = J1 + m['GM'] * (X[:, 1] * mx + X[:, 2] * my) / m['a']**3
    assert np.max(np.abs(Jbad[ok] - J0[ok]) / np.abs(J0[ok])) > 1e-6
    # and the moon matters: without it the same packets end somewhere else
    rc0 = oracle_constants(setup)
    rc0.radpres, rc0.photo, rc0.outeredge, rc0.moons = False, None, 1e30, []
    Xn, _, _ = tracking.integrate_adaptive(X0, rc0)
    assert np.median(np.linalg.norm(Xn[ok, 1:4] - X[ok, 1:4], axis=1)) > 1e-5


@pytest.fixture(scope='module')
def hc():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_hostcheck',
                        'libnexo_hostcheck.so')
    if not os.path.exists(path):
        import subprocess
        subprocess.run(['sh', os.path.join(os.path.dirname(path), 'build.sh')], check=True)
    return C.CDLL(path)


@pytest.mark.parametrize('strict', [1, 0])
def test_kernel_code_matches_oracle_with_moons(hc, strict):
    """The kernels' own physics (csrc/*.cuh compiled for the host; strict = generic path,
    0 = the single-moon fast path with angle-addition moon phases) against the NumPy
    restatement: same accept / reject sequences, states within 1e-10."""
    from test_hostcheck import run_adaptive
    setup = RunSetup(workload('Na.Io.Jupiter.input'))
    X0 = initial_state.draw_x0(setup, 200, 3)[:, :8]
    X0[:, 0] *= 0.2
    Xo, a_o, c_o = tracking.integrate_adaptive(X0, oracle_constants(setup))
    Xh, a_h, c_h, _ = run_adaptive(hc, setup, X0, strict)
    par = state_parity(Xh, Xo)
    assert par['alive_mismatch'] == 0
    assert np.array_equal(a_h, a_o) and np.array_equal(c_h, c_o)
    assert max(par['pos'], par['vel'], par['frac']) < 1e-10


def test_init_state_matches_oracle_with_moon_start(hc):
    from nexoclom_b200._lib import dptr
    setup = RunSetup(workload('Na.Io.Jupiter.input'))
    sp = setup.source_params(None)
    n = 500
    out = np.zeros((n, 14))
    hc.hc_init_state(C.c_long(n), C.byref(sp), C.c_ulonglong(9), C.c_ulonglong(77), None, None,
                     None, None, C.c_int(0), dptr(out))
    ref = initial_state.draw_x0(setup, n, 9, first_id=77)
    assert np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('strict', [False, True])
def test_gpu_moons_vs_oracle(engine, strict):
    """K1 + K2 through the C ABI on the Io / Jupiter workload against the oracle (fast
    single-moon kernels and the generic ones)."""
    setup = RunSetup(workload('Na.Io.Jupiter.input'), strict_math=strict)
    setup.upload(engine)
    n = 1500
    engine.init_state(setup.source_params(engine), 4, 0, n)
    X0 = engine.export_x0()[:8].T.copy()
    ref0 = initial_state.draw_x0(setup, n, 4)[:, :8]
    assert np.max(np.abs(X0 - ref0) / np.maximum(np.abs(ref0), 1e-3)) < 1e-12
    X0[:, 0] *= 0.3
    engine.import_state(X0)
    att, acc = engine.integrate_adaptive()
    Xg = engine.export_state().T
    a_g, c_g = engine.export_stats()
    Xo, a_o, c_o = tracking.integrate_adaptive(X0, oracle_constants(setup))
    par = state_parity(Xg, Xo)
    assert par['alive_mismatch'] == 0, par
    assert np.array_equal(a_g, a_o) and np.array_equal(c_g, c_o)
    assert max(par['pos'], par['vel'], par['frac']) < 1e-8, par
    # the streamed host-buffer path gives the same answer
    t = engine.integrate_adaptive_host(X0, nchunks=4)
    assert t == (att, acc) and np.array_equal(engine.export_state().T, Xg)
