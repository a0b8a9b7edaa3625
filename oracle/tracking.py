"""Oracle: packet tracking (RHS, Dormand-Prince step, the two drivers, bounce).

TEST INFRASTRUCTURE -- see oracle/__init__.py for who may import this.

NumPy restatement of
  reference nexoclom/particle_tracking/state.py:17-74          -> rhs()
  reference nexoclom/particle_tracking/rk5.py:5-54             -> dp_step()
  reference nexoclom/particle_tracking/Output.py:221-366       -> integrate_adaptive()
  reference nexoclom/particle_tracking/Output.py:368-455       -> integrate_constant()
  reference nexoclom/particle_tracking/bouncepackets.py:5-100  -> bounce()
  reference nexoclom/initial_state/surface_temperature.py:4-19 -> surface_temperature()
with plain arrays instead of pandas/astropy objects.  Floating-point operation
ORDER follows the reference exactly (e.g. ``(h*a_ni)*k_i`` accumulated from 0 and
the base state added last; norms as ``sqrt((x^2+y^2)+z^2)``; ``r**3`` through
NumPy's pow) so results are bit-identical to the reference functions on the same
machine -- checked in tests/test_oracle_vs_reference.py and against
tests/golden/rk5_*.npz.

Packet state columns: 0 time-remaining, 1-3 x y z [R_p], 4-6 vx vy vz [R_p/s],
7 frac.  GM is negative (R_p^3/s^2); Sun at -y; shadow = (x^2+z^2 <= 1 and y >= 0).
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

# Dormand-Prince 5(4) tableau (reference rk5.py:5-18).
DP_C = (0., 0.2, 0.3, 0.8, 8. / 9., 1., 1.)
DP_B = (35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 0.)
DP_BS = (5179. / 57600., 0., 7571. / 16695., 393. / 640., -92097. / 339200.,
         187. / 2100., 1. / 40.)
DP_BD = tuple(np.array(DP_B) - np.array(DP_BS))
DP_A = (
    (),
    (0.2,),
    (3. / 40., 9. / 40.),
    (44. / 45., -56. / 15., 32. / 9.),
    (19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729.),
    (9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656.),
    DP_B[:6],
)


@dataclass
class RunConstants:
    """Per-run scalars and tables (what reference Output.__init__ :102-133 builds)."""
    GM: float
    vrplanet: float = 0.0
    gravity: bool = True
    radpres: bool = True
    radpres_v: Optional[np.ndarray] = None     # R_p/s, ascending
    radpres_a: Optional[np.ndarray] = None     # R_p/s^2
    lifetime: float = 0.0                      # > 0: constant e-folding time [s]
    photo: Optional[float] = None              # loss_info.photo [1/s] (lifetime <= 0)
    resolution: float = 1e-4
    outeredge: float = 1e30
    step_size: float = 0.0                     # 0 -> adaptive
    endtime: float = 0.0
    # surface interaction
    sticktype: str = 'constant'
    stickcoef: float = 1.0
    accomfactor: Optional[float] = None
    A: Tuple[float, float, float] = (1.57014, -0.006262, 0.1614157)
    taa: float = 0.0
    planet_radius_km: float = 2440.53
    v_interp: object = None                    # callable (T, prob) -> km/s
    extra: dict = field(default_factory=dict)
    # EXTENSION (no reference code: it asserts for planets with moons, Output.py:153-155):
    # moons on circular prograde equatorial orbits, dicts with GM, a, omega, phi, radius in
    # R_p units; position at time-remaining tau = a (-sin phi, cos phi, 0), phi = phi0 - omega tau
    moons: list = field(default_factory=list)


def moon_xy(m, tau):
    phi = m['phi'] - m['omega'] * tau
    return -m['a'] * np.sin(phi), m['a'] * np.cos(phi)


def _out_of_shadow(x):
    rho = np.sqrt(x[:, 1] * x[:, 1] + x[:, 3] * x[:, 3])
    return (rho > 1) | (x[:, 2] < 0)


def rhs(x, rc):
    """Acceleration (M,3) and loss rate (M,) -- reference state.py:17-74."""
    m = x.shape[0]
    acc = np.zeros((m, 3))
    if rc.gravity:
        r = np.sqrt((x[:, 1] * x[:, 1] + x[:, 2] * x[:, 2]) + x[:, 3] * x[:, 3])
        r3 = r**3
        acc = rc.GM * x[:, 1:4] / r3[:, np.newaxis]
        for mo in rc.moons:       # extension: direct term + the planet's own acceleration
            mx, my = moon_xy(mo, x[:, 0])
            dx, dy, z = x[:, 1] - mx, x[:, 2] - my, x[:, 3]
            d2 = dx * dx + dy * dy + z * z
            di = 1.0 / np.sqrt(d2)
            gd = mo['GM'] * di * di * di
            ai = 1.0 / mo['a']
            gi = mo['GM'] * ai * ai * ai
            acc = acc + np.stack([gd * dx + gi * mx, gd * dy + gi * my, gd * z], axis=1)
    if rc.radpres:
        lit = _out_of_shadow(x)
        vv = x[:, 5] + rc.vrplanet
        arad_y = np.interp(vv, rc.radpres_v, rc.radpres_a) * lit
        acc = acc + np.stack([np.zeros(m), arad_y, np.zeros(m)], axis=1)
    else:
        acc = acc + np.zeros((m, 3))
    if rc.lifetime > 0:
        rate = np.ones(m) / rc.lifetime
    elif rc.photo is not None:
        rate = rc.photo * _out_of_shadow(x)
    else:
        rate = np.zeros(m)
    return acc, rate


def dp_step(X0, h, rc, want_error=True):
    """One Dormand-Prince step -- reference rk5.py:21-54.

    Returns (Xnew (M,8), delta (M,8) or None).  ``delta`` uses stages 0..5 only
    (the 7th derivative is never evaluated -- quirk Q1)."""
    m = X0.shape[0]
    stage = [X0.copy()]
    stage[0][:, 7] = np.log(stage[0][:, 7])
    acc, rate = [], []
    hcol = h[:, np.newaxis]
    for n in range(6):
        a_n, r_n = rhs(stage[n], rc)
        acc.append(a_n)
        rate.append(r_n)
        nxt = np.zeros((m, 8))
        nxt[:, 0] = -h * DP_C[n + 1]
        for i in range(n + 1):
            a_ni = DP_A[n + 1][i]
            nxt[:, 1:4] += hcol * a_ni * stage[i][:, 4:7]
            nxt[:, 4:7] += hcol * a_ni * acc[i]
            nxt[:, 7] -= h * a_ni * rate[i]
        nxt += stage[0]
        stage.append(nxt)

    delta = None
    if want_error:
        delta = np.zeros((m, 8))
        for i in range(6):
            delta[:, 1:4] += DP_BD[i] * stage[i][:, 4:7]
            delta[:, 4:7] += DP_BD[i] * acc[i]
            delta[:, 7] += DP_BD[i] * rate[i]
        delta = np.abs(hcol * delta)

    result = stage[6]
    result[:, 7] = np.exp(result[:, 7])
    return result, delta


def integrate_adaptive(X0, rc, step0=1000., max_iter=None, return_step=False):
    """Adaptive driver -- reference Output.py:221-366.

    X0: (N,8).  Returns (X (N,8) final, attempted (N,) int64, accepted (N,) int64).
    Quirks kept: accepted steps never persist a grown step (Q3); 'no error' steps
    are REJECTED with the step x9.5 (Q4); escape compares r^2 with outeredge (Q7)."""
    assert rc.sticktype == 'constant' and rc.stickcoef == 1., 'Not set up'   # Q6
    safety, shrink = 0.95, -0.25
    res = rc.resolution
    resx, resv, resf = res, 0.1 * res, res

    X = np.array(X0, dtype=np.float64, copy=True)
    n = X.shape[0]
    step = np.zeros(n) + step0
    attempted = np.zeros(n, dtype=np.int64)
    accepted = np.zeros(n, dtype=np.int64)
    live = (X[:, 0] > res) & (X[:, 7] > 0.)
    it = 0
    while live.any():
        idx = np.nonzero(live)[0]
        cur = X[idx]
        h = np.minimum(cur[:, 0], step[idx])
        assert np.all(h > 0), 'Bad step size'
        nxt, delta = dp_step(cur, h, rc, want_error=True)

        scale = np.empty((len(idx), 8))
        scale[:, 0] = 1.0
        scale[:, 1:4] = resx + np.abs(nxt[:, 1:4]) * resx
        scale[:, 4:7] = resv + np.abs(nxt[:, 4:7]) * resv
        scale[:, 7] = resf + np.abs(nxt[:, 7]) * resf
        ratio = delta / scale
        ratio[:, 0] = delta[:, 0]
        errmax = ratio.max(axis=1)
        assert np.all(np.isfinite(errmax)), '\n\tInfinite values of emax'
        assert not np.any((nxt[:, 7] < 0) & (errmax < 1)), (
            'Found new values of frac that are negative')

        errmax[(nxt[:, 7] - cur[:, 7] > scale[:, 7]) & (errmax > 1)] = 1.1   # Q9
        noerr = errmax < 1e-7
        errmax[noerr] = 1
        htried = h.copy()
        htried[noerr] *= 10
        good = errmax < 1.0
        bad = ~good
        attempted[idx] += 1

        if good.any():
            gn = nxt[good]
            r2 = (gn[:, 1]**2 + gn[:, 2]**2) + gn[:, 3]**2
            gn[r2 < 1, 7] = 0
            for mo in rc.moons:                       # extension: impact on a moon
                mx, my = moon_xy(mo, gn[:, 0])
                dx, dy = gn[:, 1] - mx, gn[:, 2] - my
                gn[dx * dx + dy * dy + gn[:, 3] * gn[:, 3] < mo['radius']**2, 7] = 0
            gn[r2 > rc.outeredge, 7] = 0
            gn[gn[:, 7] < 1e-10, 7] = 0.
            gn[gn[:, 7] == 0, 0] = 0
            X[idx[good]] = gn
            accepted[idx[good]] += 1
        if bad.any():
            old = htried[bad]
            new = safety * old * errmax[bad]**shrink
            assert not np.any(np.isclose(new, old))
            assert np.all(np.isfinite(new)), '\n\tInfinite values of step_size'
            step[idx[bad]] = np.maximum(new, 0.1 * old)

        live = (X[:, 0] > res) & (X[:, 7] > 0.)
        it += 1
        if max_iter is not None and it >= max_iter:
            break
    if return_step:
        return X, attempted, accepted, step
    return X, attempted, accepted


# --------------------------------------------------------------------------
# surface interaction
# --------------------------------------------------------------------------
def surface_temperature(taa, longitude, latitude, t0=100., n=.25):
    """Mercury surface temperature -- reference surface_temperature.py:4-19."""
    t1 = 600. + 125 * (np.cos(taa) - 1) / 2.
    t_surf = np.zeros_like(longitude) + t0
    day = (longitude <= np.pi / 2) | (longitude >= 3 * np.pi / 2)
    t_surf[day] = t0 + t1 * np.abs(np.cos(longitude[day]) * np.cos(latitude[day]))**n
    return t_surf


def stick_coefficient(rc, lon, lat):
    """T-dependent sticking -- reference SurfaceInteraction.py:15-20."""
    tsurf = surface_temperature(rc.taa, lon, lat)
    coef = rc.A[0] * np.exp(rc.A[1] * tsurf) + rc.A[2]
    coef[coef > 1.] = 1.
    coef[coef < 0.] = 0.
    return coef


def local_frame_direction(pos, sinalt, az):
    """Unit emission direction at surface point ``pos`` (K,3) for altitude
    asin(sinalt) and azimuth az -- reference bouncepackets.py:5-36 (same frame as
    source_distribution.py:229-252; 'v_tan0' multiplies NORTH, quirk Q18)."""
    k = pos.shape[0]
    alt = np.arcsin(sinalt)
    v_rad = np.sin(alt)
    v_tan0 = np.cos(alt) * np.cos(az)
    v_tan1 = np.cos(alt) * np.sin(az)
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    rad = pos / np.sqrt((x * x + y * y) + z * z)[:, np.newaxis]
    east = np.stack([y, -x, np.zeros(k)], axis=1)
    east = east / np.sqrt((east[:, 0]**2 + east[:, 1]**2) + east[:, 2]**2)[:, np.newaxis]
    north = np.stack([-z * x, -z * y, x**2 + y**2], axis=1)
    north = north / np.sqrt((north[:, 0]**2 + north[:, 1]**2) + north[:, 2]**2)[:, np.newaxis]
    return (v_tan0[:, np.newaxis] * north + v_tan1[:, np.newaxis] * east +
            v_rad[:, np.newaxis] * rad)


def bounce(rc, Xhit, r_hit, u_alt, u_az, u_prob):
    """Surface impact of the rows ``Xhit`` (K,8) whose post-step radius is
    ``r_hit`` < 1 -- reference bouncepackets.py:39-100.  Uniform deviates are
    passed in (the reference draws them as three batches from ``randgen``)."""
    X = Xhit.copy()
    pos, vel = X[:, 1:4], X[:, 4:7]
    a = (vel[:, 0]**2 + vel[:, 1]**2) + vel[:, 2]**2
    b = 2 * ((pos[:, 0] * vel[:, 0] + pos[:, 1] * vel[:, 1]) + pos[:, 2] * vel[:, 2])
    c = ((pos[:, 0]**2 + pos[:, 1]**2) + pos[:, 2]**2) - 1.
    disc = np.sqrt(b**2 - 4 * a * c)
    t = np.minimum((-b - disc) / (2 * a), (-b + disc) / (2 * a))
    pos = pos + vel * t[:, np.newaxis]
    X[:, 1:4] = pos

    pe = 2 * rc.GM * (1. / r_hit - 1)
    v_old2 = a + pe
    v_old2[v_old2 < 0] = 0.

    direction = local_frame_direction(pos, u_alt, 2 * np.pi * u_az)

    def hit_lonlat():
        lon = (np.arctan2(pos[:, 0], -pos[:, 1]) + 2 * np.pi) % (2 * np.pi)
        return lon, np.arcsin(pos[:, 2])

    if rc.accomfactor == 0:
        v_new = np.sqrt(v_old2)
    else:
        lon, lat = hit_lonlat()
        tsurf = surface_temperature(rc.taa, lon, lat)
        v_emit = rc.v_interp(tsurf, u_prob)
        v_emit = v_emit / rc.planet_radius_km
        af = rc.accomfactor
        v_new = np.sqrt(v_emit**2 * af + v_old2 * (1 - af))
    X[:, 4:7] = direction * v_new[:, np.newaxis]

    if rc.sticktype == 'temperature dependent':
        lon, lat = hit_lonlat()
        coef = stick_coefficient(rc, lon, lat)
        assert np.all(coef <= 1) and np.all(coef >= 0)
        X[:, 7] *= (1 - coef)
    elif rc.stickcoef > 0:
        X[:, 7] *= (1 - rc.stickcoef)
    return X


def integrate_constant(X0, rc, uniforms=None, on_step=None, keep_trajectory=True):
    """Constant-step driver -- reference Output.py:368-455.

    ``uniforms(step_index, packet_index_array) -> (u_alt, u_az, u_prob)`` supplies
    the bounce deviates (reference: three ``randgen.random(K)`` batches per step).
    Returns (results (N,8,nsteps) or None, nsteps, attempted-steps (N,) int64)."""
    n = X0.shape[0]
    h0 = rc.step_size
    nsteps = int(np.ceil(rc.endtime / h0 + 1))
    cur = np.array(X0, dtype=np.float64, copy=True)
    results = None
    if keep_trajectory:
        results = np.zeros((n, 8, nsteps))
        results[:, :, 0] = cur
    nattempt = np.zeros(n, dtype=np.int64)
    simple_stick = (rc.sticktype == 'constant') and (rc.stickcoef == 1.)

    curtime = rc.endtime
    ct = 1
    live = cur[:, 7] > 0
    while (curtime > 0) and live.any():
        idx = np.nonzero(live)[0]
        todo = cur[idx]
        assert np.all(todo[:, 7] > 0) and np.all(np.isfinite(todo))
        nxt, _ = dp_step(todo, np.zeros(len(idx)) + h0, rc, want_error=False)
        nattempt[idx] += 1
        r = np.sqrt((nxt[:, 1]**2 + nxt[:, 2]**2) + nxt[:, 3]**2)
        hit = (r - 1.) < 0
        if simple_stick:
            nxt[hit, 7] = 0.
        elif hit.any():
            u_alt, u_az, u_prob = uniforms(ct, idx[hit])
            nxt[hit] = bounce(rc, nxt[hit], r[hit], u_alt, u_az, u_prob)
        nxt[r > rc.outeredge, 7] = 0
        nxt[nxt[:, 7] < 1e-10, 7] = 0.
        nxt[nxt[:, 7] == 0, 0] = 0.
        # packets that stopped keep all-zero rows from here on (dense tensor
        # is zero-initialised in the reference, Output.py:376)
        new = np.zeros_like(cur)
        new[idx] = nxt
        cur = new
        if keep_trajectory:
            results[:, :, ct] = cur
        if on_step is not None:
            on_step(ct, cur)
        live = cur[:, 7] > 0
        ct += 1
        curtime -= h0
    return results, nsteps, nattempt
