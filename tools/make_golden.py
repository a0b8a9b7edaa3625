#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions
(rk5, state, bouncepackets, surface_temperature and the two Output drivers)
imported from /root/reference through tools/refimport.py.

Build-container only (needs /root/reference).  The produced fixtures travel to
the GPU box; tests/test_oracle_golden.py replays them through oracle/ and
tests/test_gpu_parity.py through the CUDA path.
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

import refimport                                       # noqa: E402
from common import workload, oracle_constants         # noqa: E402
from nexoclom_b200.runsetup import RunSetup            # noqa: E402
from oracle import initial_state                       # noqa: E402

COLS = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']
GOLD = os.path.join(REPO, 'tests', 'golden')


def fake_from_setup(setup, seed=0, ref_surfaceint=None):
    p = setup.params
    sint = setup.inputs.surfaceinteraction
    fo = refimport.fake_output(
        GM=p.GM, vrplanet=p.vrplanet, radpres_v=setup.radpres_v, radpres_a=setup.radpres_a,
        gravity=bool(p.gravity), radpres=bool(p.radpres),
        lifetime=(1.0 / p.loss_rate if p.loss_mode == 1 else 0.0),
        photo=(p.loss_rate if p.loss_mode == 2 else None), step_size=p.step_size,
        resolution=(p.resolution if p.resolution > 0 else None), outeredge=p.outeredge,
        endtime=p.endtime, stickcoef=getattr(sint, 'stickcoef', None), sticktype=sint.sticktype,
        accomfactor=sint.accomfactor, A=tuple(p.stick_A),
        taa=float(np.asarray(setup.inputs.geometry.taa)), planet_radius_km=p.planet_radius_km,
        seed=seed)
    if setup.surfaceint is not None:
        # the accommodation table / sticking closure of the REFERENCE's own
        # SurfaceInteraction.__init__ when the caller built one (reference_surfaceint),
        # else this repo's port (bit-identical: tests/test_host_tables.py)
        fo.surfaceint = ref_surfaceint if ref_surfaceint is not None else setup.surfaceint
    return fo


def reference_surfaceints():
    """{workload: reference SurfaceInteraction} built by the unmodified reference class
    (tools/reftables.py) BEFORE the lighter refimport stubs are installed."""
    import reftables
    out = {}
    for wl in ('Na.bounce.input', 'Na.bounce.stick05.input'):
        out[wl] = reftables.surface_interaction(workload(wl))
    reftables.purge()
    return out


def main():
    surfaceints = reference_surfaceints()
    ref = refimport.install()
    os.makedirs(GOLD, exist_ok=True)

    # ---- 1. single Dormand-Prince steps (rk5 + state), three force configs ----
    out = {}
    rng = np.random.default_rng(2024)
    for tag, wl in (('na', 'Na.maxwellian.radpres.input'), ('ca', 'Ca.isotropic.flat.input'),
                    ('grav', 'Gravity.input')):
        setup = RunSetup(workload(wl))
        fo = fake_from_setup(setup)
        fo.inputs.options.step_size = 0.0           # ask rk5 for the error vector
        n = 512
        x0 = initial_state.draw_x0(setup, n, 11)[:, :8]
        x0[:, 1:4] *= (1 + 3 * rng.random(n))[:, None]      # spread radii 1..4 R_p
        x0[:, 7] = 0.05 + 0.95 * rng.random(n)
        x0[:, 0] = 1 + 5e4 * rng.random(n)
        h = np.minimum(x0[:, 0], 10**rng.uniform(0, 3, n))
        res, delta = ref.rk5(fo, x0.copy(), h)
        acc, rate = ref.state(x0.copy(), fo)
        out[f'{tag}_x0'], out[f'{tag}_h'] = x0, h
        out[f'{tag}_result'], out[f'{tag}_delta'] = res, delta
        out[f'{tag}_accel'], out[f'{tag}_rate'] = acc, rate
    np.savez_compressed(os.path.join(GOLD, 'rk5_steps.npz'), **out)
    print('rk5_steps.npz')

    # ---- 2. adaptive driver (Output.variable_step_size_driver) ----
    out = {}
    for tag, wl, n in (('na', 'Na.maxwellian.radpres.input', 400),
                       ('ca', 'Ca.isotropic.flat.input', 300)):
        setup = RunSetup(workload(wl))
        fo = fake_from_setup(setup)
        x0 = initial_state.draw_x0(setup, n, 5)[:, :8]
        fo.X = pd.DataFrame(x0.copy(), columns=COLS)
        fo.X['lossfrac'] = 0.
        fo.npackets = n
        ref.Output.variable_step_size_driver(fo)
        out[f'{tag}_x0'] = x0
        out[f'{tag}_final'] = fo.X[COLS].values
        out[f'{tag}_step'] = fo.X['step_size'].values
        print(tag, 'adaptive alive', (fo.X.frac > 0).mean())
    np.savez_compressed(os.path.join(GOLD, 'adaptive_driver.npz'), **out)
    print('adaptive_driver.npz')

    # ---- 3. constant-step driver + bouncepackets ----
    out = {}
    for tag, wl, n in (('tdep', 'Na.bounce.input', 200), ('c05', 'Na.bounce.stick05.input', 200),
                       ('grav', 'Gravity.input', 100)):
        inputs = workload(wl)
        if tag == 'tdep':
            inputs.options.endtime = type(inputs.options.endtime)(3000., 's')
        setup = RunSetup(inputs)
        seed = 31
        fo = fake_from_setup(setup, seed=seed, ref_surfaceint=surfaceints.get(wl))
        x0 = initial_state.draw_x0(setup, n, 9)[:, :8]
        fo.X0 = pd.DataFrame(x0.copy(), columns=COLS)
        fo.npackets = n
        fo.totalsource = float(n)
        ref.Output.constant_step_size_driver(fo)
        nsteps = fo.nsteps
        traj = fo.X[COLS].values.reshape(n, nsteps, 8).transpose(0, 2, 1)
        out[f'{tag}_x0'] = x0
        out[f'{tag}_traj'] = traj
        out[f'{tag}_seed'] = np.int64(seed)
        out[f'{tag}_endtime'] = np.float64(setup.params.endtime)
        print(tag, 'const alive at end', (traj[:, 7, -1] > 0).mean(), 'nsteps', nsteps)
    np.savez_compressed(os.path.join(GOLD, 'constant_driver.npz'), **out)
    print('constant_driver.npz')

    # ---- 4. surface temperature + rebound direction ----
    rng = np.random.default_rng(3)
    lon = rng.random(256) * 2 * np.pi
    lat = np.arcsin(rng.random(256) * 2 - 1)
    geo = types.SimpleNamespace(startpoint='Mercury', taa=refimport.TaaQuantity(1.3))
    ts = ref.surface_temperature(geo, lon, lat)
    pos = np.stack([np.sin(lon) * np.cos(lat), -np.cos(lon) * np.cos(lat), np.sin(lat)], 1)
    fo = types.SimpleNamespace(randgen=np.random.default_rng(77))
    direction = ref.rebound_direction(fo, pos.copy())
    np.savez_compressed(os.path.join(GOLD, 'surface.npz'), lon=lon, lat=lat, taa=1.3,
                        tsurf=ts, pos=pos, direction=direction, seed=77)
    print('surface.npz')


if __name__ == '__main__':
    main()
