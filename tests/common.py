"""Shared helpers for the tests: workload loading, oracle constants from a
RunSetup, comparison metrics."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = os.path.join(REPO, 'nexoclom_b200', 'workloads')
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def workload(name):
    from nexoclom_b200.Input import Input
    return Input(os.path.join(WORKLOADS, name))


def oracle_constants(setup):
    """oracle.tracking.RunConstants carrying the same numbers as setup.params."""
    from oracle.tracking import RunConstants
    p = setup.params
    sint = setup.inputs.surfaceinteraction
    return RunConstants(
        GM=p.GM, vrplanet=p.vrplanet, gravity=bool(p.gravity), radpres=bool(p.radpres),
        radpres_v=setup.radpres_v, radpres_a=setup.radpres_a,
        lifetime=(1.0 / p.loss_rate if p.loss_mode == 1 else 0.0),
        photo=(p.loss_rate if p.loss_mode == 2 else None),
        resolution=p.resolution if p.resolution > 0 else 1e-4, outeredge=p.outeredge,
        step_size=p.step_size, endtime=p.endtime, sticktype=sint.sticktype,
        stickcoef=getattr(sint, 'stickcoef', 1.0),
        accomfactor=sint.accomfactor, A=tuple(p.stick_A),
        taa=float(np.asarray(setup.inputs.geometry.taa)),
        planet_radius_km=p.planet_radius_km,
        v_interp=(setup.surfaceint.v_interp if setup.surfaceint is not None and
                  setup.surfaceint.tck is not None else None),
        moons=list(getattr(setup, 'moons', [])))


def vec_rel(a, b):
    """|a-b| / |b| per row for (N,3) arrays."""
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


def state_parity(X_gpu, X_ref):
    """Relative differences between two (N,8) packet tables.  `pos` / `vel` (vector norms)
    and `time` cover EVERY packet whose alive flag agrees -- the dead ones too: their last
    position decides a pixel of `packet_image` when compress=False -- `frac` the packets
    alive in both (it is exactly 0 for the dead).  Returns maxima and mismatch counts."""
    alive_g, alive_r = X_gpu[:, 7] > 0, X_ref[:, 7] > 0
    both = alive_g & alive_r
    same = alive_g == alive_r
    out = dict(alive_mismatch=int((~same).sum()), n_both=int(both.sum()),
               n_dead_both=int((same & ~alive_g).sum()))
    if same.any():
        out['pos'] = float(vec_rel(X_gpu[same, 1:4], X_ref[same, 1:4]).max())
        out['vel'] = float(vec_rel(X_gpu[same, 4:7], X_ref[same, 4:7]).max())
        out['time'] = float(np.abs(X_gpu[same, 0] - X_ref[same, 0]).max())
        out['frac'] = 0.0
    if both.any():
        out['frac'] = float((np.abs(X_gpu[both, 7] - X_ref[both, 7]) / X_ref[both, 7]).max())
    return out


def source_case_input(tag):
    """Parsed tests/golden/source_cases/<tag>.input; table files named relative to the repo
    root are made absolute so the tests run from any directory."""
    from nexoclom_b200.Input import Input
    inputs = Input(os.path.join(GOLDEN, 'source_cases', tag + '.input'))
    for group, attr in ((inputs.spatialdist, 'mapfile'), (inputs.speeddist, 'vdistfile')):
        path = getattr(group, attr, None)
        if isinstance(path, str) and path != 'default' and not os.path.isabs(path):
            setattr(group, attr, os.path.join(REPO, path))
    return inputs
