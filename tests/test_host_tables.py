"""Host-side tables against the reference's own golden vectors / known answers
(SURVEY section 8c): tests/unit_tests/atomicdata/g_value_test_data.pkl (converted
to tests/golden/gvalue_golden.npz), test_photolossrates.py:5-7,
test_xyz_from_latlon.py:8-65, test_SSObject.py:6-34."""
import os

import numpy as np
import pytest
from pytest import approx

from common import GOLDEN
from nexoclom_b200.atomicdata import gValue, RadPresConst, PhotoRate, atomicmass
from nexoclom_b200.solarsystem import SSObject, planet_dist
from oracle.initial_state import xyz_from_lonlat


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'gvalue_golden.npz'))


@pytest.mark.parametrize('i, species, wavelength, aplanet',
                         [(0, 'Na', 5891, 1.5), (1, 'Ca', 4227, 0.3), (2, 'X', 3333, 1.0)])
def test_gvalue(gold, i, species, wavelength, aplanet):
    g = gValue(species, wavelength, aplanet)
    assert g.species == str(gold[f'g{i}_species'])
    assert g.wavelength.value == float(gold[f'g{i}_wavelength'])
    assert g.aplanet.value == float(gold[f'g{i}_aplanet'])
    assert g.velocity.value == approx(gold[f'g{i}_velocity'], rel=1e-14)
    assert g.g.value == approx(gold[f'g{i}_g'], rel=1e-14)


@pytest.mark.parametrize('i, species, aplanet', [(0, 'Na', 1.5), (1, 'Ca', 0.3), (2, 'X', 1.0)])
def test_radpresconst(gold, i, species, aplanet):
    rp = RadPresConst(species, aplanet)
    assert rp.species == str(gold[f'r{i}_species'])
    assert rp.aplanet.value == float(gold[f'r{i}_aplanet'])
    assert rp.velocity.value == approx(gold[f'r{i}_velocity'], rel=1e-14)
    assert rp.accel.value == approx(gold[f'r{i}_accel'], rel=1e-13, abs=0)


@pytest.mark.parametrize('species, aplanet, result',
                         [('Na', 1.5, 3.2266666666666665e-06),
                          ('Ca', 0.3, 0.0007777777777777777), ('X', 1.0, 1e-30)])
def test_photorate(species, aplanet, result):
    rate = PhotoRate(species, aplanet)
    assert rate.species == species
    assert rate.aplanet.value == aplanet
    assert rate.rate.value == approx(result, rel=1e-15)


def test_atomicmass():
    assert atomicmass('Na').value == 22.98977          # reference atomicmass.py:32
    assert atomicmass('X') is None


def test_ssobject():
    m = SSObject('Mercury')
    assert m.object == 'Mercury' and m.orbits == 'Sun' and m.type == 'Planet'
    assert m.moons is None and len(m) == 1
    assert m.radius.value == 2440.53
    assert m.GM.value < 0
    j = SSObject('jupiter')
    assert j.object == 'Jupiter'
    assert [x.object for x in j.moons] == ['Io', 'Europa', 'Ganymede', 'Callisto']
    assert len(j) == 5
    assert SSObject('Io').type == 'Moon'
    assert SSObject('Nowhere').object is None


def test_planet_dist():
    m = SSObject('Mercury')
    r0, v0 = planet_dist(m, 0.)
    r1, v1 = planet_dist(m, np.pi)
    a, e = m.a.value, m.e
    assert r0.value == approx(a * (1 - e)) and r1.value == approx(a * (1 + e))
    assert abs(v0.value) < 0.1 and abs(v1.value) < 0.1
    r, v = planet_dist(m, 1.3)
    # outbound leg: positive radial velocity of order 10 km/s
    assert 9 < v.value < 10.5 and 0.34 < r.value < 0.36


def test_xyz_from_lonlat_known_answers():
    s2 = np.sqrt(2) / 2
    lon = np.arange(0, 2 * np.pi, np.pi / 4)
    lat = np.zeros_like(lon)
    xyz = xyz_from_lonlat(lon, lat, True, 1.)
    x = [0, s2, 1., s2, 0., -s2, -1, -s2]
    y = [-1, -s2, 0, s2, 1, s2, 0, -s2]
    assert xyz == approx(np.array([x, y, np.zeros(8)]), abs=1e-15)
    xyz = xyz_from_lonlat(lon, lat, False, 1.)
    assert xyz == approx(np.array([[-v for v in x], y, np.zeros(8)]), abs=1e-15)
    lon = np.linspace(0, np.pi, 5)
    lat = np.linspace(-np.pi / 2, np.pi / 2, 5)
    xyz = xyz_from_lonlat(lon, lat, True, 2.)
    z = np.array([-1, -s2, 0, s2, 1])
    z_ = np.array([0, s2, 1, s2, 0])
    assert xyz == approx(np.array([np.array([0, s2, 1., s2, 0]) * z_,
                                   np.array([-1, -s2, 0., s2, 1]) * z_, z]) * 2., abs=1e-15)


# ---------------------------------------------------------------------------
# host tables pinned to outputs of the UNMODIFIED reference builders
# (tools/make_golden_tables.py -> tests/golden/host_tables.npz)
# ---------------------------------------------------------------------------
@pytest.fixture(scope='module')
def tables_gold():
    return np.load(os.path.join(GOLDEN, 'host_tables.npz'))


@pytest.mark.parametrize('tag, wl, taa', [('tdep314', 'Na.bounce.input', None),
                                           ('tdep130', 'Na.bounce.input', 1.3),
                                           ('c05', 'Na.bounce.stick05.input', None)])
def test_surface_interaction_table_vs_reference(tables_gold, tag, wl, taa):
    """probgrid / accommodation spline / sticking closure of the port are bit-identical to
    what the reference's SurfaceInteraction.__init__ (SurfaceInteraction.py:10-61) built."""
    from common import workload
    from nexoclom_b200.surfaceinteraction import SurfaceInteraction
    from nexoclom_b200.units import Quantity
    g = tables_gold
    inputs = workload(wl)
    if taa is not None:
        inputs.geometry.taa = Quantity(float(taa), 'rad')
    si = SurfaceInteraction(inputs)
    assert np.array_equal(si.probgrid, g[f'{tag}_probgrid'])
    assert np.array_equal(si.temperature, g[f'{tag}_temperature'])
    assert np.array_equal(si.probability, g[f'{tag}_probability'])
    v = si.v_interp(g[f'{tag}_sample_T'], g[f'{tag}_sample_P'])
    assert np.array_equal(v, g[f'{tag}_v_interp'])
    if f'{tag}_stickcoef' in g.files:
        assert np.array_equal(si.stickcoef(g[f'{tag}_lon'], g[f'{tag}_lat']),
                              g[f'{tag}_stickcoef'])
    else:
        assert not hasattr(si, 'stickcoef')


def test_planet_dist_vs_reference(tables_gold):
    """(r, v_r) of the port == the reference's planet_dist (planet_dist.py:29-74) at the
    workloads' true anomalies 0 / 1.3 / 3.14 and on a sweep; vrplanet shifts every
    radiation-pressure and g-value lookup, so this is bit-exact, not approximate."""
    g = tables_gold
    for planet, key in (('Mercury', 'mercury'), ('Jupiter', 'jupiter'), ('Mars', 'mars')):
        got = np.array([[float(q.value) for q in planet_dist(planet, float(t))]
                        for t in g[f'{key}_taa']])
        assert np.array_equal(got, g[f'{key}_r_vr']), planet
    for name in ('Mercury', 'Jupiter', 'Io'):
        o = SSObject(name)
        got = np.array([o.radius.value, o.mass.value, o.a.value, o.e, o.orbperiod.value,
                        o.GM.value])
        assert got == approx(g[f'ss_{name.lower()}'], rel=1e-15), name
