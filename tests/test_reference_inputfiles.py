"""The reference's own boundary regression (tests/unit_tests/Initial_state/
test_input_classes.py:17-143), restated against THIS package's Input class and run on the
reference's own inputfiles where they lie (tests/test_data/inputfiles/*.input; build
container only -- skipped where /root/reference is absent).  Also test_SSObject.py:6-34."""
import os

import numpy as np
import pytest

from nexoclom_b200 import Input, SSObject
from nexoclom_b200.units import Quantity

INPUTS = os.path.join(os.environ.get('NEXOCLOM_REFERENCE', '/root/reference'), 'tests',
                      'test_data', 'inputfiles')
pytestmark = pytest.mark.skipif(not os.path.isdir(INPUTS), reason='reference tree not present')


def rad(v):
    return Quantity(v, 'rad')


def same(d, expected):
    assert set(d) == set(expected), (sorted(d), sorted(expected))
    for k, v in expected.items():
        got = d[k]
        if isinstance(v, tuple):
            assert len(got) == len(v) and all(float(a) == pytest.approx(float(b)) for a, b in zip(got, v)), k
        elif isinstance(v, (float, Quantity)) and not isinstance(v, bool):
            assert float(got) == pytest.approx(float(v)), k
        else:
            assert got == v, (k, got, v)


def test_geometry():
    g1 = Input(os.path.join(INPUTS, 'Geometry.01.input')).geometry
    same(g1.__dict__, {'planet': SSObject('Jupiter'), 'startpoint': 'Io',
                       'objects': {SSObject('Jupiter'), SSObject('Io'), SSObject('Europa')},
                       'type': 'geometry without starttime', 'phi': (rad(1), rad(2)),
                       'subsolarpoint': (rad(3.14), rad(0)), 'taa': rad(1.57)})
    g2 = Input(os.path.join(INPUTS, 'Geometry.02.input')).geometry
    assert g2.planet == SSObject('Jupiter') and g2.startpoint == 'Io'
    assert g2.objects == {SSObject('Jupiter'), SSObject('Io')}
    assert g2.type == 'geometry with starttime' and '2022-03-08T19:53:21' in str(g2.time)
    g3 = Input(os.path.join(INPUTS, 'Geometry.03.input')).geometry
    same(g3.__dict__, {'planet': SSObject('Mercury'), 'startpoint': 'Mercury',
                       'objects': {SSObject('Mercury')}, 'type': 'geometry without starttime',
                       'subsolarpoint': (rad(0), rad(0)), 'phi': None, 'taa': rad(3.14)})
    assert g1 == g1 and g1 != g2 and g1 != g3


def test_surface_interaction():
    def si(k):
        return Input(os.path.join(INPUTS, f'SurfaceInteraction.0{k}.input')).surfaceinteraction
    same(si(1).__dict__, {'sticktype': 'constant', 'stickcoef': 1., 'accomfactor': None})
    same(si(2).__dict__, {'sticktype': 'constant', 'stickcoef': 0.5, 'accomfactor': 0.2})
    assert si(1) == si(1) and si(1) != si(2)
    same(si(3).__dict__, {'sticktype': 'temperature dependent', 'accomfactor': 0.2,
                          'A': (1.57014, -0.006262, 0.1614157)})
    same(si(4).__dict__, {'sticktype': 'temperature dependent', 'accomfactor': 0.5,
                          'A': (1., 0.001, 0.2)})
    # The reference's test expects a 'coordinate_system' key here, but its code
    # (input_classes.py:277-295, v3.7.4) sets 'stick_map' instead -- the test is stale and
    # fails on the reference itself; the code is the specification.
    same(si(5).__dict__, {'sticktype': 'surface map', 'stick_mapfile': 'default',
                          'stick_map': None, 'subsolarlon': None, 'accomfactor': 0.5})
    same(si(6).__dict__, {'sticktype': 'surface map', 'stick_mapfile': 'Orbit3576.Ca.pkl',
                          'stick_map': None, 'subsolarlon': None, 'accomfactor': 0.5})


def test_forces():
    for k, (g, r) in enumerate([(True, True), (False, True), (True, False)], start=1):
        f = Input(os.path.join(INPUTS, f'Forces.0{k}.input')).forces
        assert f.__dict__ == {'gravity': g, 'radpres': r}


def test_spatial_dist():
    s1 = Input(os.path.join(INPUTS, 'Spatial.01.input')).spatialdist
    same(s1.__dict__, {'type': 'uniform', 'longitude': (rad(0), rad(2 * np.pi)),
                       'latitude': (rad(-np.pi / 2), rad(np.pi / 2)), 'exobase': 1.})
    s2 = Input(os.path.join(INPUTS, 'Spatial.02.input')).spatialdist
    same(s2.__dict__, {'type': 'uniform', 'longitude': (rad(0), rad(3.14)),
                       'latitude': (rad(0), rad(0.79)), 'exobase': 2.1})


def test_every_reference_inputfile_parses():
    for fn in sorted(os.listdir(INPUTS)):
        if fn.endswith('.input'):
            inp = Input(os.path.join(INPUTS, fn))
            assert inp.geometry.planet.object is not None, fn


def test_ssobject():
    """reference tests/unit_tests/solarsystem/test_SSObject.py:6-34."""
    m = SSObject('Mercury')
    assert m.object == 'Mercury' and m.moons is None and m.type == 'Planet' and len(m) == 1
    j = SSObject('Jupiter')
    assert j.moons is not None and len(j) == len(j.moons) + 1
    assert {x.object for x in j.moons} >= {'Io', 'Europa', 'Ganymede', 'Callisto'}
    io = SSObject('Io')
    assert io.type == 'Moon' and io.orbits == 'Jupiter' and io.GM.value < 0
