"""files written by the reference are readable.

* the reader's protocol facts (class paths, astropy Quantity / unit pickling) are pinned
  against the reference's OWN pickle fixture, written by the reference with real astropy
  (``/root/reference/tests/test_data/input_classes_data.pkl`` -- build container only);
* ``tests/golden/reference_output.pkl`` is an Output file in the reference's on-disk format
  (tools/make_golden_refpickle.py: genuine reference inputs / units, oracle-made packets):
  ``Output.restore`` reads it, and on the GPU ``ModelImage`` bins it like any other run."""
import os

import numpy as np
import pytest

from common import GOLDEN, REPO

REF = os.environ.get('NEXOCLOM_REFERENCE', '/root/reference')
FIXTURE = os.path.join(REF, 'tests', 'test_data', 'input_classes_data.pkl')


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason='needs /root/reference (build container)')
def test_reads_the_references_own_pickles():
    from nexoclom_b200 import Input, refpickle
    from nexoclom_b200.input_classes import Geometry, Options
    from nexoclom_b200.solarsystem import SSObject
    names, inputs = refpickle.load(FIXTURE)
    assert len(names) == len(inputs) == 21
    for inp in inputs:
        assert isinstance(inp, Input) and isinstance(inp.geometry, Geometry)
        assert isinstance(inp.options, Options) and isinstance(inp.geometry.planet, SSObject)
        assert inp.options.endtime.unit == 's' and float(inp.options.endtime) > 0
    # the one inputfile of that (older) fixture that still exists in the reference tree
    name = 'test_data/inputfiles/Ca.surfacemap.maxwellian.input'
    ref = inputs[names.index(name)]
    mine = Input(os.path.join(REF, 'tests', name))
    for group in ('geometry', 'surfaceinteraction', 'forces', 'speeddist', 'angulardist',
                  'options'):
        assert getattr(ref, group) == getattr(mine, group), group
    assert ref.spatialdist.type == mine.spatialdist.type == 'surface map'
    # quantities and units: Jupiter system
    jup = inputs[names.index('test_data/inputfiles/Jupiter.01.input')].geometry
    ours = SSObject('Jupiter')
    assert jup.planet == ours and jup.startpoint == 'Io'
    for attr in ('radius', 'mass', 'a', 'tilt', 'rotperiod', 'orbperiod'):
        a, b = getattr(jup.planet, attr), getattr(ours, attr)
        assert a.unit == b.unit and float(a) == pytest.approx(float(b), rel=1e-12), attr
    assert jup.planet.GM.unit == 'm3/s2'
    assert float(jup.planet.GM) == pytest.approx(float(ours.GM), rel=1e-4)   # G of that astropy
    assert [float(p) for p in jup.phi] == [1.0, 2.0] and jup.phi[0].unit == 'rad'
    assert sorted(m.object for m in jup.planet.moons) == ['Callisto', 'Europa', 'Ganymede', 'Io']
    # the g-value fixture of the reference's tests goes through the same reader
    gv, rp = refpickle.load(os.path.join(REF, 'tests/unit_tests/atomicdata/g_value_test_data.pkl'))
    assert gv[0]['velocity'].unit == 'km/s' and len(np.asarray(gv[0]['velocity'])) > 100


def test_reference_format_output_is_restored():
    from nexoclom_b200 import Input, Output
    from nexoclom_b200 import catalogue
    g = np.load(os.path.join(GOLDEN, 'reference_output_packets.npz'))
    path = os.path.join(GOLDEN, 'reference_output.pkl')
    out = Output.restore(path)
    assert isinstance(out, Output) and isinstance(out.inputs, Input)
    assert out.npackets == int(g['npackets']) and out.totalsource == float(g['totalsource'])
    assert out.compress and out.idnum == 17 and out.planet.object == 'Mercury'
    assert out.aplanet.unit == 'au' and float(out.aplanet) == pytest.approx(float(g['aplanet']))
    assert out.vrplanet.unit == 'km/s'
    assert float(out.vrplanet) == pytest.approx(float(g['vrplanet_kms']), rel=1e-12)
    assert out.unit == 'R_Mercury'
    # restore() up-casts to 64 bit (Output.py:555-570); the values are the saved float32 ones
    X = out.X
    assert str(X['x'].dtype) == 'float64' and str(X['Index'].dtype) == 'int64'
    cols = ['time', 'x', 'y', 'z', 'vx', 'vy', 'vz', 'frac']
    assert np.array_equal(X[cols].values, g['X'].astype(np.float64))
    assert np.array_equal(X['Index'].values, g['index']) and (X.frac > 0).all()
    assert out.X0.shape == (out.npackets, 14)
    assert np.array_equal(out.X0.values, g['X0'].astype(np.float64))
    # the inputs inside the file are the reference's parse of Ca.isotropic.flat.input
    ours = Input(os.path.join(REPO, 'nexoclom_b200', 'workloads', 'Ca.isotropic.flat.input'))
    assert out.inputs.options.species == ours.options.species == 'Ca'
    assert float(out.inputs.options.endtime) == float(ours.options.endtime)
    catalogue._outputs.pop(path, None)


@pytest.mark.gpu
def test_model_image_over_a_reference_written_file(engine):
    """catalogue.adopt() makes the file known; ModelImage / LOSResult then treat it like a run
    of this package (the packets are uploaded as a resident table); deleting only forgets it."""
    from nexoclom_b200 import ModelImage, catalogue
    from nexoclom_b200.runsetup import RunSetup
    from oracle import imaging
    path = os.path.join(GOLDEN, 'reference_output.pkl')
    out = catalogue.adopt(path)
    inputs = out.inputs
    ids, files, npk, tot = inputs.search()
    assert files == [path] and npk == out.npackets
    im = ModelImage(inputs, {'quantity': 'radiance', 'dims': '250,250'})
    g = np.load(os.path.join(GOLDEN, 'reference_output_packets.npz'))
    P = g['X'].astype(np.float64)
    setup = RunSetup(inputs)
    oi, oc, _, _ = imaging.create_image(
        P[:, 1], P[:, 2], P[:, 3], P[:, 5], P[:, 7], vrplanet=setup.vrplanet,
        M=imaging.image_rotation(0, np.pi / 2), dims=[250, 250], xrange=(-4, 4), zrange=(-4, 4),
        apix=float(im.Apix), quantity='radiance', gtables=setup.gtables([4227]))
    assert np.array_equal(im.packet_image, oc) and oc.sum() > 500
    oi *= im.atoms_per_packet
    nz = oi > 0
    assert nz.sum() > 50 and np.max(np.abs(im.image[nz] - oi[nz]) / oi[nz]) < 1e-6
    inputs.delete_files()
    assert os.path.exists(path) and inputs.search()[1] == []
