"""The local run catalogue (stands in for the reference's `outputfile` table and
Input.search / delete_files, reference Input.py:121-172, Output.py:457-548): runs saved by
one process are found, summed, fetched and deleted by a later one."""
import os
import types

import pytest

from common import workload
from nexoclom_b200 import catalogue


def _forget():
    catalogue._outputs.clear()
    catalogue._by_key.clear()
    catalogue._counter[0] = 0


def _fake_output(npackets, totalsource):
    return types.SimpleNamespace(npackets=npackets, totalsource=totalsource, payload=list(range(5)))


def test_runs_persist_across_processes(tmp_path, monkeypatch):
    monkeypatch.setenv('NEXOCLOM_B200_SAVEPATH', str(tmp_path))
    _forget()
    a = workload('Na.maxwellian.radpres.input')
    b = workload('Ca.isotropic.flat.input')
    id1, f1 = catalogue.register(a, _fake_output(1000, 1000.0))
    id2, f2 = catalogue.register(b, _fake_output(50, 50.0))
    id3, f3 = catalogue.register(a, _fake_output(2000, 1500.5))
    assert (id1, id2, id3) == (1, 2, 3) and all(os.path.exists(f) for f in (f1, f2, f3))
    assert catalogue.search(a) == ([1, 3], [f1, f3], 3000, 2500.5)

    _forget()                                    # "a new process"
    assert catalogue.search(a) == ([1, 3], [f1, f3], 3000, 2500.5)
    assert catalogue.search(b) == ([2], [f2], 50, 50.0)
    assert catalogue.fetch(f3).payload == list(range(5))
    id4, f4 = catalogue.register(b, _fake_output(7, 7.0))
    assert id4 == 4                              # idnums keep counting over the directory
    assert catalogue.search(b)[0] == [2, 4] and catalogue.search(b)[2] == 57

    _forget()
    catalogue.delete(a, f1)
    assert catalogue.search(a) == ([3], [f3], 2000, 1500.5) and not os.path.exists(f1)
    a.delete_files()
    assert catalogue.search(a) == ([], [], 0, 0)
    assert sorted(os.listdir(tmp_path)) == sorted(
        os.path.basename(p) for f in (f2, f4) for p in (f, f[:-4] + '.json'))
    _forget()


def test_everything_an_output_carries_can_be_pickled():
    """Saved Outputs are pickles of the object (reference Output.py:545-548): the pieces it
    holds besides the packet tables must survive a round trip for every bundled workload."""
    import pickle
    from common import WORKLOADS
    from nexoclom_b200.Output import RadPres
    from nexoclom_b200.runsetup import RunSetup
    from nexoclom_b200.units import Quantity, def_unit
    names = sorted(f for f in os.listdir(WORKLOADS) if f.endswith('.input'))
    assert len(names) >= 5
    for name in names:
        inputs = workload(name)
        setup = RunSetup(inputs)
        rp = RadPres()
        rp.velocity, rp.accel = setup.radpres_v, setup.radpres_a
        for what, obj in (('inputs', inputs), ('loss_info', setup.loss_info),
                          ('surfaceint', setup.surfaceint), ('planet', inputs.geometry.planet),
                          ('unit', def_unit('R_x', 'length', 2.44e6)),
                          ('quantity', Quantity(1.5, 'km/s')), ('radpres', rp)):
            try:
                back = pickle.loads(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))
            except Exception as exc:                                 # noqa: BLE001
                pytest.fail(f'{name}: {what} does not pickle: {exc!r}')
            assert type(back) is type(obj)
        if setup.surfaceint is not None and callable(getattr(setup.surfaceint, 'stickcoef', None)):
            import numpy as np
            lon, lat = np.array([0.1, 3.0]), np.array([0.2, -0.4])
            back = pickle.loads(pickle.dumps(setup.surfaceint))
            assert np.array_equal(back.stickcoef(lon, lat), setup.surfaceint.stickcoef(lon, lat))


def test_memory_only_without_savepath(monkeypatch):
    monkeypatch.delenv('NEXOCLOM_B200_SAVEPATH', raising=False)
    _forget()
    a = workload('Na.maxwellian.radpres.input')
    idn, fname = catalogue.register(a, _fake_output(10, 10.0))
    assert fname.startswith('mem://') and catalogue.search(a) == ([idn], [fname], 10, 10.0)
    a.delete_files()
    assert catalogue.search(a) == ([], [], 0, 0)
    _forget()


@pytest.mark.gpu
def test_output_saved_by_one_process_is_used_by_another(tmp_path, monkeypatch):
    """Input.run in this process; search / Output.restore / ModelImage in a fresh one."""
    import subprocess
    import sys
    import numpy as np
    from nexoclom_b200 import Output
    monkeypatch.setenv('NEXOCLOM_B200_SAVEPATH', str(tmp_path))
    _forget()
    inputs = workload('Na.maxwellian.radpres.input')
    out = Output(inputs, 20000, seed=3)
    assert os.path.exists(out.filename)
    frac_sum = float(out.X.frac.sum())
    code = f"""
import sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
from common import workload
from nexoclom_b200 import Output, ModelImage
inputs = workload('Na.maxwellian.radpres.input')
ids, files, npk, tot = inputs.search()
assert files == [{out.filename!r}] and npk == 20000, (files, npk)
o = Output.restore(files[0])
print('FRAC', repr(float(o.X.frac.sum())), len(o.X), o.X.x.dtype)
im = ModelImage(inputs, {{'quantity': 'radiance', 'dims': '200,200'}})
print('IMAGE', float(im.image.sum()) > 0, int(im.packet_image.sum()))
"""
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    frac_line = [l for l in res.stdout.splitlines() if l.startswith('FRAC')][0].split()
    assert float(frac_line[1]) == pytest.approx(frac_sum, rel=1e-6) and frac_line[3] == 'float64'
    assert 'IMAGE True' in res.stdout
    inputs.delete_files()
    assert os.listdir(tmp_path) == []
    _forget()
