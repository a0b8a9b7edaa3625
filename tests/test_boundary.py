"""The drop-in boundary: C-ABI exports, struct layouts, inputfile grammar and
defaults, no-oracle / no-CPU-fallback rules."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from common import REPO, WORKLOADS, workload
from nexoclom_b200 import _lib
from nexoclom_b200.Input import Input
from nexoclom_b200.input_classes import (Options, SurfaceInteraction, SpatialDist, AngularDist,
                                         Forces, Geometry, InputError)


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(REPO, 'include', 'nexoclom_b200.h')).read()
    declared = set(re.findall(r'^\s*(?:int|const char\*)\s+(nx_\w+)\s*\(', header, re.M))
    assert len(declared) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    bound = _lib.load()
    for s in declared:
        assert hasattr(bound, s)


def test_struct_layouts_match_header():
    # sizes the static_asserts in nx_api.cu also pin against the kernels' structs
    assert C.sizeof(_lib.RunParams) == 15 * 8 + 6 * 4 + 5 * 4 * 8          # + 4 moons x 5 doubles
    assert C.sizeof(_lib.SourceParams) == 4 * 4 + 14 * 8 + 4 * 4 + 8 + 2 * 4 + 4 * 8
    assert C.sizeof(_lib.SourceMapParams) == 2 * 8 + 6 * 4
    assert C.sizeof(_lib.ImageParams) == 15 * 8 + 6 * 4
    assert C.sizeof(_lib.LosParams) == 4 * 8 + 4 * 4


def test_sass_is_sm100a(built):
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out.stdout


def test_product_never_imports_oracle():
    bad = []
    for root, _, files in os.walk(os.path.join(REPO, 'nexoclom_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(root, f)).read()
                if (re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M) or
                        re.search(r'(CDLL|#include|import).*hostcheck', txt)):
                    bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback():
    """Without a CUDA device context creation fails loudly (rc != 0 -> raises)."""
    code = ('import sys; sys.path.insert(0, %r)\n'
            'from nexoclom_b200.engine import Engine, NexoclomCudaError\n'
            'try:\n    Engine(0)\nexcept NexoclomCudaError as e:\n    print("RAISED")\n' % REPO)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env)
    assert 'RAISED' in out.stdout, out.stdout + out.stderr


def _write(tmp_path, text):
    p = tmp_path / 'x.input'
    p.write_text(text)
    return str(p)


def test_inputfile_grammar_and_defaults(tmp_path):
    inp = Input(_write(tmp_path, '''
# comment line
Geometry.Planet = mercury   ; trailing comment
geometry.TAA = 1.3 # hash comment
SpatialDist.type = uniform
SpeedDist.type = flat
SpeedDist.vprob = 4.
SpeedDist.delv = 4. ; semicolon wins # over hash
options.endtime = 100.
options.species = na
bad line without equals
a.b.c = 3
'''))
    assert inp.geometry.planet.object == 'Mercury' and inp.geometry.startpoint == 'Mercury'
    assert float(inp.geometry.taa) == 1.3 and inp.geometry.phi is None
    assert inp.forces.gravity is True and inp.forces.radpres is True
    assert inp.surfaceinteraction.sticktype == 'constant'
    assert inp.surfaceinteraction.stickcoef == 1. and inp.surfaceinteraction.accomfactor is None
    assert inp.angulardist.type == 'isotropic'
    assert [float(a) for a in inp.angulardist.altitude] == [0., np.pi / 2]
    assert [float(a) for a in inp.angulardist.azimuth] == [0., 2 * np.pi]
    assert inp.options.species == 'Na' and inp.options.outeredge == 1e30
    assert inp.options.step_size == 0. and inp.options.resolution == 1e-4
    assert float(inp.options.lifetime) == 0. and inp.options.fitted is False
    assert float(inp.speeddist.delv) == 4.
    assert [float(a) for a in inp.spatialdist.latitude] == [-np.pi / 2, np.pi / 2]
    assert inp == Input(inp._inputfile)


def test_input_quirks_and_errors(tmp_path):
    # Q10: resolution from a file stays a string; 'stepsize' alias raises KeyError
    assert Options({'endtime': '1', 'species': 'Na', 'resolution': '1e-5'}).resolution == '1e-5'
    with pytest.raises(KeyError):
        Options({'endtime': '1', 'species': 'Na', 'stepsize': '30'})
    assert Options({'endtime': '1', 'atom': 'ca', 'step_size': '30'}).resolution is None
    with pytest.raises(InputError):
        Options({'species': 'Na'})
    with pytest.raises(InputError):
        SpatialDist({})
    with pytest.raises(InputError):
        SpatialDist({'type': 'uniform', 'latitude': '1.0, 0.5'})
    with pytest.raises(InputError):
        SurfaceInteraction({'stickcoef': '0.5'})
    s = SurfaceInteraction({'sticktype': 'temperature dependent', 'accomfactor': '0.2'})
    assert s.A == (1.57014, -0.006262, 0.1614157) and 'stickcoef' not in s.__dict__
    assert SurfaceInteraction({'stickcoef': '7'}).stickcoef == 1
    assert SurfaceInteraction({'stickcoef': '-1', 'accomfactor': '0'}).stickcoef == 0
    assert Forces({'gravity': 'false'}).gravity is False
    with pytest.raises(InputError):
        Geometry({})
    with pytest.raises(ValueError):
        Geometry({'planet': 'Mercury', 'startpoint': 'Io'})
    g = Geometry({'planet': 'Jupiter', 'startpoint': 'Io', 'objects': 'Jupiter, Io', 'phi': '1'})
    assert len(g.phi) == 1 and float(g.phi[0]) == 1.0
    a = AngularDist({'type': 'Isotropic', 'altitude': '0.2, 9'})
    assert [float(v) for v in a.altitude] == [0.2, np.pi / 2]
    with pytest.raises(FileNotFoundError):
        Input(str(tmp_path / 'missing.input'))


def test_workload_files_parse():
    for f in sorted(os.listdir(WORKLOADS)):
        inp = workload(f)
        assert inp.options.species in ('Na', 'Ca')


def test_pinned_pool_capacities_and_fallback(built):
    """Result buffers of the host side: capacities grow in eighths of an octave (buffers of
    slightly different sizes are shared), and without a usable device ``nx_host_alloc`` fails
    and the pool hands out ordinary NumPy arrays (the library then stages the copy)."""
    from nexoclom_b200.engine import _PinnedPool
    caps = [_PinnedPool._capacity(n) for n in (1, 4096, 4097, 10 ** 6, 10 ** 7, 10 ** 7 + 1, 10 ** 8)]
    assert caps == sorted(caps) and caps[0] == 4096
    for n, c in zip((10 ** 6, 10 ** 7, 10 ** 8), caps[3::1][0:1] + caps[4:5] + caps[6:7]):
        assert n <= c <= 1.07 * n
    pool = _PinnedPool(_lib.load())
    a = pool.array((3, 5))
    assert a.shape == (3, 5) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable
    b = pool.array((7,), dtype=np.uint8)
    assert b.shape == (7,) and b.dtype == np.uint8
    assert pool.array((0,)).size == 0
