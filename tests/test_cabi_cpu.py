'''CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/tfb200.h declares; host-side preparation agrees with the oracle's independent copy.'''
import ctypes
import os
import re

import numpy

from transiflow_b200 import _lib, hostprep, recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'tfb200.h')).read()
    declared = set(re.findall(r'\b(tfb_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(_lib.SO)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_params_struct_layout_matches_header():
    assert ctypes.sizeof(hostprep.TfbParams) == 5 * 8 + 2 * 8 * hostprep.TFB_MAX_FORCE + 4 * 4


def test_no_gpu_fails_loudly():
    if _lib.device_count() > 0:
        return
    import pytest
    from transiflow_b200 import Interface
    with pytest.raises(RuntimeError):
        Interface({'Reynolds Number': 100}, 4, 4, 4)


def test_coordinate_vectors_match_oracle_copy():
    from oracle import tf_oracle
    for p in ({}, {'Grid Stretching Factor': 1.5}, {'Grid Stretching': True, 'Grid Stretching Method': 'sin'},
              {'Grid Stretching Factor': 2.0, 'X-max': 5}):
        for n in (1, 4, 13):
            a = hostprep.coordinate_vector(p, 0.0, p.get('X-max', 1.0), n)
            b = tf_oracle.coordinate_vector(p, 0.0, p.get('X-max', 1.0), n)
            assert numpy.array_equal(a, b)


def test_every_problem_type_has_a_kernel_family():
    for problem, dim, nz, dof in ((recipes.LDC, 2, 1, 3), (recipes.LDC, 3, 8, 4), (recipes.RB, 2, 1, 4),
                                  (recipes.RBP, 3, 8, 5), (recipes.DHC, 2, 1, 4), (recipes.DHC, 3, 4, 5),
                                  (recipes.QG, 2, 1, 3), (recipes.AMOC, 2, 1, 5), (recipes.LDC, 3, 1, 4),
                                  (recipes.RB, 3, 1, 5), (recipes.DHC, 3, 1, 5)):
        cfg = recipes.find_config(problem, dim, nz, dof)
        assert cfg is not None
        assert _lib.lib().tfb_config_name(cfg.cid).decode() == cfg.name


def test_host_pipeline_pieces_cover_the_slab():
    '''The z-pieces over which tfb_jacobian pipelines upload, assembly and download (csrc/tfb_core.cu, tfb_pipe_pieces):
    strictly increasing boundaries from 0 to nzl for every slab height, small pieces at both ends of tall slabs.'''
    L = _lib.lib()
    out = (ctypes.c_int * 128)()
    for nzl in list(range(1, 200)) + [255, 256, 257, 512, 1000, 1024]:
        n = L.tfb_pipe_pieces_of(nzl, out, 128)
        b = list(out[:n])
        assert n >= 2 and b[0] == 0 and b[-1] == nzl, (nzl, b)
        assert all(y > x for x, y in zip(b, b[1:])), (nzl, b)
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert max(sizes) <= 32 or nzl < 64, (nzl, sizes)
        if nzl >= 96:
            assert sizes[0] == 4 and sizes[-1] == 4, (nzl, sizes)       # short fill and drain of the pipeline
    assert list(out[:L.tfb_pipe_pieces_of(128, out, 128)]) == [0, 4, 16, 32, 64, 96, 112, 124, 128]
    assert L.tfb_pipe_pieces_of(0, out, 128) < 0
