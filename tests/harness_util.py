'''Drives tests/cpu_harness (the generated row functions compiled with g++) from Python.'''
import ctypes
import os
import subprocess

import numpy

from transiflow_b200 import hostprep, recipes

HERE = os.path.dirname(os.path.abspath(__file__))
HDIR = os.path.join(HERE, 'cpu_harness')
_LIB = None


class TfbGrid(ctypes.Structure):
    _fields_ = [
        ('nx', ctypes.c_int), ('ny', ctypes.c_int), ('nz', ctypes.c_int), ('dim', ctypes.c_int), ('dof', ctypes.c_int),
        ('zfold', ctypes.c_int),
        ('met', ctypes.c_void_p * 3), ('cor', ctypes.c_void_p),
        ('fval', ctypes.c_void_p * hostprep.TFB_MAX_FORCE),
        ('fdir', ctypes.c_byte * hostprep.TFB_MAX_FORCE),
    ]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HDIR, 'libtfharness.so')
        src = os.path.join(HDIR, 'harness.cpp')
        gen = os.path.join(os.path.dirname(HERE), 'transiflow_b200', 'csrc', 'gen', 'all_configs.h')
        newest = max(os.path.getmtime(p) for p in (src, gen))
        if not os.path.exists(so) or os.path.getmtime(so) < newest:
            subprocess.check_call(['/usr/bin/g++', '-O1', '-fPIC', '-shared', '-ffp-contract=off', '-std=c++17',
                                   '-w', '-o', so, src])
        _LIB = ctypes.CDLL(so)
        _LIB.tfh_assemble.restype = ctypes.c_int
    return _LIB


def assemble(parameters, nx, ny, nz, dim, dof, state, x=None, y=None, z=None, boundary_conditions=None):
    '''Returns (vals, cols, row_ptr, rhs) of the fixed structural pattern (explicit zeros kept).  With a user
    ``boundary_conditions`` callback the kernel family is the one whose recipe has the structure of the recorded ops
    (recipes.match_recorded), exactly as Interface.__init__ chooses it.'''
    p = parameters
    problem = recipes.PROBLEM_IDS[p.get('Problem Type', 'Lid-driven Cavity').lower()]
    if boundary_conditions is not None:
        cfg = recipes.match_recorded(recipes.record_boundary_conditions(boundary_conditions), dim, nz, dof)
    else:
        cfg = recipes.find_config(problem, dim, nz, dof)
    assert cfg is not None, 'no generated config'
    x = hostprep.coordinate_vector(p, p.get('X-min', 0.0), p.get('X-max', 1.0), nx) if x is None else x
    y = hostprep.coordinate_vector(p, p.get('Y-min', 0.0), p.get('Y-max', 1.0), ny) if y is None else y
    z = hostprep.coordinate_vector(p, p.get('Z-min', 0.0), p.get('Z-max', 1.0), nz) if z is None else z
    mets = [hostprep.axis_metrics(v, m) for v, m in ((x, nx), (y, ny), (z, nz))]
    cor = hostprep.coriolis_metrics(y, ny)
    prm, arrays = hostprep.make_params(cfg, problem, p, nx, ny, nz, x, y, z)
    g = TfbGrid()
    g.nx, g.ny, g.nz, g.dim, g.dof, g.zfold = nx, ny, nz, dim, dof, int(nz == 1)
    for a in range(3):
        g.met[a] = mets[a].ctypes.data
    g.cor = cor.ctypes.data
    fi = 0
    for op in cfg.recipe:
        if op[0] == 'force':
            g.fdir[fi] = op[1]
            if fi in arrays:
                g.fval[fi] = arrays[fi].ctypes.data
            fi += 1
    wind = None
    if problem == recipes.QG:
        wind = hostprep.wind_stress(p, nx, ny, nz, dof, x, y, z)
    n = nx * ny * nz * dof
    cap = 32 * n
    row_ptr = numpy.zeros(n + 1, dtype=numpy.int64)
    col = numpy.zeros(cap, dtype=numpy.int64)
    val = numpy.zeros(cap)
    rhs = numpy.zeros(n)
    state = numpy.ascontiguousarray(state, dtype=numpy.float64)
    rc = lib().tfh_assemble(cfg.cid, ctypes.byref(g), ctypes.byref(prm), ctypes.c_void_p(state.ctypes.data),
                            ctypes.c_void_p(wind.ctypes.data if wind is not None else None),
                            ctypes.c_void_p(row_ptr.ctypes.data), ctypes.c_void_p(col.ctypes.data),
                            ctypes.c_void_p(val.ctypes.data), ctypes.c_void_p(rhs.ctypes.data), ctypes.c_int64(cap))
    assert rc == 0, 'harness rc=%d' % rc
    nnz = row_ptr[-1]
    return val[:nnz], col[:nnz], row_ptr, rhs
