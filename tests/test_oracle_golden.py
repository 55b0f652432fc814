'''Pins the CPU oracle (oracle/tf_oracle.c) against (a) the reference's own golden CSR/RHS
files and (b) vectors generated from the unmodified Python reference (tests/golden/make_golden.py).
CPU-only; no /root/reference access.'''
import numpy
import pytest

from cases import CASES, make_state
from golden_io import assert_csr_equal, load_case, read_ref_matrix, read_ref_vector
from oracle.tf_oracle import Oracle, direct_solve


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_matches_generated_golden(name):
    params, nx, ny, nz, dim, dof, kind = CASES[name]
    g = load_case(name)
    orc = Oracle(dict(params), nx, ny, nz, dim, dof)
    assert orc.dim == int(g['dim']) and orc.dof == int(g['dof'])
    for a, b in ((orc.x, g['x']), (orc.y, g['y']), (orc.z, g['z'])):
        # tanh/sin are numpy calls on both sides; allow 1 ulp across CPU generations
        assert numpy.allclose(a, b, rtol=4e-16, atol=0)
    orc.x, orc.y, orc.z = (numpy.ascontiguousarray(g[c]) for c in 'xyz')
    state = make_state(kind, orc.n)
    assert numpy.array_equal(state, g['state'])
    # bit-exact: integer pattern AND fp64 values (same operation order as the reference)
    assert_csr_equal(orc.jacobian(state), (g['coA'], g['jcoA'], g['begA']), 0.0, name + ' jacobian')
    rtol = 0.0 if 'Wind Stress Parameter' not in params and params.get('Problem Type') != 'AMOC' else 1e-14
    f = orc.rhs(state)
    if rtol == 0.0:
        assert numpy.array_equal(f, g['rhs'])
    else:  # cos() of the forcing profiles is evaluated by numpy on this host
        assert numpy.allclose(f, g['rhs'], rtol=1e-13, atol=1e-13 * numpy.abs(g['rhs']).max())
    assert_csr_equal(orc.mass_matrix(), (g['mcoA'], g['mjcoA'], g['mbegA']), 0.0, name + ' mass')


REF_FILES = [
    # (file stem, parameters, dim, dof, nz)  -- parameters from reference tests/test_fvm.py:843-1257
    ('ldc', {'Reynolds Number': 100}, 3, 4, 4),
    ('ldc_stretched', {'Reynolds Number': 100, 'Grid Stretching': True}, 3, 4, 4),
    ('bous', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
              'Problem Type': 'Rayleigh-Benard'}, 3, 5, 4),
    ('bous_stretched', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
                        'Problem Type': 'Rayleigh-Benard', 'Grid Stretching': True}, 3, 5, 4),
    ('dhc', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
             'Problem Type': 'Differentially Heated Cavity'}, 3, 5, 4),
    ('qg', {'Reynolds Number': 16, 'Rossby Parameter': 1000, 'Wind Stress Parameter': 1000,
            'Problem Type': 'Double Gyre'}, 2, 4, 1),
    ('amoc', {'Reynolds Number': 16, 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1,
              'Temperature Forcing': 1, 'Freshwater Flux': 1, 'Problem Type': 'AMOC'}, 2, 5, 1),
]


@pytest.mark.parametrize('stem,params,dim,dof,nz', REF_FILES, ids=[r[0] for r in REF_FILES])
def test_oracle_matches_reference_golden_files(stem, params, dim, dof, nz):
    nx = ny = 4
    orc = Oracle(dict(params), nx, ny, nz, dim, dof)
    state = make_state('lin', orc.n)
    want = read_ref_matrix('%s_%dx%dx%d.txt' % (stem, nx, ny, nz), orc.n)
    # files were written by an earlier implementation with a different operation order:
    # pattern exact, values to 1e-12 (the reference test itself uses pytest.approx, rel 1e-6)
    assert_csr_equal(orc.jacobian(state), want, 1e-12, stem)
    want_rhs = read_ref_vector('%s_rhs_%dx%dx%d.txt' % (stem, nx, ny, nz))
    f = orc.rhs(state)
    assert numpy.all(numpy.abs(f - want_rhs) <= 1e-12 * numpy.maximum(numpy.abs(want_rhs), numpy.abs(f).max() * 1e-3))


def test_direct_solve_restatement():
    '''Manufactured solution through the pinned-pressure direct solve (mirrors
    reference tests/test_SciPy.py:6-30).'''
    params, nx, ny, nz = {'Reynolds Number': 100}, 4, 4, 4
    orc = Oracle(params, nx, ny, nz)
    state = make_state(0, orc.n)
    J = orc.jacobian_csr(state)
    xs = make_state(1, orc.n)
    xs[3::4] -= xs[3]          # pressure is determined up to the pinned constant
    b = J @ xs
    y = direct_solve(J, b, orc.dim, orc.dof)
    assert numpy.allclose(y, xs, rtol=0, atol=1e-9)
