'''The algebra of the blocked Gauss-Jordan inversion that csrc/tfb_direct.cu runs on the device (k_gj_blocked: panels
eliminated on their own, one rank-NB update of the other columns per panel, row swaps undone by a column gather at the end),
as its numpy model tools/proto/blocked_gauss_jordan.py, against numpy.linalg.inv.  CPU-only; the device kernel is checked
against SuperLU in tests/test_solve_gpu.py.'''
import importlib.util
import os

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('blocked_gauss_jordan', os.path.join(ROOT, 'tools', 'proto', 'blocked_gauss_jordan.py'))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


@pytest.mark.parametrize('m,nb', [(7, 3), (40, 8), (96, 16), (165, 32), (33, 33)])
def test_blocked_inverse_matches_numpy(m, nb):
    rng = numpy.random.default_rng(m)
    A = rng.standard_normal((m, m))
    for d in range(0, m, 3):
        A[d, d] = 0.0                       # zero diagonal entries, as the pressure rows of the saddle-point blocks have
    inv = model.blocked_inverse(A, nb)
    assert numpy.abs(inv @ A - numpy.eye(m)).max() < 1e-9
    assert numpy.abs(inv - numpy.linalg.inv(A)).max() <= 1e-8 * numpy.abs(inv).max()


def test_panel_elimination_is_the_unblocked_algorithm_on_the_panel_columns():
    '''One panel over all columns is plain Gauss-Jordan with partial pivoting: the pivots are the column maxima among the
    rows not used yet.'''
    rng = numpy.random.default_rng(3)
    A = rng.standard_normal((12, 12))
    P = A.copy()
    piv = model.gj_panel(P, 0)
    B = A.copy()
    for k, p in enumerate(piv):
        assert p == k + int(numpy.argmax(numpy.abs(B[k:, k])))
        B[[k, p]] = B[[p, k]]
        pivrow = B[k] / B[k, k]
        f = B[:, k].copy()
        B -= numpy.outer(f, pivrow)
        B[k] = pivrow
        # (B now holds the reduced matrix, not the in-place inverse: only the pivot choice is compared here)
