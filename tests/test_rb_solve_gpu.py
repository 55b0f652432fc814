'''Rayleigh-Benard linear solves on the device: the coupled (w, T) line solve against its numpy mirror
(tests/test_joint_cpu.py) and Newton updates at sub- and super-critical Rayleigh numbers against the pinned
SuperLU solve of the SciPy backend (oracle.direct_solve restates SciPy.py:204-258).'''
import ctypes

import numpy
import pytest

from test_joint_cpu import JointModel

pytestmark = pytest.mark.gpu

RB = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
      'X-max': 10, 'Y-max': 10}


def _conduction_state(orc):
    from oracle.tf_oracle import direct_solve
    x0 = numpy.zeros(orc.n)
    return x0 + direct_solve(orc.jacobian_csr(x0), -orc.rhs(x0), orc.dim, orc.dof)


def test_device_joint_solve_matches_numpy_mirror():
    from transiflow_b200 import Interface, _lib
    nx, ny, nz = 10, 7, 9
    model = JointModel(RB, nx, ny, nz)
    it = Interface(dict(RB), nx, ny, nz)
    it._sync_solver()
    jac = it.jacobian(model.state)
    rng = numpy.random.default_rng(0)
    r = rng.standard_normal(it.n)
    z = numpy.zeros(it.n)
    table = numpy.zeros((12, nz))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().tfb_joint_apply(jac._h, P(r), P(z), P(table)))
    assert numpy.allclose(table[:8], model.zc[:8], rtol=0, atol=0)
    assert numpy.allclose(table[8:], model.zc[8:], rtol=1e-12, atol=1e-14)
    T = model.cfg.T
    w, t = model.solve(r[2::5].copy(), r[T::5].copy())
    assert numpy.abs(z[2::5] - w).max() <= 1e-10 * numpy.abs(w).max()
    assert numpy.abs(z[T::5] - t).max() <= 1e-10 * numpy.abs(t).max()
    assert not z[0::5].any() and not z[1::5].any() and not z[3::5].any()


@pytest.mark.parametrize('Ra,limit', [(100.0, 120), (1000.0, 120), (3000.0, 400)])
def test_rb_newton_update_matches_superlu(Ra, limit):
    '''J(conduction state) dx = b: the block-triangular preconditioner stalls beyond Ra ~ 500;
    the coupled solve must reach 1e-10 and SuperLU's update.'''
    from oracle.tf_oracle import Oracle, direct_solve
    from transiflow_b200 import Interface
    nx, ny, nz = 16, 16, 8
    params = dict(RB)
    params['Rayleigh Number'] = Ra
    orc = Oracle(dict(params), nx, ny, nz)
    x = _conduction_state(orc)
    it = Interface(dict(params), nx, ny, nz)
    jac = it.jacobian(x)
    b = numpy.random.default_rng(0).standard_normal(it.n)
    b[3] = 0
    y = it.solve(jac, b)
    assert it.last_solve['converged'], it.last_solve
    assert it.last_solve['iterations'] <= limit, it.last_solve
    want = direct_solve(orc.jacobian_csr(x), b, orc.dim, orc.dof)
    assert numpy.abs(y - want).max() <= 1e-8 * numpy.abs(want).max(), it.last_solve


def test_rb_newton_returns_to_the_conduction_state():
    '''Newton at Ra = 1000 from the conduction state plus a finite roll-like perturbation (the regime where the
    block-triangular preconditioner stalls): every linear solve converges and the iteration ends on the oracle's
    conduction state.'''
    from oracle.tf_oracle import Oracle
    from transiflow_b200 import Interface
    nx, ny, nz = 16, 16, 8
    orc = Oracle(dict(RB), nx, ny, nz)
    want = _conduction_state(orc)
    k3, _, i3 = numpy.indices((nz, ny, nx)).astype(float)
    roll = numpy.sin(numpy.pi * (k3 + 1) / nz) * numpy.cos(4 * numpy.pi * (i3 + 0.5) / nx)
    x = want.copy().reshape(nz, ny, nx, 5)
    x[..., 2] += 1e-2 * roll
    x[-1, :, :, 2] = 0
    x[..., 4] += 1e-2 * roll
    x = x.ravel()
    it = Interface(dict(RB), nx, ny, nz)
    for k in range(8):
        f = it.rhs(x)
        if numpy.linalg.norm(f) < 1e-10:
            break
        x = x + it.solve(it.jacobian(x), -f)
        assert it.last_solve['converged'], (k, it.last_solve)
    assert numpy.linalg.norm(orc.rhs(x)) < 1e-9
    dx = x - want
    dx[3::5] -= dx[3]                      # pressure is determined up to its pinned constant
    assert numpy.abs(dx).max() <= 1e-8 * numpy.abs(want).max()


def test_block_triangular_option_still_available():
    from transiflow_b200 import Interface
    params = dict(RB)
    params['Rayleigh Number'] = 100.0
    params['Iterative Solver'] = {'Scalar Coupling': 'none'}
    it = Interface(params, 12, 12, 6)
    x = numpy.zeros(it.n)
    y = it.solve(it.jacobian(x), -it.rhs(x))
    assert it.last_solve['converged'], it.last_solve
    assert numpy.linalg.norm(it.rhs(x + y)) < 1e-8      # linear at u = 0: one step reaches the conduction state
