'''The Interface surface callers use besides assembly and solve, on a device: state / parameter files (port of the
reference's tests/test_interface.py:87-115), the attributes the JaDa glue and utils read, and the error behaviour.'''
import os
import warnings

import numpy
import pytest

pytestmark = pytest.mark.gpu


def test_save_load(tmp_path, nx=4):
    '''reference tests/test_interface.py:87-115 (complex parameter values survive, the caller's dict is updated)'''
    from transiflow_b200 import Interface
    parameters = {'Eigenvalue Solver': {'Target': 1 + 3j}}
    interface = Interface(parameters, nx, nx, nx)
    x = interface.vector_from_array(numpy.random.random(interface.vector().size))
    name = str(tmp_path / 'x-test')
    interface.save_state(name, x)
    assert os.path.isfile(name + '.npy')
    assert os.path.isfile(name + '.params')
    parameters['Eigenvalue Solver']['Target'] = 33
    x2 = interface.load_state(name)
    assert parameters['Eigenvalue Solver']['Target'] == 1 + 3j
    assert numpy.linalg.norm(x - x2) < 1e-14
    # a name that already carries the extension keeps `name + '.params'` for the parameters (BaseInterface.py:128-160)
    interface.save_state(name + '.npy', x)
    assert os.path.isfile(name + '.npy.params')
    with pytest.raises(FileNotFoundError):
        interface.load_state(str(tmp_path / 'missing'))


def test_matrix_handle_and_discretization_attributes():
    from transiflow_b200 import Interface
    it = Interface({'Reynolds Number': 10}, 5, 4, 3)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac = it.jacobian(x)
    assert jac.data.dtype == numpy.float64 and jac.dtype == numpy.float64 and jac.shape == (it.n, it.n)
    assert numpy.array_equal(jac.data, jac.tocsr().data)
    d = it.discretization
    assert (d.x_periodic, d.y_periodic, d.z_periodic) == (False, False, False)
    assert d.nx == 5 and d.dof == 4 and d.x is it.x
    assert Interface({}, 4, 4, 1).discretization.z_periodic       # Discretization.py:127-128
    it.set_parameter('Reynolds Number', 20)
    assert it.get_parameter('Reynolds Number') == 20 and it.get_parameter('missing') == 0 and it.get_parameter('missing', 3) == 3


def test_unconverged_solve_warns_and_returns_the_best_iterate():
    from transiflow_b200 import Interface
    it = Interface({'Reynolds Number': 100, 'Iterative Solver': {'Maximum Iterations': 3}}, 8, 8, 8)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, f = it.jacobian_rhs(x)
    with pytest.warns(RuntimeWarning, match='relative residual'):
        y = it.solve(jac, -f)
    assert not it.last_solve['converged'] and numpy.all(numpy.isfinite(y))
    it.parameters['Iterative Solver'] = {}
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        it.solve(jac, -f)
    assert it.last_solve['converged']


def test_small_restart_keeps_iterating_instead_of_declaring_stagnation():
    '''A user-set small 'Restart' must not stop after two cycles far from the tolerance.'''
    from transiflow_b200 import Interface
    it = Interface({'Reynolds Number': 100, 'Iterative Solver': {'Restart': 20, 'Method': 'FGMRES'}}, 12, 12, 12)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, f = it.jacobian_rhs(x)
    it.solve(jac, -f)
    assert it.last_solve['converged'], it.last_solve


def test_host_layer_refuses_z_slabs():
    from transiflow_b200 import Interface
    it = Interface({'Reynolds Number': 10}, 4, 4, 6, slab=(0, 3))
    for call in (it.mass_matrix, lambda: it.eigs(it.vector())):
        with pytest.raises(NotImplementedError, match='z-slab'):
            call()
