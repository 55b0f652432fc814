'''The reference's own drivers -- Continuation.newton / Continuation.continuation (Continuation.py:69-113,
:362-470) and TimeIntegration.integration (TimeIntegration.py:40-115) -- run UNMODIFIED on top of the B200
Interface.  They come from the pip-installed reference under baseline/_ref (git-ignored, travels to the GPU
box with the snapshot; `python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref
<copy of /root/reference>`); when it is absent the tests skip.  Goldens: the same drivers on the reference's SciPy
backend (tests/golden/make_golden_continuation.py, make_golden_time.py).'''
import contextlib
import io
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, 'tests', 'golden', 'generated')
REF = os.path.join(ROOT, 'baseline', '_ref')

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def transiflow():
    if not os.path.isdir(os.path.join(REF, 'transiflow')):
        pytest.skip('reference package not installed under baseline/_ref')
    sys.path.insert(0, REF)
    try:
        import transiflow as tf
        yield tf
    finally:
        sys.path.remove(REF)


LDC = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 0, 'Lid Velocity': 1, 'Grid Stretching Factor': 1.5}


@pytest.mark.parametrize('bordered', [False, True])
def test_reference_continuation_drives_the_b200_interface(transiflow, bordered):
    """Pseudo-arclength continuation in the Reynolds number, 0 -> 400 on the stretched 16x16 cavity
    (the reference's tests/test_continuation.py:75-78 at a smaller grid).  Every corrector step calls
    set_parameter, rhs (twice), jacobian and solve (twice, or once with the border) on the device backend.
    A 1e-9 relative perturbation of every solve moves the reference's own end point by 7e-11, so the
    device solver's 1e-10 residual tolerance has to land within 1e-6 of the golden end point."""
    g = numpy.load(os.path.join(GEN, 'continuation_ldc2d.npz'))
    from transiflow_b200 import Interface
    it = Interface(dict(LDC), int(g['nx']), int(g['ny']))
    cont = transiflow.Continuation(it, bordered_solver=bordered)
    with contextlib.redirect_stdout(io.StringIO()):
        x0 = cont.newton(it.vector())
        x, mu = cont.continuation(x0, 'Reynolds Number', 0, float(g['mu']), 100)
    scale = numpy.abs(g['x']).max()
    assert abs(mu - float(g['mu'])) <= 1e-4                      # Continuation's destination_tolerance
    assert it.get_parameter('Reynolds Number') == float(g['mu'])
    assert numpy.abs(x - g['x_continuation']).max() <= 1e-6 * scale
    x = cont.newton(x, 1e-11)
    assert numpy.abs(x - g['x']).max() <= 1e-8 * scale


def test_corrector_iterations_upload_the_state_once(transiflow):
    """Device-resident continuation fast path: a corrector iteration of the UNMODIFIED Continuation
    evaluates rhs(x; mu), rhs(x; mu + delta) and jacobian(x) for the same x (Continuation.py:126,145-150); the state is
    uploaded once for the three, recognised by an exact checksum, so in-place edits of single entries are still seen."""
    from transiflow_b200 import Interface
    it = Interface(dict(LDC), 16, 16)
    x = 0.01 * numpy.random.default_rng(0).standard_normal(it.n)
    calls = {'rhs': 0, 'jacobian': 0}
    rhs0, jac0 = it.rhs, it.jacobian
    it.rhs = lambda s: (calls.__setitem__('rhs', calls['rhs'] + 1), rhs0(s))[1]
    it.jacobian = lambda s: (calls.__setitem__('jacobian', calls['jacobian'] + 1), jac0(s))[1]
    cont = transiflow.Continuation(it)
    with contextlib.redirect_stdout(io.StringIO()):
        x0 = cont.newton(x.copy())
        before, calls['rhs'], calls['jacobian'] = it.state_uploads, 0, 0
        cont.continuation(x0, 'Reynolds Number', 0, 20, 10)
    uploads = it.state_uploads - before
    assert calls['jacobian'] > 0 and calls['rhs'] >= 2 * calls['jacobian']
    # one upload per distinct state: every corrector iteration has one (three evaluations), plus the converged checks
    assert uploads <= calls['rhs'] - calls['jacobian'], (uploads, calls)
    # an in-place edit of a single entry is a different state
    f0 = rhs0(x0)
    x0[7] += 1e-9
    n0 = it.state_uploads
    f1 = rhs0(x0)
    assert it.state_uploads == n0 + 1 and not numpy.array_equal(f0, f1)
    # opting out restores one upload per call
    it.parameters['State Cache'] = False
    n0 = it.state_uploads
    rhs0(x0), rhs0(x0)
    assert it.state_uploads == n0 + 2


def test_reference_time_integration_drives_the_b200_interface(transiflow):
    """Implicit Euler of the reference (theta = 1): mass_matrix() @ v, jacobian(x) - mass / (theta dt) on the
    DeviceMatrix, solve."""
    g = numpy.load(os.path.join(GEN, 'time_ldc2d.npz'))
    from transiflow_b200 import Interface
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
    it = Interface(params, int(g['nx']), int(g['ny']))
    ti = transiflow.TimeIntegration(it, theta=1.0)
    dt, steps = float(g['dt']), int(g['steps'])
    with contextlib.redirect_stdout(io.StringIO()):
        x, t = ti.integration(it.vector(), dt, dt * steps - 1e-9)
    assert abs(t - float(g['t'])) <= 1e-12
    assert numpy.abs(x - g['x']).max() <= 1e-8 * numpy.abs(g['x']).max()


def test_reference_newton_on_a_3d_grid(transiflow):
    """Continuation.newton on the 3-D cavity at Re = 100: converges to |F| < 1e-10 in the reference's
    iteration limit, through the IDR / FGMRES device solves."""
    from transiflow_b200 import Interface
    it = Interface({'Reynolds Number': 100, 'Lid Velocity': 1}, 16, 16, 16)
    cont = transiflow.Continuation(it)
    with contextlib.redirect_stdout(io.StringIO()):
        x = cont.newton(it.vector(), 1e-10)
    assert numpy.linalg.norm(it.rhs(x)) < 1e-10
    assert cont.newton_iterations < cont.maximum_newton_iterations - 1
