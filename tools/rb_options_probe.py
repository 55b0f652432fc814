'''Rayleigh-Benard 128^3: solver options on the second Newton system after the perturbed conduction state (diagnostic
script, same set-up as bench.py --problem rb): python tools/rb_options_probe.py [grid]'''
import sys, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}
it = tb.Interface(p, g, g, g)
x = it.vector()
jac, f = it.jacobian_rhs(x)
x = x + it.solve(jac, -f)
c3 = numpy.indices((g, g, g)).astype(float)
roll = numpy.sin(numpy.pi * (c3[0] + 1) / g) * numpy.cos(6 * numpy.pi * (c3[2] + 0.5) / g)
xs = x.reshape(g, g, g, it.dof)
xs[..., 2] += 1e-2 * roll
xs[-1, :, :, 2] = 0.0
xs[..., 4] += 1e-2 * roll
jac, f = it.jacobian_rhs(x)
x = x + it.solve(jac, -f)
jac, f = it.jacobian_rhs(x)
ref = None
import warnings
warnings.simplefilter('ignore')
for opts in ({}, {'Velocity Iterations': 4}, {'Velocity Iterations': 0},
             {'Velocity Iterations': 0, 'Method': 'IDR'},
             {'Velocity Iterations': 0, 'Method': 'IDR', 'Schur Complement': 'Scaled Mass'},
             {'Velocity Iterations': 0, 'Schur Complement': 'Scaled Mass', 'Method': 'FGMRES'},
             {'Velocity Iterations': 4, 'Schur Complement': 'Scaled Mass'},
             {'Schur Complement': 'Scaled Mass'},
             {'Preconditioner Precision': 'double'}):
    it.parameters['Iterative Solver'] = dict(opts)
    dx = it.solve(jac, -f)
    if ref is None:
        ref = dx
    print('%-50s %4d its %8.1f ms  relres %.2e  conv %s  %s |dx - dx_ref|/|dx| %.1e' % (
        opts, it.last_solve['iterations'], it.last_solve['solve_ms'], it.last_solve['relres'], it.last_solve['converged'],
        it.last_solve['method'], numpy.abs(dx - ref).max() / numpy.abs(ref).max()), flush=True)
