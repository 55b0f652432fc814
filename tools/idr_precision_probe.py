'''IDR(8) with the FDM sub-solves of the preconditioner in fp64 / fp32 / TF32 on the third Newton system of the 128^3 cavity
(diagnostic script, not a test): python tests/idr_precision_probe.py [grid]'''
import sys, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
it = tb.Interface(p, g, g, g)
x = it.vector()
for k in range(2):
    jac, f = it.jacobian_rhs(x)
    x = x + it.solve(jac, -f)
jac, f = it.jacobian_rhs(x)
ref = None
for opts in ({'Method': 'IDR'}, {'Method': 'IDR', 'Preconditioner Precision': 'single'},
             {'Method': 'IDR', 'Preconditioner Precision': 'tf32'}, {'Method': 'IDR', 'IDR Dimension': 4, 'Preconditioner Precision': 'single'},
             {'Method': 'IDR', 'Schur Complement': 'Scaled Mass'}, {'Method': 'IDR', 'Schur Complement': 'Scaled Mass', 'Preconditioner Precision': 'single'},
             {'Method': 'FGMRES'}):
    it.parameters['Iterative Solver'] = dict(opts)
    for rep in range(2):
        dx = it.solve(jac, -f)
    if ref is None:
        ref = dx
    print('%-90s %4d its %8.1f ms  relres %.2e  conv %s  |dx - dx_ref|/|dx| %.1e' % (
        opts, it.last_solve['iterations'], it.last_solve['solve_ms'], it.last_solve['relres'], it.last_solve['converged'],
        numpy.abs(dx - ref).max() / numpy.abs(ref).max()), flush=True)
