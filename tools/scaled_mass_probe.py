'''Accuracy / robustness of IDR with the scaled-mass Schur complement against the LSC default (diagnostic script):
python tools/scaled_mass_probe.py'''
import sys, time, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb

SM = {'Method': 'IDR', 'Schur Complement': 'Scaled Mass'}

# 1. Newton at 20^3 / 16^3 (SuperLU feasible): converged state against the spsolve path (tests/golden/make_golden_newton_oracle.py)
import os
GEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'generated')
for name, params in (('ldc3d_20_re100', {'Reynolds Number': 100}), ('ldc3d_16_re400_str', {'Reynolds Number': 400, 'Grid Stretching Factor': 1.5})):
    gold = numpy.load(os.path.join(GEN, 'newton_oracle_' + name + '.npz'))
    N, xr = int(gold['N']), gold['x']
    for label, opts in (('LSC 1e-10', {'Method': 'IDR'}), ('scaled mass 1e-10', SM), ('scaled mass 1e-12', dict(SM, **{'Convergence Tolerance': 1e-12}))):
        it = tb.Interface(dict(params, **{'Iterative Solver': dict(opts)}), N, N, N)
        x = it.vector()
        log = []
        for k in range(12):
            f = it.rhs(x)
            if numpy.linalg.norm(f) < 1e-10:
                break
            jac = it.jacobian(x)
            x = x + it.solve(jac, -f)
            log.append('%d' % it.last_solve['iterations'])
        print('%s %-18s: %d Newton steps, |F| %.1e, state err vs spsolve Newton %.2e; its: %s' % (
            name, label, k, numpy.linalg.norm(it.rhs(x)), numpy.abs(x - xr).max() / numpy.abs(xr).max(), ' '.join(log)), flush=True)

# 2. Newton sequence at 128^3 (and 64^3 at Re 400) with both Schur complements
for params, g in (({'Reynolds Number': 100, 'Lid Velocity': 1}, 128), ({'Reynolds Number': 400, 'Lid Velocity': 1}, 64)):
    sols = {}
    for name, opts in (('LSC', {'Method': 'IDR'}), ('scaled mass', SM), ('scaled mass 1e-12', dict(SM, **{'Convergence Tolerance': 1e-12}))):
        it = tb.Interface(dict(params, **{'Iterative Solver': dict(opts)}), g, g, g)
        x = it.vector()
        for k in range(6):
            t0 = time.perf_counter()
            jac, f = it.jacobian_rhs(x)
            dx = it.solve(jac, -f)
            x = x + dx
            print('%d^3 Re %g %-18s step %d: |F| %.2e  %4d its %7.1f ms (wall %.1f)  relres %.1e %s' % (
                g, params['Reynolds Number'], name, k, numpy.linalg.norm(f), it.last_solve['iterations'], it.last_solve['solve_ms'],
                1e3 * (time.perf_counter() - t0), it.last_solve['relres'], it.last_solve['converged']), flush=True)
        sols[name] = x
        del it, jac
    for name in sols:
        print('%d^3 state after 6 steps, %s vs LSC: %.2e' % (g, name, numpy.abs(sols[name] - sols['LSC']).max() / numpy.abs(sols['LSC']).max()), flush=True)
