'''Solver-option scan on one Newton system of the 3-D cavity (diagnostic script, not a test):
python tools/solve_scan.py [grid]'''
import sys, time, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': float(__import__('os').environ.get('RE', 100)), 'Lid Velocity': 1}
if __import__('os').environ.get('STRETCH'):
    p['Grid Stretching Factor'] = 1.5
it = tb.Interface(p, g, g, g)
x = it.vector()
for k in range(2):
    jac, f = it.jacobian_rhs(x)
    x = x + it.solve(jac, -f)
jac, f = it.jacobian_rhs(x)
ref = it.solve(jac, -f)
variants = [{}, {'Schur Complement': 'Scaled Mass'}, {'Schur Complement': 'Scaled Mass', 'Method': 'FGMRES'}]
if len(sys.argv) > 2:
    variants = eval(sys.argv[2])
p['Verbose'] = True
for v in variants:
    p['Iterative Solver'] = v
    t0 = time.perf_counter()
    y = it.solve(jac, -f)
    w = (time.perf_counter() - t0) * 1e3
    err = numpy.abs(y - ref).max() / numpy.abs(ref).max()
    print('%-55s its %4d  device %.1f ms  wall %.1f ms  relres %.2e  diff %.1e' % (v, it.last_solve['iterations'],
          it.last_solve['solve_ms'], w, it.last_solve['relres'], err), flush=True)
