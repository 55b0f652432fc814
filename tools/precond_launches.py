'''One short IDR solve at 128^3 for an ncu launch list of the preconditioner (diagnostic script):
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file out.csv python tests/precond_launches.py'''
import sys, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb
g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1,
     'Iterative Solver': {'Method': sys.argv[2] if len(sys.argv) > 2 else 'IDR', 'Maximum Iterations': 9}}
it = tb.Interface(p, g, g, g)
x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
jac, f = it.jacobian_rhs(x)
it.solve(jac, -f)
print(it.last_solve)
