'''Wall-clock split of Newton steps at 128^3 (diagnostic script, not a test): python tools/newton_probe.py [grid]'''
import sys, time, numpy
sys.path.insert(0, '.')
import transiflow_b200 as tb

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
it = tb.Interface(p, g, g, g)
x = it.vector()
for k in range(6):
    t = [time.perf_counter()]
    jac, f = it.jacobian_rhs(x); t.append(time.perf_counter())
    nf = -f; t.append(time.perf_counter())
    dx = it.solve(jac, nf); t.append(time.perf_counter())
    x = x + dx; t.append(time.perf_counter())
    d = numpy.diff(t) * 1e3
    print('step %d: jacobian_rhs %.1f  neg %.1f  solve(wall) %.1f [device %.1f, its %d]  update %.1f ms'
          % (k, d[0], d[1], d[2], it.last_solve['solve_ms'], it.last_solve['iterations'], d[3]), flush=True)
