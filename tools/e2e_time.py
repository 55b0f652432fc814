'''e2e leg of bench.py alone: Interface.jacobian_rhs_into with pinned host buffers at 128^3 (TFB_PIPE_PLANES / TFB_NO_PIPELINE
select the host pipeline).  python tools/e2e_time.py [grid]'''
import ctypes
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transiflow_b200 import DeviceMatrix, Interface, _lib  # noqa: E402
from transiflow_b200._lib import check  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = _lib.lib()
it = Interface({'Reynolds Number': 100.0, 'Lid Velocity': 1.0}, grid, grid, grid)
state = _lib.pinned_array(it.n_local)
state[:] = numpy.random.default_rng(0).uniform(-0.5, 0.5, it.n_local)
out = _lib.pinned_array(it.n_local)
mat = DeviceMatrix(it)
for _ in range(5):
    it.jacobian_rhs_into(state, mat, out)
reps = 50
check(L.tfb_sync(it._ctx))
t0 = time.perf_counter()
for _ in range(reps):
    it.jacobian_rhs_into(state, mat, out)
wall = (time.perf_counter() - t0) / reps * 1e3
print('TFB_PIPE_PLANES=%s NO_PIPELINE=%s: %.3f ms per call (wall), rhs checksum %.17g' % (
    os.environ.get('TFB_PIPE_PLANES', '-'), os.environ.get('TFB_NO_PIPELINE', '-'), wall, float(numpy.abs(out).sum())))
