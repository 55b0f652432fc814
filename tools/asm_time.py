'''Times the fused Jacobian+RHS launch (state resident, L2 flushed) for the kernel variant selected by
TFB_ASM_VARIANT.  Measurement script, not a test:  python tools/asm_time.py [ldc|rb] [grid]'''
import ctypes
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transiflow_b200 import DeviceMatrix, Interface, _lib  # noqa: E402
from transiflow_b200._lib import check, ptr  # noqa: E402

problem = sys.argv[1] if len(sys.argv) > 1 else 'ldc'
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
if problem == 'rb':
    params = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
              'X-max': 10.0, 'Y-max': 10.0}
else:
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}
L = _lib.lib()
it = Interface(params, grid, grid, grid)
state = _lib.pinned_array(it.n_local)
state[:] = numpy.random.default_rng(0).uniform(-0.5, 0.5, it.n_local)
it._sync_params()
mat = DeviceMatrix(it)
check(L.tfb_state_upload(it._ctx, ptr(state)))
for _ in range(10):
    check(L.tfb_flush_l2(it._ctx))
    check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
ms_all = []
for _ in range(40):
    check(L.tfb_flush_l2(it._ctx))
    check(L.tfb_event_record(it._ctx, 0))
    check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
    check(L.tfb_event_record(it._ctx, 1))
    ms = ctypes.c_float()
    check(L.tfb_event_elapsed_ms(it._ctx, 0, 1, ctypes.byref(ms)))
    ms_all.append(ms.value)
ms_all.sort()
vals = mat.values()
print('variant %s %s %d^3: median %.4f ms  min %.4f  mean %.4f   checksum %.17g' % (
    os.environ.get('TFB_ASM_VARIANT', '0'), problem, grid, ms_all[len(ms_all) // 2], ms_all[0], sum(ms_all) / len(ms_all),
    float(numpy.abs(vals).sum())))
