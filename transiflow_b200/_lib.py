'''ctypes binding of libtfb200.so (the C ABI declared in include/tfb200.h).

The product path has NO CPU fallback: if the CUDA library is missing or no device is
visible, constructing an Interface raises.'''
import ctypes
import os

import numpy

from .hostprep import TFB_MAX_FORCE, TfbParams

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'lib', 'libtfb200.so')
_LIB = None

c_double_p = ctypes.POINTER(ctypes.c_double)


class TfbDesc(ctypes.Structure):
    _fields_ = [
        ('config', ctypes.c_int32), ('nx', ctypes.c_int32), ('ny', ctypes.c_int32), ('nz', ctypes.c_int32),
        ('dim', ctypes.c_int32), ('dof', ctypes.c_int32), ('device', ctypes.c_int32),
        ('k0', ctypes.c_int32), ('k1', ctypes.c_int32),
        ('met', ctypes.c_void_p * 3), ('cor', ctypes.c_void_p),
    ]


class TfbSolveOpts(ctypes.Structure):
    _fields_ = [
        ('tol', ctypes.c_double), ('maxit', ctypes.c_int32), ('restart', ctypes.c_int32),
        ('pressure_row', ctypes.c_int32), ('precond', ctypes.c_int32), ('verbose', ctypes.c_int32),
        ('basis_fp32', ctypes.c_int32), ('method', ctypes.c_int32), ('idr_s', ctypes.c_int32),
        ('precond_flags', ctypes.c_int32), ('inner_its', ctypes.c_int32), ('stall_cycles', ctypes.c_int32),
    ]


METHOD_FGMRES, METHOD_BICGSTAB, METHOD_IDR = 0, 1, 2
PREC_FP32, PREC_NO_JOINT, PREC_SCALED_MASS, PREC_TENSOR = 1, 2, 8, 16


class TfbSolveInfo(ctypes.Structure):
    _fields_ = [
        ('iters', ctypes.c_int32), ('converged', ctypes.c_int32), ('relres', ctypes.c_double),
        ('setup_ms', ctypes.c_float), ('solve_ms', ctypes.c_float),
    ]


EXPORTS = [
    'tfb_device_count', 'tfb_last_error', 'tfb_config_name', 'tfb_create', 'tfb_destroy', 'tfb_set_params',
    'tfb_sizes', 'tfb_get_pattern', 'tfb_mat_create', 'tfb_mat_destroy', 'tfb_mat_get_values',
    'tfb_mat_set_values', 'tfb_mat_add_diag', 'tfb_mat_set_shift', 'tfb_rhs', 'tfb_jacobian', 'tfb_mass_diag', 'tfb_state_upload',
    'tfb_assemble_resident', 'tfb_rhs_download', 'tfb_host_checksum', 'tfb_upload_count', 'tfb_sync', 'tfb_event_record', 'tfb_event_elapsed_ms',
    'tfb_flush_l2', 'tfb_pipe_pieces_of', 'tfb_pinned_alloc', 'tfb_pinned_free', 'tfb_launch_count', 'tfb_spmv', 'tfb_spmv_bench', 'tfb_solve', 'tfb_direct_solve', 'tfb_fdm_set', 'tfb_fdm_set_pencil', 'tfb_fdm_pin', 'tfb_joint_set', 'tfb_joint_apply', 'tfb_precond_apply', 'tfb_precond_apply_opts', 'tfb_nccl_unique_id', 'tfb_comm_init',
]


def lib():
    '''Load libtfb200.so (built in-tree by ``python -m transiflow_b200.build``).'''
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            raise RuntimeError('libtfb200.so is missing: run `python -m transiflow_b200.build` '
                               '(there is no CPU fallback for the B200 backend)')
        L = ctypes.CDLL(SO)
        L.tfb_last_error.restype = ctypes.c_char_p
        L.tfb_config_name.restype = ctypes.c_char_p
        L.tfb_launch_count.restype = ctypes.c_int64
        L.tfb_upload_count.restype = ctypes.c_int64
        L.tfb_upload_count.argtypes = [ctypes.c_void_p]
        for name in EXPORTS:
            if name not in ('tfb_last_error', 'tfb_config_name', 'tfb_launch_count', 'tfb_upload_count', 'tfb_destroy', 'tfb_mat_destroy'):
                getattr(L, name).restype = ctypes.c_int
        L.tfb_destroy.restype = None
        L.tfb_mat_destroy.restype = None
        L.tfb_destroy.argtypes = [ctypes.c_void_p]
        L.tfb_mat_destroy.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def device_count():
    return lib().tfb_device_count()


def check(rc):
    if rc < 0:
        raise RuntimeError('libtfb200: ' + lib().tfb_last_error().decode())
    return rc


def ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(None)


def as_f64(a):
    return numpy.ascontiguousarray(a, dtype=numpy.float64)


def pinned_array(n):
    """fp64 numpy array of length n backed by page-locked host memory."""
    p = ctypes.c_void_p()
    check(lib().tfb_pinned_alloc(ctypes.c_size_t(8 * n), ctypes.byref(p)))
    buf = (ctypes.c_double * n).from_address(p.value)
    arr = numpy.frombuffer(buf, dtype=numpy.float64)
    return arr


class _PinnedBlock:
    """One page-locked buffer lent out as the base object of a numpy array; when the last array (or view) referring to it
    is collected the buffer goes back to its pool -- never earlier, so a returned vector is not aliased by a later one."""

    def __init__(self, pool, address, n):
        self._pool, self._address = pool, address
        self.__array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (address, False), 'version': 3}

    def __del__(self):
        pool = self._pool
        if pool is not None:
            pool._free.append(self._address)


class PinnedPool:
    """Recycled page-locked result vectors of one length (the device-to-host copy of a result into fresh pageable memory
    costs 8 ms of staging + page faults per 67 MB vector at 128^3, against 1.3 ms into a pinned buffer).  At most `limit`
    buffers are kept; beyond that, and if the allocation fails, `empty()` returns an ordinary numpy array."""

    def __init__(self, n, limit=6):
        self.n, self.limit = n, limit
        self._free, self._owned = [], []

    def empty(self):
        if not self._free and len(self._owned) < self.limit:
            p = ctypes.c_void_p()
            if lib().tfb_pinned_alloc(ctypes.c_size_t(8 * max(self.n, 1)), ctypes.byref(p)) == 0 and p.value:
                self._owned.append(p.value)
                self._free.append(p.value)
        if not self._free:
            return numpy.empty(self.n)
        return numpy.asarray(_PinnedBlock(self, self._free.pop(), self.n))

    def close(self):
        """Frees the buffers that are not lent out; lent ones are released with the process."""
        if _LIB is not None:
            for a in self._free:
                _LIB.tfb_pinned_free(ctypes.c_void_p(a))
                self._owned.remove(a)
        self._free = []
