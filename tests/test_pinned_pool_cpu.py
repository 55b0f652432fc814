'''Host logic of the recycled page-locked result vectors (transiflow_b200/_lib.py: PinnedPool): a buffer returns to the
pool only when the last array or view on it is gone, so a vector handed to the caller (the reference mutates them in
place, Continuation.py:169) is never aliased by a later result.  The allocator is stubbed with malloc: no GPU.'''
import ctypes
import gc

import numpy

from transiflow_b200 import _lib


class _StubLib:
    def __init__(self, fail_after=None):
        self.libc = ctypes.CDLL(None)
        self.libc.malloc.restype = ctypes.c_void_p
        self.libc.free.argtypes = [ctypes.c_void_p]
        self.allocs, self.frees, self.fail_after = 0, 0, fail_after

    def tfb_pinned_alloc(self, nbytes, out):
        if self.fail_after is not None and self.allocs >= self.fail_after:
            return -1
        out._obj.value = self.libc.malloc(nbytes)
        self.allocs += 1
        return 0

    def tfb_pinned_free(self, p):
        self.libc.free(p)
        self.frees += 1
        return 0


def test_buffers_are_recycled_only_after_the_last_view_is_gone(monkeypatch):
    stub = _StubLib()
    monkeypatch.setattr(_lib, 'lib', lambda: stub)
    monkeypatch.setattr(_lib, '_LIB', stub)
    pool = _lib.PinnedPool(1000, limit=3)
    a = pool.empty()
    assert a.shape == (1000,) and a.dtype == numpy.float64 and a.flags.writeable
    a[:] = 1.0
    b = pool.empty()
    b[:] = 2.0
    assert stub.allocs == 2 and a.ctypes.data != b.ctypes.data
    view = a[10:20]
    addr_a = a.ctypes.data
    del a
    gc.collect()
    c = pool.empty()                       # `view` still refers to the first buffer: a third one is allocated
    assert stub.allocs == 3 and c.ctypes.data != addr_a
    assert numpy.all(view == 1.0)
    del view
    gc.collect()
    d = pool.empty()                       # now it is free again
    assert stub.allocs == 3 and d.ctypes.data == addr_a
    e = pool.empty()                       # limit reached, nothing free: ordinary numpy memory
    assert stub.allocs == 3 and e.base is None and e.shape == (1000,)
    x = b + c                              # arithmetic results are ordinary arrays
    assert x.base is None
    del b, c, d, e
    gc.collect()
    pool.close()
    assert stub.frees == 3


def test_allocation_failure_falls_back_to_numpy(monkeypatch):
    stub = _StubLib(fail_after=0)
    monkeypatch.setattr(_lib, 'lib', lambda: stub)
    pool = _lib.PinnedPool(64)
    a = pool.empty()
    assert a.shape == (64,) and a.base is None


def _bare_interface(n=12, dim=3, dof=4, row0=0):
    from transiflow_b200.interface import Interface
    it = Interface.__new__(Interface)          # no device: only the host logic around the solver call is exercised
    it.dim, it.dof, it.pressure_row, it.row0, it.n_local, it.n = dim, dof, dim, row0, n, n
    it._ctx = None
    it._pool = None
    it._sync_solver = lambda: None
    return it


def test_pressure_pin_is_applied_in_place_and_undone():
    """rhs[dim] = 0 for the solve (SciPy.py:216) without copying the vector; the caller's array is unchanged afterwards,
    also when the solver raises, and a read-only array is never written."""
    it = _bare_interface()
    seen = {}

    def fake(jac, b, prow):
        seen['b'], seen['prow'], seen['same'] = b.copy(), prow, b
        return b * 2

    it._solve_pinned = fake
    rhs = numpy.arange(1.0, 13.0)
    y = it._solve1(None, rhs)
    assert seen['prow'] == 3 and seen['b'][3] == 0.0 and seen['same'] is rhs
    assert numpy.array_equal(rhs, numpy.arange(1.0, 13.0))          # restored
    assert y[3] == 0.0 and y[4] == 10.0

    def boom(jac, b, prow):
        raise RuntimeError('x')

    it._solve_pinned = boom
    try:
        it._solve1(None, rhs)
    except RuntimeError:
        pass
    assert numpy.array_equal(rhs, numpy.arange(1.0, 13.0))

    it._solve_pinned = fake
    ro = numpy.arange(1.0, 13.0)
    ro.flags.writeable = False
    it._solve1(None, ro)
    assert seen['same'] is not ro and seen['b'][3] == 0.0 and ro[3] == 4.0

    # the slab that does not own the pinned row passes the vector through untouched
    it2 = _bare_interface(row0=48)
    it2._solve_pinned = fake
    it2._solve1(None, rhs)
    assert seen['prow'] == 3 and numpy.array_equal(seen['b'], rhs) and seen['same'] is rhs

    # no pressure unknown (dof == dim): nothing is pinned
    it3 = _bare_interface(dim=3, dof=3)
    it3._solve_pinned = fake
    it3._solve1(None, rhs)
    assert seen['prow'] == -1 and numpy.array_equal(seen['b'], rhs)
