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
