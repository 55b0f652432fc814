'''z-slab partitioning over the GPUs of one box (one process per GPU).

Rows are ordered k-slowest (Discretization.py:513), so a z-slab is a contiguous row range and
needs no permutation (the reference's MPI backends split the domain in
ParallelBaseInterface.py:90-164).  torch.distributed (any backend) is used for plumbing only:
broadcasting the NCCL unique id and max-reducing timings; the data path uses NCCL inside
libtfb200 (halo exchange, Krylov reductions).
'''
import ctypes


def slab_range(nz, world, rank):
    '''Planes [k0, k1) owned by `rank`: as even as possible, earlier ranks get the remainder.'''
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    if nz < world:
        raise ValueError('cannot split %d planes over %d ranks' % (nz, world))
    base, rem = divmod(nz, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def owned_rows(nx, ny, dof, k0, k1):
    '''Global row range [r0, r1) of a slab.'''
    plane = nx * ny * dof
    return k0 * plane, k1 * plane


def broadcast_unique_id(dist, make_id, rank):
    '''Rank 0 creates the 128-byte NCCL unique id, everyone receives it (torch.distributed).'''
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.tensor(list(make_id()), dtype=torch.uint8)
    dist.broadcast(buf, 0)
    return bytes(buf.tolist())


def init_comm(interface, dist, rank, world):
    '''Create the NCCL communicator of a slab Interface.'''
    from . import _lib
    L = _lib.lib()

    def make_id():
        raw = (ctypes.c_uint8 * 128)()
        _lib.check(L.tfb_nccl_unique_id(raw))
        return bytes(raw)

    uid = broadcast_unique_id(dist, make_id, rank)
    raw = (ctypes.c_uint8 * 128)(*uid)
    _lib.check(L.tfb_comm_init(interface._ctx, world, rank, raw))
