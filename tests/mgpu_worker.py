'''Multi-GPU parity worker (one rank per GPU; launched by tests/test_mgpu_gpu.py through torch.distributed.run):
every rank assembles its z-slab (NCCL halo exchange inside the library) and compares its owned rows with the
oracle's result for the whole grid -- the reference's own distributed-vs-serial test pattern
(/root/reference/tests/test_PETSc.py:195-286).  CSR values must be bit-identical to the oracle, distributed Newton
updates within 1e-8 of the pinned SuperLU solve.  Grids have at least two planes per rank and ragged slabs.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_worker.py
'''
import os
import sys
import warnings

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch.distributed as dist  # noqa: E402

from golden_io import compress  # noqa: E402
from oracle.tf_oracle import Oracle, direct_solve  # noqa: E402
from transiflow_b200 import Interface, parallel  # noqa: E402

LDC = {'Reynolds Number': 100}
RB = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
      'X-max': 10, 'Y-max': 10}
TOL = 1e-8


def say(rank, world, msg, good):
    print('rank %d/%d %s -> %s' % (rank, world, msg, 'ok' if good else 'MISMATCH'), flush=True)
    return good


def make(params, nx, ny, nz, rank, world, local):
    k0, k1 = parallel.slab_range(nz, world, rank)
    it = Interface(dict(params), nx, ny, nz, device=local, slab=(k0, k1))
    parallel.init_comm(it, dist, rank, world)
    orc = Oracle(dict(params), nx, ny, nz)
    r0, r1 = parallel.owned_rows(nx, ny, it.dof, k0, k1)
    return it, orc, r0, r1


def assembly_parity(it, orc, r0, r1, rank, world, name):
    state = numpy.random.default_rng(3).uniform(-0.5, 0.5, orc.n)
    jac, f = it.jacobian_rhs(state[r0:r1].copy())
    row_ptr, col = it.pattern()
    gv, gc, gp = compress(jac.values(), col, row_ptr)
    coA, jcoA, begA = orc.jacobian(state)
    e0, e1 = begA[r0], begA[r1]
    parts = {'row_ptr': numpy.array_equal(gp, begA[r0:r1 + 1] - e0), 'cols': numpy.array_equal(gc, jcoA[e0:e1]),
             'values': numpy.array_equal(gv, coA[e0:e1]), 'rhs': numpy.array_equal(f, orc.rhs(state)[r0:r1])}
    good = all(parts.values())
    if not good:
        print('rank', rank, parts, flush=True)
    return say(rank, world, '%s slab rows [%d,%d): owned CSR rows and RHS bit-identical to the oracle' % (name, r0, r1), good)


def solve_parity(it, jac, b_local, want, r0, r1, rank, world, label, opts, required=True):
    it.parameters['Iterative Solver'] = dict(opts)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dx = it.solve(jac, b_local.copy())
    err = numpy.abs(dx - want[r0:r1]).max() / numpy.abs(want).max()
    ls = it.last_solve
    good = err <= TOL and ls['converged']
    say(rank, world, '%s %s: %s/%s/%s %d its, relres %.1e, err vs spsolve %.1e' % (
        label, opts, ls['method'], ls['schur'], ls.get('precond_precision'), ls['iterations'], ls['relres'], err), good or not required)
    it.parameters.pop('Iterative Solver')
    return good or not required


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', rank))
    nz = 2 * world + 3          # >= 2 planes per rank, ragged: the first three ranks own one plane more
    ok = True

    # ---- lid-driven cavity ----
    nx, ny = 12, 9
    it, orc, r0, r1 = make(LDC, nx, ny, nz, rank, world, local)
    ok &= assembly_parity(it, orc, r0, r1, rank, world, 'LDC %dx%dx%d' % (nx, ny, nz))
    x = numpy.zeros(orc.n)
    for step in range(2):       # distributed Newton updates vs the oracle's pinned SuperLU solve of the whole system
        jac, f = it.jacobian_rhs(x[r0:r1].copy())
        want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
        ok &= solve_parity(it, jac, -f, want, r0, r1, rank, world, 'LDC Newton step %d' % step, {})
        x = x + want            # all ranks advance the same global state
    jac, f = it.jacobian_rhs(x[r0:r1].copy())
    want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
    # what 'auto' selects on large grids (IDR(8) + scaled-mass Schur complement + tensor-core FDM sub-solves), forced onto
    # this small one, then the other Krylov / preconditioner variants
    it.AUTO_IDR_MIN_UNKNOWNS = 100
    ok &= solve_parity(it, jac, -f, want, r0, r1, rank, world, 'LDC large-grid defaults', {})
    it.AUTO_IDR_MIN_UNKNOWNS = Interface.AUTO_IDR_MIN_UNKNOWNS
    for opts in ({'Schur Complement': 'Scaled Mass', 'Method': 'IDR', 'Preconditioner Precision': 'double'},
                 {'Schur Complement': 'Scaled Mass', 'Method': 'IDR', 'Preconditioner Precision': 'tf32x3'},
                 {'Schur Complement': 'Scaled Mass', 'Method': 'FGMRES'},
                 {'Preconditioner Precision': 'single'}, {'Method': 'BiCGStab'}, {'Velocity Iterations': 3},
                 {'Basis Precision': 'single'}, {'Method': 'IDR'}, {'Method': 'IDR', 'IDR Dimension': 4}):
        ok &= solve_parity(it, jac, -f, want, r0, r1, rank, world, 'LDC', opts)
    del it, jac

    # ---- taller slabs (>= 4 planes on every rank, ragged): every rank takes the path that exchanges the halo on a side
    # stream next to the interior planes when TFB_OVERLAP is set; RHS-only and Jacobian-only launches as well ----
    nzt = 5 * world + 1
    for prm, (nxt, nyt), name in ((LDC, (34, 5), 'LDC'), (RB, (7, 6), 'Rayleigh-Benard')):
        it, orc, r0, r1 = make(prm, nxt, nyt, nzt, rank, world, local)
        ok &= assembly_parity(it, orc, r0, r1, rank, world, '%s %dx%dx%d' % (name, nxt, nyt, nzt))
        state = numpy.random.default_rng(4).uniform(-0.5, 0.5, orc.n)
        good = numpy.array_equal(it.rhs(state[r0:r1].copy()), orc.rhs(state)[r0:r1])
        jonly = it.jacobian(state[r0:r1].copy())
        both, _ = it.jacobian_rhs(state[r0:r1].copy())
        good = good and numpy.array_equal(jonly.values(), both.values())
        ok &= say(rank, world, '%s tall slabs: rhs() and jacobian() alone equal the fused launch / the oracle' % name, good)
        del it, jonly, both

    # ---- slabs of 17-18 planes: the host path of jacobian_rhs is pipelined over z-pieces (edge planes first, halo exchange
    # while the rest of the slab is uploaded) -- same bits ----
    it, orc, r0, r1 = make(LDC, 7, 5, 17 * world + 1, rank, world, local)
    ok &= assembly_parity(it, orc, r0, r1, rank, world, 'LDC 7x5x%d (pipelined host path)' % (17 * world + 1))
    del it

    # ---- Rayleigh-Benard: assembly, and the coupled (w, T) line solve on pencils (all z, a chunk of y) ----
    nx, ny = 9, 10
    it, orc, r0, r1 = make(RB, nx, ny, nz, rank, world, local)
    ok &= assembly_parity(it, orc, r0, r1, rank, world, 'Rayleigh-Benard %dx%dx%d' % (nx, ny, nz))
    x0 = numpy.zeros(orc.n)
    want0 = direct_solve(orc.jacobian_csr(x0), -orc.rhs(x0), orc.dim, orc.dof)
    jac, f = it.jacobian_rhs(x0[r0:r1].copy())
    ok &= solve_parity(it, jac, -f, want0, r0, r1, rank, world, 'RB Newton step from zero', {})
    x = x0 + want0              # the conduction state
    b = numpy.random.default_rng(5).standard_normal(orc.n)
    b[orc.dim] = 0
    want = direct_solve(orc.jacobian_csr(x), b, orc.dim, orc.dof)
    jac, f = it.jacobian_rhs(x[r0:r1].copy())
    for opts, required in (({}, True), ({'Velocity Iterations': 0}, True),
                           # the block-triangular variant is only reported (it may stall)
                           ({'Scalar Coupling': 'none', 'Maximum Iterations': 2000, 'Restart': 2000}, False)):
        ok &= solve_parity(it, jac, b[r0:r1], want, r0, r1, rank, world, 'RB at the conduction state', opts, required)
    dist.barrier()
    print('rank %d/%d %s' % (rank, world, 'ALL OK' if ok else 'FAILED'), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
