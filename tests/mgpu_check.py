'''Multi-GPU parity check (run under torchrun on the GPU box, one rank per GPU):
every rank assembles its z-slab (NCCL halo exchange inside the library) and compares its owned
rows with the oracle's result for the whole grid -- the reference's own distributed-vs-serial
test pattern (tests/test_PETSc.py:195-286).  CSR values must be bit-identical to 1 GPU.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py
'''
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch.distributed as dist  # noqa: E402

from golden_io import compress  # noqa: E402
from oracle.tf_oracle import Oracle  # noqa: E402
from transiflow_b200 import Interface, parallel  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get('LOCAL_RANK', rank))
    ok = True
    for params, nx, ny, nz in (({'Reynolds Number': 100}, 12, 9, 11),
                               ({'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 2000.0, 'Prandtl Number': 10.0,
                                 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}, 9, 10, 13)):
        k0, k1 = parallel.slab_range(nz, world, rank)
        it = Interface(dict(params), nx, ny, nz, device=local, slab=(k0, k1))
        parallel.init_comm(it, dist, rank, world)
        orc = Oracle(dict(params), nx, ny, nz)
        state = numpy.random.default_rng(3).uniform(-0.5, 0.5, orc.n)
        r0, r1 = parallel.owned_rows(nx, ny, it.dof, k0, k1)
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
        print('rank %d/%d %s slab [%d,%d): %s' % (rank, world, params.get('Problem Type', 'LDC'), k0, k1,
                                                 'bit-identical to the oracle' if good else 'MISMATCH'), flush=True)
        ok = ok and good
        if params.get('Problem Type') is None:
            # distributed Newton update vs the oracle's pinned SuperLU solve of the whole system
            from oracle.tf_oracle import direct_solve
            x = numpy.zeros(orc.n)
            for step in range(2):
                jac, f = it.jacobian_rhs(x[r0:r1].copy())
                dx = it.solve(jac, -f)
                want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
                err = numpy.abs(dx - want[r0:r1]).max() / numpy.abs(want).max()
                good = err <= 1e-8 and it.last_solve['converged']
                print('rank %d/%d distributed solve step %d: %d its, relres %.1e, err vs spsolve %.1e -> %s' % (
                    rank, world, step, it.last_solve['iterations'], it.last_solve['relres'], err, 'ok' if good else 'MISMATCH'), flush=True)
                ok = ok and good
                # all ranks advance the same global state (gather through the oracle's full solution)
                x = x + want
            # the optional Krylov / preconditioner variants must reach the same update on z-slabs
            # (fp32 FDM sub-solves transpose fp32 pencils through the all-to-all)
            jac, f = it.jacobian_rhs(x[r0:r1].copy())
            want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
            for opts in ({'Preconditioner Precision': 'single'}, {'Method': 'BiCGStab'}, {'Velocity Iterations': 3},
                         {'Basis Precision': 'single'}, {'Method': 'IDR'}, {'Method': 'IDR', 'IDR Dimension': 4},
                         {'Schur Complement': 'Scaled Mass', 'Method': 'FGMRES'},
                         {'Schur Complement': 'Scaled Mass', 'Method': 'IDR'}):     # what 'auto' selects on large grids
                it.parameters['Iterative Solver'] = dict(opts)
                dx = it.solve(jac, -f)
                err = numpy.abs(dx - want[r0:r1]).max() / numpy.abs(want).max()
                good = err <= 1e-8 and it.last_solve['converged']
                print('rank %d/%d distributed solve %s: %d its, relres %.1e, err %.1e -> %s' % (
                    rank, world, opts, it.last_solve['iterations'], it.last_solve['relres'], err, 'ok' if good else 'MISMATCH'), flush=True)
                ok = ok and good
            it.parameters.pop('Iterative Solver')
    # Rayleigh-Benard: the coupled (w, T) line solve runs on pencils (all z, a chunk of y) reached through two more
    # all-to-alls; the distributed update at the conduction state must equal the pinned SuperLU solve
    from oracle.tf_oracle import direct_solve
    params = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
              'X-max': 10, 'Y-max': 10}
    nx, ny, nz = 12, 10, 9
    k0, k1 = parallel.slab_range(nz, world, rank)
    it = Interface(dict(params), nx, ny, nz, device=local, slab=(k0, k1))
    parallel.init_comm(it, dist, rank, world)
    orc = Oracle(dict(params), nx, ny, nz)
    r0, r1 = parallel.owned_rows(nx, ny, it.dof, k0, k1)
    x0 = numpy.zeros(orc.n)
    x = x0 + direct_solve(orc.jacobian_csr(x0), -orc.rhs(x0), orc.dim, orc.dof)
    b = numpy.random.default_rng(5).standard_normal(orc.n)
    b[3] = 0
    want = direct_solve(orc.jacobian_csr(x), b, orc.dim, orc.dof)
    jac, f = it.jacobian_rhs(x[r0:r1].copy())
    for opts in ({}, {'Velocity Iterations': 0}, {'Scalar Coupling': 'none', 'Maximum Iterations': 2000, 'Restart': 2000}):
        it.parameters['Iterative Solver'] = dict(opts)
        dx = it.solve(jac, b[r0:r1].copy())
        err = numpy.abs(dx - want[r0:r1]).max() / numpy.abs(want).max()
        good = err <= 1e-8 and it.last_solve['converged']
        print('rank %d/%d Rayleigh-Benard distributed solve %s: %d its, relres %.1e, err vs spsolve %.1e -> %s' % (
            rank, world, opts, it.last_solve['iterations'], it.last_solve['relres'], err, 'ok' if good else 'MISMATCH'), flush=True)
        if 'Scalar Coupling' not in opts:      # the block-triangular variant is only reported (it may stall)
            ok = ok and good
    dist.barrier()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
