#!/usr/bin/env python3
'''Per-phase device-time breakdown of the z-slab Krylov solve (weak scaling: grid^3 cells per rank).  Launch with
torch.distributed.run, one rank per GPU:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mgpu_solve_probe.py [grid]
Every rank prints the mean device time per operator product by phase (tfb_solve, Verbose).'''
import os
import sys
import time
import warnings

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from transiflow_b200 import Interface, parallel  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
local = int(os.environ.get('LOCAL_RANK', rank))
nz = grid * world
params = {'Reynolds Number': 100, 'Z-max': float(world)}
it = Interface(params, grid, grid, nz, device=local, slab=parallel.slab_range(nz, world, rank))
parallel.init_comm(it, dist, rank, world)
warnings.simplefilter('ignore')
x = it.vector()
for k in range(3):
    if k == 2:
        params['Verbose'] = True
    dist.barrier()
    t0 = time.perf_counter()
    jac, f = it.jacobian_rhs(x)
    dx = it.solve(jac, -f)
    x = x + dx
    ls = it.last_solve
    if rank == 0:
        print('step %d: %s %d products, solve %.1f ms, step %.1f ms' % (k, ls['method'], ls['iterations'], ls['solve_ms'],
                                                                       1e3 * (time.perf_counter() - t0)), flush=True)
dist.barrier()
