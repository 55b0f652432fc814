#!/usr/bin/env python3
'''Benchmark of the B200 TransiFlow backend's hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 128] [--impl reference]

Workload (BASELINE.json north_star / SURVEY.md section 8d): 3D lid-driven cavity, Re = 100,
grid^3 cells (default 128^3, n = 8.4 M unknowns, nnz = 118 M), synthetic state
uniform(-0.5, 0.5) seed 0.  One "step" = one fused Jacobian+RHS assembly on the fixed
sparsity pattern.  `value` = cells/s with the state resident in HBM; `e2e` = the same through
Interface.jacobian_rhs() with host buffers (H2D state + D2H F(x) inside the timed region).
Further keys: `spmv` (y = J x on the assembled matrix), `newton` (fused assembly + FGMRES solve to
1e-10 per step), `roofline`, `cpu_baseline`, `clocks`.
With N > 1 (torchrun, one process per GPU) the grid is split into z-slabs, weak scaling: every
rank owns `grid` planes of a (grid x grid x N*grid) cavity with Z-max = N (cubic cells); halos
are exchanged with NCCL inside the library.

`--impl reference` times the CPU implementation of the same path (the C port of the
reference in oracle/, all host threads) on a bounded sample of the workload.
'''
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
# secondary workloads (--problem ...): the other BASELINE configurations
RB_PARAMS = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
             'X-max': 10, 'Y-max': 10}
# name: (parameters, nx, ny, description)  -- 2-D configurations run on one GPU at their BASELINE size
PROBLEMS_2D = {
    'ldc2d': ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1, 'Grid Stretching Factor': 1.5},
              32, 32, '2D lid-driven cavity 32x32 (stretched), Re=100'),
    'dhc2d': ({'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e6, 'Prandtl Number': 1000,
               'Reynolds Number': 1, 'X-max': 4.08 / 80, 'Y-max': 1}, 64, 64, '2D differentially heated cavity 64x64, Ra=1e6'),
    'qg': ({'Problem Type': 'Double Gyre', 'Reynolds Number': 16, 'Rossby Parameter': 1000, 'Wind Stress Parameter': 100},
           256, 128, '2D quasi-geostrophic double gyre 256x128'),
    'amoc': ({'Problem Type': 'AMOC', 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1, 'Freshwater Flux': 0,
              'Temperature Forcing': 0.1, 'X-max': 5}, 256, 128, '2D AMOC 256x128'),
}


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    '''nvidia-smi clocks / throttle reasons during the timed region.'''
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([s.strip() for s in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        reasons = []
        for idx, name in ((2, 'hw_slowdown'), (3, 'hw_thermal_slowdown'), (4, 'sw_thermal_slowdown'), (5, 'sw_power_cap')):
            if any(len(s) > idx and s[idx].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def cpu_port_cells_per_s(grid, planes, repeats=1):
    '''Jacobian+RHS with the oracle C port (all OpenMP threads) on a (grid x grid x planes) sample.'''
    from oracle.tf_oracle import Oracle, lib
    # all host cores, set explicitly: launchers such as torch.distributed.run export OMP_NUM_THREADS=1
    lib().tfo_set_num_threads(os.cpu_count() or 1)
    orc = Oracle(dict(PARAMS), grid, grid, planes)
    state = numpy.random.default_rng(0).uniform(-0.5, 0.5, orc.n)
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        orc.jacobian(state)
        orc.rhs(state)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return grid * grid * planes / best, lib().tfo_num_threads(), best


def python_reference_sample(n=16, grid2d=None):
    '''The UNMODIFIED Python reference (SciPy backend, pip-installed under baseline/_ref) on this box's host:
    Interface.rhs + Interface.jacobian (single-threaded Python) and Interface.solve (SuperLU) -- on a small n^3 instance
    of the 3-D workload (about 10 s), or on the full (nx, ny) grid of a 2-D configuration.  None when the package is
    not there.'''
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'transiflow')):
        return None
    sys.path.insert(0, ref)
    try:
        from transiflow import Interface as RefInterface
        params = {k: v for k, v in PARAMS.items() if k != 'Iterative Solver'}
        if grid2d:
            it = RefInterface(dict(params), grid2d[0], grid2d[1])
            x = numpy.random.default_rng(0).uniform(-0.01, 0.01, it.vector().size)
            t0 = time.perf_counter()
            f = it.rhs(x)
            t1 = time.perf_counter()
            jac = it.jacobian(x)
            t2 = time.perf_counter()
            it.solve(jac, -f)
            t3 = time.perf_counter()
            cells = grid2d[0] * grid2d[1]
            return {'grid': [grid2d[0], grid2d[1], 1], 'rhs_s': t1 - t0, 'jacobian_s': t2 - t1, 'solve_s': t3 - t2,
                    'assembly_cells_per_s': cells / (t2 - t0), 'newton_steps_per_s': 1.0 / (t3 - t0), 'cores': 1,
                    'note': 'unmodified reference, SciPy backend, the same grid and parameters'}
        it = RefInterface(dict(params), n, n, n)
        x = numpy.random.default_rng(0).uniform(-0.1, 0.1, n * n * n * 4)
        t0 = time.perf_counter()
        f = it.rhs(x)
        t1 = time.perf_counter()
        jac = it.jacobian(x)
        t2 = time.perf_counter()
        it.solve(jac, -f)
        t3 = time.perf_counter()
        return {'grid': [n, n, n], 'rhs_s': t1 - t0, 'jacobian_s': t2 - t1, 'solve_s': t3 - t2,
                'assembly_cells_per_s': n ** 3 / (t2 - t0), 'newton_steps_per_s': 1.0 / (t3 - t0), 'cores': 1,
                'note': 'unmodified reference, SciPy backend (Python assembly on one core, SuperLU solve); its cost per cell is '
                        'size-independent for the assembly, the direct solve is infeasible beyond ~64^3'}
    except Exception as e:     # noqa: BLE001
        return {'error': str(e)}
    finally:
        sys.path.remove(ref)


def parity_gate(device, flat):
    '''Correctness gate reported with every perf number: a small instance of the same
    problem against the CPU oracle -- pattern after compress, CSR values, RHS, Newton update.'''
    from oracle.tf_oracle import Oracle, direct_solve
    from transiflow_b200 import Interface
    n = 12
    params = {k: v for k, v in PARAMS.items() if k != 'Iterative Solver'}
    it = Interface(dict(params), n, n, 1 if flat else n, device=device)
    # the solver configuration 'auto' selects for the timed 128^3 Newton steps (IDR(8) + scaled-mass Schur complement +
    # tensor-core FDM sub-solves + CUDA-graph replay), forced onto this small instance
    it.AUTO_IDR_MIN_UNKNOWNS = 1000
    orc = Oracle(dict(params), it.nx, it.ny, it.nz)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, f = it.jacobian_rhs(x)
    csr = jac.tocsr()
    coA, jcoA, begA = orc.jacobian(x)
    fo = orc.rhs(x)
    out = {'checked_on': '%dx%dx%d instance vs CPU oracle' % (it.nx, it.ny, it.nz),
           'pattern_exact': bool(numpy.array_equal(csr.indptr, begA) and numpy.array_equal(csr.indices, jcoA)),
           'values_max_rel_err': float(numpy.max(numpy.abs(csr.data - coA) / numpy.abs(coA))) if len(coA) == csr.nnz else None,
           'rhs_max_abs_err': float(numpy.abs(f - fo).max())}
    try:
        dx = it.solve(jac, -f)
        want = direct_solve(orc.jacobian_csr(x), -fo, orc.dim, orc.dof)
        out['newton_update_max_rel_err_vs_spsolve'] = float(numpy.abs(dx - want).max() / numpy.abs(want).max())
        out['newton_update_solver'] = '%s / %s / preconditioner %s, %d iterations, relres %.1e' % (
            it.last_solve['method'], it.last_solve['schur'], it.last_solve['precond_precision'],
            it.last_solve['iterations'], it.last_solve['relres'])
    except Exception as e:     # noqa: BLE001
        out['newton_update_error'] = str(e)
    return out


def parity_gate_slabs(dist, rank, world, device):
    '''The same gate for a z-slab run: every rank's owned CSR rows and RHS against the oracle (bit-identical) and the
    distributed Newton update against the pinned SuperLU solve, on a ragged grid with >= 2 planes per rank.  Collective:
    every rank calls it; the worst rank is reported.'''
    import torch
    from oracle.tf_oracle import Oracle, direct_solve
    from transiflow_b200 import Interface, parallel
    nx, ny, nz = 12, 10, 2 * world + 3
    params = {k: v for k, v in PARAMS.items() if k != 'Iterative Solver'}
    k0, k1 = parallel.slab_range(nz, world, rank)
    it = Interface(dict(params), nx, ny, nz, device=device, slab=(k0, k1))
    parallel.init_comm(it, dist, rank, world)
    it.AUTO_IDR_MIN_UNKNOWNS = 1000
    orc = Oracle(dict(params), nx, ny, nz)
    r0, r1 = parallel.owned_rows(nx, ny, it.dof, k0, k1)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, orc.n)
    jac, f = it.jacobian_rhs(x[r0:r1].copy())
    row_ptr, col = it.pattern()
    vals = jac.values()
    keep = numpy.abs(vals) > 1e-14
    csum = numpy.concatenate(([0], numpy.cumsum(keep, dtype=numpy.int64)))
    coA, jcoA, begA = orc.jacobian(x)
    e0, e1 = begA[r0], begA[r1]
    exact = bool(numpy.array_equal(csum[row_ptr], begA[r0:r1 + 1] - e0) and numpy.array_equal(col[keep], jcoA[e0:e1])
                 and numpy.array_equal(vals[keep], coA[e0:e1]) and numpy.array_equal(f, orc.rhs(x)[r0:r1]))
    err, solver = float('inf'), ''
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dx = it.solve(jac, -f)
        want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
        err = float(numpy.abs(dx - want[r0:r1]).max() / numpy.abs(want).max())
        solver = '%s / %s / preconditioner %s, %d iterations' % (it.last_solve['method'], it.last_solve['schur'],
                                                                 it.last_solve['precond_precision'], it.last_solve['iterations'])
    except Exception as e:     # noqa: BLE001
        solver = 'error: ' + str(e)
    t = torch.tensor([0.0 if exact else 1.0, err if numpy.isfinite(err) else 1e300], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {'checked_on': '%dx%dx%d cavity on %d z-slabs vs CPU oracle (every rank, worst reported)' % (nx, ny, nz, world),
            'owned_rows_and_rhs_bit_identical': bool(t[0] == 0.0),
            'newton_update_max_rel_err_vs_spsolve': float(t[1]), 'newton_update_solver': solver}


def solve_roofline(it, step, peak_gbs):
    '''Algorithmic bytes of one operator product of the Krylov solve that was timed over the
    measured time per product.  Components are listed so the figure can be recomputed.'''
    n, nnz, dof, dim = it.n_local, it.nnz, it.dof, it.dim
    ncell = n // dof
    ms_per_product = step['solve_ms'] / max(1, step['iterations'])
    comp = {'operator': 8 * nnz + 4 * (n + 1) + 16 * n}         # CSR values + row pointers + x read + y written
    if step.get('method') == 'IDR':
        s = 8
        # vectors of 8n bytes moved per product by IDR(s) in the bi-orthogonal form, averaged over k = 0..s-1 and the
        # omega step: v = r - G c (s-k+2), U_k (s-k+2), P^T G_k (1), fused bi-orthogonalisation + update sweeps (2(k+4))
        # (shadow vectors are generated on the fly: P^T G_k reads G_k only; the bi-orthogonalisation of G_k / U_k is fused
        # with the update of r / x: k + 4 vectors per family)
        per_k = [(s - k + 2) + (s - k + 2) + 1 + 2 * (k + 4) for k in range(s)]
        comp['recurrence_vectors'] = 8 * n * (sum(per_k) + 10) // (s + 1)
    else:
        its = max(1, step['iterations'])
        comp['orthogonalisation_vectors'] = 8 * n * (its + 6)       # mean basis size its/2, read twice, plus w, z
    if step.get('precond') == 'tf32x3':
        # r read, z written (fp64); per cell: two-slot gradient block (2 floats per velocity row), fp32 planes of the `dim`
        # velocity components written by the head, x/y forward (r+w), Thomas (r+w + two factor arrays), x/y backward (r+w),
        # read by the tail, pressure update (w+r)
        comp['preconditioner'] = 16 * n + ncell * (8 * dim + 4 * dim + 8 * dim + 16 * dim + 8 * dim + 4 * dim + 8)
    else:
        comp['preconditioner'] = 16 * n + dim * ncell * 8 * 14       # six fp64 transforms (r+w) + scaling + (de)interleave
    total = sum(comp.values())
    achieved = total / (ms_per_product * 1e-3) / 1e9
    return {'bound': 'hbm', 'unit': 'GB/s', 'achieved': achieved, 'peak': peak_gbs, 'frac': achieved / peak_gbs,
            'ms_per_product': ms_per_product, 'products': step['iterations'], 'algorithmic_bytes_per_product': total,
            'components': comp}


def rb_strong_leg(args, dist, rank, world, local_rank, max_over_ranks):
    '''BASELINE config 4: ONE Rayleigh-Benard grid^3 problem (dof 5, Ra = 1000) split over `world` z-slabs -- fused
    assembly time, one Newton step from the perturbed conduction state, iterations.  Collective.'''
    import ctypes
    from transiflow_b200 import DeviceMatrix, Interface, _lib, parallel
    from transiflow_b200._lib import check, ptr
    L = _lib.lib()
    grid = args.grid
    if world > 1:
        slab = parallel.slab_range(grid, world, rank)
        it = Interface(dict(RB_PARAMS), grid, grid, grid, device=local_rank, slab=slab)
        parallel.init_comm(it, dist, rank, world)
    else:
        it = Interface(dict(RB_PARAMS), grid, grid, grid, device=local_rank)
    it._sync_params()
    state = _lib.pinned_array(it.n_local)
    state[:] = numpy.random.default_rng(rank).uniform(-0.5, 0.5, it.n_local)
    mat = DeviceMatrix(it)
    check(L.tfb_state_upload(it._ctx, ptr(state)))
    for _ in range(5):
        check(L.tfb_flush_l2(it._ctx))
        check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
    check(L.tfb_sync(it._ctx))
    if dist:
        dist.barrier()
    ms_list = []
    for _ in range(20):
        check(L.tfb_flush_l2(it._ctx))
        check(L.tfb_event_record(it._ctx, 6))
        check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
        check(L.tfb_event_record(it._ctx, 7))
        ms = ctypes.c_float()
        check(L.tfb_event_elapsed_ms(it._ctx, 6, 7, ctypes.byref(ms)))
        ms_list.append(ms.value)
    asm_ms = max_over_ranks(sum(ms_list) / len(ms_list))
    del mat
    # Newton: the first step from zero lands on the conduction state (linear); the timed steps start from that state
    # plus a smooth roll-like perturbation
    import warnings
    x = it.vector()
    steps = []
    for k in range(3):
        t0 = time.perf_counter()
        check(L.tfb_event_record(it._ctx, 6))
        jac, f = it.jacobian_rhs(x)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            dx = it.solve(jac, -f)
        check(L.tfb_event_record(it._ctx, 7))
        ms = ctypes.c_float()
        check(L.tfb_event_elapsed_ms(it._ctx, 6, 7, ctypes.byref(ms)))
        x = x + dx
        if k == 0:
            k0s, k1s = it.slab
            c3 = numpy.indices((k1s - k0s, it.ny, it.nx)).astype(float)
            roll = numpy.sin(numpy.pi * (c3[0] + k0s + 1) / it.nz) * numpy.cos(6 * numpy.pi * (c3[2] + 0.5) / it.nx)
            xs = x.reshape(k1s - k0s, it.ny, it.nx, it.dof)
            xs[..., 2] += 1e-2 * roll
            if k1s == it.nz:
                xs[-1, :, :, 2] = 0.0
            xs[..., 4] += 1e-2 * roll
        steps.append({'ms': max_over_ranks(ms.value), 'iterations': it.last_solve['iterations'], 'relres': it.last_solve['relres'],
                      'solve_ms': max_over_ranks(it.last_solve['solve_ms']), 'converged': bool(it.last_solve['converged']),
                      'method': it.last_solve['method'], 'precond': it.last_solve.get('precond_precision')})
    timed = steps[1:]
    alg = 8 * it.nnz + 16 * it.n_local
    return {'workload': 'ONE 3D Rayleigh-Benard %d^3 problem (dof 5, Ra=1000) on %d z-slab(s)' % (grid, world), 'scaling': 'strong',
            'assembly_ms': asm_ms, 'assembly_cells_per_s': grid ** 3 / (asm_ms * 1e-3),
            'assembly_roofline_frac': alg / (asm_ms * 1e-3) / 1e9 / measured_peaks()[0],
            'newton_ms_per_step': sum(h['ms'] for h in timed) / len(timed),
            'newton_steps_per_s': 1e3 * len(timed) / sum(h['ms'] for h in timed),
            'krylov_iterations': [h['iterations'] for h in timed], 'relres': [h['relres'] for h in timed],
            'solve_ms': [round(h['solve_ms'], 1) for h in timed], 'all_converged': all(h['converged'] for h in steps),
            'method': timed[-1]['method'], 'preconditioner_precision': timed[-1]['precond'], 'unknowns': it.n}


def run_reference(args, rank, world):
    if rank != 0:
        return
    grid = args.grid
    planes = max(2, min(grid, 16))
    sample = '3D LDC %dx%dx%d slab of the %d^3 grid (Jacobian+RHS, C port of the reference, OpenMP)' % (grid, grid, planes, grid)
    for _ in range(args.warmup):
        cpu_port_cells_per_s(grid, planes)
    times = []
    cores = 1
    for _ in range(args.steps):
        v, cores, dt = cpu_port_cells_per_s(grid, planes)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = grid * grid * planes / (ms * 1e-3)
    out = {
        'impl': 'reference', 'metric': 'jacobian_rhs_assembly_cells_per_s', 'value': value, 'unit': 'cells/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '3D lid-driven cavity %d^3, Re=100, fused Jacobian+RHS assembly' % grid},
        'cpu_baseline': {'value': value, 'unit': 'cells/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cells/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    pyref = python_reference_sample()
    if pyref is not None:
        out['cpu_baseline']['python_reference'] = pyref
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--newton-steps', type=int, default=3, help='timed Newton steps (0 = skip the solver leg)')
    ap.add_argument('--grid', type=int, default=128)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--problem', default='ldc', choices=['ldc', 'rb'] + sorted(PROBLEMS_2D))
    ap.add_argument('--rb-strong', type=int, default=1,
                    help='1 (default): add the strong-scaling leg of one Rayleigh-Benard grid^3 problem on N z-slabs; 0: skip')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak (default): grid^3 cells per GPU; strong: one grid^3 problem split over the GPUs')
    args = ap.parse_args()
    if args.problem == 'rb':
        PARAMS.clear()
        PARAMS.update(RB_PARAMS)
    two_d = PROBLEMS_2D.get(args.problem)
    if two_d:
        PARAMS.clear()
        PARAMS.update(two_d[0])
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import ctypes
    from transiflow_b200 import Interface, _lib
    from transiflow_b200._lib import check, ptr
    L = _lib.lib()

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('gloo')     # plumbing only: id broadcast, barrier, max-reduce

    grid = args.grid
    # weak scaling: every rank owns `grid` planes of a (grid x grid x world*grid) domain
    nz = grid * world if args.scaling == 'weak' else grid
    params = dict(PARAMS)
    if world > 1 and args.scaling == 'weak' and args.problem != 'rb':
        params['Z-max'] = float(world)      # the cavity grows with the GPU count: cells stay cubic
    # (Rayleigh-Benard keeps its unit layer height -- a taller layer is a different, far more supercritical
    # problem: the effective Rayleigh number grows with height^3 -- and refines z instead)
    if two_d:
        if world > 1:
            raise SystemExit('2-D configurations are single-GPU ("replicas only", DESIGN.md section 5)')
        it = Interface(params, two_d[1], two_d[2], 1, device=local_rank)
    elif world > 1:
        from transiflow_b200 import parallel
        it = Interface(params, grid, grid, nz, device=local_rank, slab=parallel.slab_range(nz, world, rank))
    else:
        it = Interface(params, grid, grid, grid, device=local_rank)
    if world > 1:
        from transiflow_b200 import parallel
        parallel.init_comm(it, dist, rank, world)

    n_local = it.n_local
    cells_local = n_local // it.dof
    state = _lib.pinned_array(n_local)
    state[:] = numpy.random.default_rng(rank).uniform(-0.5, 0.5, n_local)
    out = _lib.pinned_array(n_local)
    it._sync_params()
    from transiflow_b200 import DeviceMatrix
    mat = DeviceMatrix(it)

    def barrier():
        check(L.tfb_sync(it._ctx))
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    launches0 = L.tfb_launch_count()
    # ---- kernel-only: state resident in HBM; L2 flushed between iterations ----
    check(L.tfb_state_upload(it._ctx, ptr(state)))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_warm = time.perf_counter()
    nwarm = 0
    while nwarm < args.warmup or time.perf_counter() - t_warm < 0.8:   # nvidia-smi needs ~0.5 s to start sampling
        check(L.tfb_flush_l2(it._ctx))
        check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
        nwarm += 1
        if nwarm % 50 == 0:
            check(L.tfb_sync(it._ctx))
    barrier()
    # one (start, stop) event pair per iteration, read after the loop: a host synchronisation inside the loop would put
    # the launch jitter of the slowest rank into every z-slab iteration (each one meets its neighbours in the halo exchange)
    kern_ms = []
    t_wall0 = time.perf_counter()
    for s0 in range(0, args.steps, 512):
        block = range(s0, min(s0 + 512, args.steps))
        for s in block:
            check(L.tfb_flush_l2(it._ctx))
            check(L.tfb_event_record(it._ctx, 16 + 2 * (s - s0)))
            check(L.tfb_assemble_resident(it._ctx, mat._h, 1, 1))
            check(L.tfb_event_record(it._ctx, 17 + 2 * (s - s0)))
        for s in block:
            ms = ctypes.c_float()
            check(L.tfb_event_elapsed_ms(it._ctx, 16 + 2 * (s - s0), 17 + 2 * (s - s0), ctypes.byref(ms)))
            kern_ms.append(ms.value)
    barrier()
    launches = L.tfb_launch_count() - launches0
    step_ms = max_over_ranks(sum(kern_ms) / len(kern_ms))
    # ---- e2e: host buffers through the public API, copies inside the timed region ----
    for _ in range(3):
        it.jacobian_rhs_into(state, mat, out)
    barrier()
    t0 = time.perf_counter()
    check(L.tfb_event_record(it._ctx, 2))
    for s in range(args.steps):
        it.jacobian_rhs_into(state, mat, out)
    check(L.tfb_event_record(it._ctx, 3))
    ms = ctypes.c_float()
    check(L.tfb_event_elapsed_ms(it._ctx, 2, 3, ctypes.byref(ms)))
    barrier()
    e2e_ms = max_over_ranks(ms.value / args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- SpMV y = J x on the assembled matrix (resident operands; 1.6 GB of operands >> L2) ----
    spmv = None
    try:
        ms_sp = ctypes.c_float()
        check(L.tfb_spmv_bench(mat._h, 50, 0, ctypes.byref(ms_sp)))
        sp_ms = max_over_ranks(ms_sp.value)
        csr_bytes = 12 * it.nnz + 4 * (n_local + 1) + 16 * n_local    # CSR algorithmic bytes
        own_bytes = 8 * it.nnz + 4 * (n_local + 1) + 16 * n_local     # what this kernel moves: it never reads column indices
        spmv = {'ms': sp_ms, 'bytes_moved': own_bytes, 'achieved_gbs': own_bytes / (sp_ms * 1e-3) / 1e9,
                'frac_of_hbm_peak': own_bytes / (sp_ms * 1e-3) / 1e9 / measured_peaks()[0],
                'csr_equivalent_bytes': csr_bytes, 'csr_equivalent_gbs': csr_bytes / (sp_ms * 1e-3) / 1e9,
                'kernel': 'tfb_spmv_march_kernel (values-only stream, TMA bulk loads, no column indices)'}
        if world > 1:
            # every product of this loop meets both z-neighbours in its halo exchange: the figure is the median of 50
            # device-timed products (a late rank lengthens single samples by milliseconds)
            spmv['includes'] = 'halo exchange (NCCL send/recv with both z-neighbours); median of 50 device-timed products'
    except Exception as e:     # noqa: BLE001
        spmv = {'error': str(e)}

    # ---- Newton step: fused assembly + preconditioned FGMRES to 1e-10 (single GPU) ----
    newton = None
    if args.newton_steps > 0:
        x = it.vector()
        hist = []
        for k in range(2 + args.newton_steps):          # 2 untimed steps bring the state into the convective regime
            t0 = time.perf_counter()
            check(L.tfb_event_record(it._ctx, 4))
            jac, f = it.jacobian_rhs(x)
            dx = it.solve(jac, -f)
            check(L.tfb_event_record(it._ctx, 5))
            ms5 = ctypes.c_float()
            check(L.tfb_event_elapsed_ms(it._ctx, 4, 5, ctypes.byref(ms5)))
            x = x + dx
            if args.problem == 'rb' and k == 0:
                # Newton from zero lands on the conduction state in one (linear) step; the timed steps start from
                # that state plus a smooth roll-like perturbation so that they are genuine Newton steps
                k0s, k1s = it.slab
                c3 = numpy.indices((k1s - k0s, it.ny, it.nx)).astype(float)
                roll = numpy.sin(numpy.pi * (c3[0] + k0s + 1) / it.nz) * numpy.cos(6 * numpy.pi * (c3[2] + 0.5) / it.nx)
                xs = x.reshape(k1s - k0s, it.ny, it.nx, it.dof)
                xs[..., 2] += 1e-2 * roll
                if k1s == it.nz:
                    xs[-1, :, :, 2] = 0.0
                xs[..., 4] += 1e-2 * roll
            hist.append({'ms': max_over_ranks(ms5.value), 'wall_ms': 1e3 * (time.perf_counter() - t0),
                         'fnorm': float(numpy.sqrt(max_over_ranks(float(f @ f)) if world > 1 else f @ f)),
                         'iterations': it.last_solve['iterations'], 'relres': it.last_solve['relres'],
                         'solve_ms': max_over_ranks(it.last_solve['solve_ms']),
                         'precond': it.last_solve.get('precond_precision'), 'method': it.last_solve.get('method'),
                         'schur': it.last_solve.get('schur'),
                         'converged': bool(it.last_solve['converged'])})
        timed = hist[2:]
        nms = sum(h['ms'] for h in timed) / len(timed)
        newton = {'steps_per_s': 1e3 / nms, 'ms_per_step': nms, 'timed_steps': len(timed),
                  'ms_of_each_step': [round(h['ms'], 1) for h in timed], 'solve_ms_of_each_step': [round(h['solve_ms'], 1) for h in timed],
                  'krylov_iterations': [h['iterations'] for h in timed], 'relres': [h['relres'] for h in timed],
                  'fnorm_before': [h['fnorm'] for h in timed], 'all_converged': all(h['converged'] for h in hist),
                  'tolerance': 1e-10, 'unknowns': it.n,
                  'krylov_method': it.last_solve.get('method'),
                  'solver': ('direct: block-tridiagonal elimination over the grid lines, pivoted dense line inverses (csrc/tfb_direct.cu)'
                             if it.last_solve.get('method') == 'Direct' else
                             'FGMRES + LSC block preconditioner, coupled (w,T) line solve + inner GMRES on the (u,T) block'
                             if args.problem == 'rb' else
                             ('IDR(8)' if it.last_solve.get('method') == 'IDR' else 'FGMRES')
                             + (' + block preconditioner (FDM velocity solves, scaled-mass Schur complement)'
                                if it.last_solve.get('schur') == 'Scaled Mass' else ' + LSC block preconditioner (FDM sub-solves)')
                             + ('; FDM x/y transforms on tcgen05 (3xTF32), Thomas sweeps along z'
                                if hist[-1].get('precond') == 'tf32x3' else ''))
                            + ', host vectors in/out'
                            + ('; z-slabs: NCCL halo exchange, all-reduce, all-to-all transposes' if world > 1 else '')}
        # per-product roofline of the solve: algorithmic bytes of one IDR(s) operator product over
        # the measured time per product of the last timed solve
        try:
            if not two_d:
                newton['roofline'] = solve_roofline(it, timed[-1], measured_peaks()[0])
        except Exception as e:     # noqa: BLE001
            newton['roofline'] = {'error': str(e)}
        # the same linear system with the fp64 SIMT preconditioner (the tensor-core path only steers the iteration: both
        # reach the same 1e-10 true residual) -- shows what the tcgen05 transforms buy
        try:
            if two_d:
                raise RuntimeError('not applicable to the direct solve of 2-D grids')
            saved = it.parameters.get('Iterative Solver')
            it.parameters['Iterative Solver'] = dict(saved or {}, **{'Preconditioner Precision': 'double'})
            it.solve(jac, -f)
            newton['fp64_preconditioner_solve'] = {'solve_ms': max_over_ranks(it.last_solve['solve_ms']),
                                                   'iterations': it.last_solve['iterations'],
                                                   'relres': it.last_solve['relres'],
                                                   'converged': bool(it.last_solve['converged']),
                                                   'default_solve_ms': timed[-1]['solve_ms'],
                                                   'default_preconditioner': timed[-1].get('precond')}
            if saved is None:
                it.parameters.pop('Iterative Solver')
            else:
                it.parameters['Iterative Solver'] = saved
        except Exception as e:     # noqa: BLE001
            newton['fp64_preconditioner_solve'] = {'error': str(e)}

    # ---- strong scaling of ONE Rayleigh-Benard grid^3 problem on `world` z-slabs (BASELINE config 4) ----
    rb_strong = None
    if args.rb_strong and not two_d:
        try:
            rb_strong = rb_strong_leg(args, dist, rank, world, local_rank, max_over_ranks)
        except Exception as e:     # noqa: BLE001
            rb_strong = {'error': str(e)}
    slab_gate = parity_gate_slabs(dist, rank, world, local_rank) if world > 1 else None

    if rank != 0:
        return
    total_cells = (cells_local * world) if (args.scaling == 'weak' or two_d) else grid ** 3
    value = total_cells / (step_ms * 1e-3)
    nnz = it.nnz
    alg_bytes = 8 * nnz + 16 * n_local       # CSR values written + state read + RHS written
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (step_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the timed kernel, from the committed `ncu --set full`
    # capture of the CURRENT kernel on this workload: profiles/traffic.json, written by profiles/summarize_ncu.py
    ncu_traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as tf:
            entry = json.load(tf).get('%s_%d_%dgpu' % (args.problem, grid, world))
        if entry:
            ncu_traffic, traffic_src = entry['dram_bytes_per_launch'], entry['source']
    except (OSError, ValueError, KeyError):
        pass
    line = {
        'metric': 'jacobian_rhs_assembly_cells_per_s', 'value': value, 'unit': 'cells/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_ms,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': (two_d[3] + ', fused Jacobian+RHS assembly (launch-latency bound: absolute numbers only)') if two_d else
                               ('3D lid-driven cavity %d^3 per GPU, Re=100, fused Jacobian+RHS assembly' if args.problem == 'ldc'
                                else '3D Rayleigh-Benard %d^3 per GPU (dof 5), Ra=1000, fused Jacobian+RHS assembly') % grid,
                   'grid': [it.nx, it.ny, it.nz], 'unknowns': it.n, 'nnz_per_gpu': nnz,
                   'partition': 'z-slabs' if world > 1 else 'single GPU', 'l2': 'flushed between timed iterations (256 MiB write)'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': ncu_traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                     'kernel': ('tfb_assemble_kernel (2-D tile kernel)' if two_d else
                                'tfb_assemble_march_kernel<Cfg_%s,J=1,F=1>' % ('rb3d' if args.problem == 'rb' else 'ldc3d')),
                     'algorithmic_bytes_per_launch': alg_bytes},
        'e2e': {'value': total_cells / (e2e_ms * 1e-3), 'unit': 'cells/s', 'ms_per_step': e2e_ms,
                'h2d_bytes_per_step': 8 * n_local, 'd2h_bytes_per_step': 8 * n_local},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'newton': newton,
        'spmv': spmv,
        'parity': parity_gate(local_rank, bool(two_d)),
    }
    if slab_gate is not None:
        line['parity_slabs'] = slab_gate
    if rb_strong is not None:
        line['rb_strong'] = rb_strong
    if not args.no_cpu_baseline and world == 1 and two_d:
        pyref = python_reference_sample(grid2d=(two_d[1], two_d[2]))
        line['cpu_baseline'] = {'value': pyref['assembly_cells_per_s'] if pyref and 'error' not in pyref else None, 'unit': 'cells/s',
                                'cores': 1, 'kind': 'reference', 'sample': 'the full %dx%d grid, unmodified reference (SciPy backend)' % (two_d[1], two_d[2]),
                                'python_reference': pyref}
    if not args.no_cpu_baseline and world == 1 and not two_d:
        planes = max(2, min(grid, 32))
        v, cores, dt = cpu_port_cells_per_s(grid, planes, repeats=2)
        line['cpu_baseline'] = {'value': v, 'unit': 'cells/s', 'cores': cores, 'kind': 'port',
                                'sample': '%dx%dx%d slab of the workload, Jacobian+RHS, oracle C port with OpenMP' % (grid, grid, planes)}
        pyref = python_reference_sample()
        if pyref is not None:
            line['cpu_baseline']['python_reference'] = pyref
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
