'''Host-side preconditioner set-up (hostprep.fdm_operators): the fast-diagonalisation data must
invert the diffusion blocks of the oracle's Jacobian exactly (wall folds, stretched grids, Robin
and Dirichlet scalars included).  CPU-only: numpy applies the same transforms the device does.'''
import numpy
import pytest
from scipy.sparse.linalg import splu

from oracle.tf_oracle import Oracle
from transiflow_b200 import hostprep, recipes


def fdm_apply(ops, v, nx, ny, nz, r3):
    '''Op_v^-1 r on a (nz, ny, nx) array -- numpy mirror of fdm_solve() in csrc/tfb_solver.cu.'''
    parts = sorted([o for o in ops if o[0] == v], key=lambda o: o[1])
    coef = parts[0][5]
    Q = [p[3] for p in parts]
    lam = [p[4] for p in parts]
    m = [q.shape[0] for q in Q]
    three = len(parts) == 3
    act = r3[:m[2] if three else nz, :m[1], :m[0]]
    t = numpy.einsum('kji,ia->kja', act, Q[0])
    t = numpy.einsum('kja,jb->kba', t, Q[1])
    den = lam[0][None, None, :] + lam[1][None, :, None]
    if three:
        t = numpy.einsum('kba,kc->cba', t, Q[2])
        den = den + lam[2][:, None, None]
    den = coef * den
    t = numpy.where(numpy.abs(den) > 1e-12 * numpy.abs(den).max(), t / numpy.where(den == 0, 1, den), 0.0)
    if three:
        t = numpy.einsum('cba,kc->kba', t, Q[2])
    t = numpy.einsum('kba,jb->kja', t, Q[1])
    t = numpy.einsum('kja,ia->kji', t, Q[0])
    out = -r3.copy()
    out[:act.shape[0], :act.shape[1], :act.shape[2]] = t
    return out


CASES = [
    ('ldc3d', {'Reynolds Number': 0}, 6, 5, 4),
    ('ldc3d_stretched', {'Reynolds Number': 0, 'Grid Stretching Factor': 1.5}, 5, 6, 7),
    ('ldc2d', {'Reynolds Number': 0}, 7, 6, 1),
    ('dhc2d', {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 100.0,
               'Reynolds Number': 1, 'X-max': 0.3}, 6, 7, 1),
    ('rb3d', {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 500.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
              'X-max': 4, 'Y-max': 3, 'Grid Stretching Factor': 1.2}, 5, 4, 6),
]


@pytest.mark.parametrize('name,params,nx,ny,nz', CASES, ids=[c[0] for c in CASES])
def test_fdm_inverts_the_diffusion_blocks(name, params, nx, ny, nz):
    orc = Oracle(dict(params), nx, ny, nz)
    problem = recipes.PROBLEM_IDS[params.get('Problem Type', 'Lid-driven Cavity').lower()]
    cfg = recipes.find_config(problem, orc.dim, nz, orc.dof)
    prm, _ = hostprep.make_params(cfg, problem, params, nx, ny, nz, orc.x, orc.y, orc.z)
    mets = [hostprep.axis_metrics(v, m) for v, m in ((orc.x, nx), (orc.y, ny), (orc.z, nz))]
    ops = hostprep.fdm_operators(cfg, prm, mets, nx, ny, nz)
    J = orc.jacobian_csr(numpy.zeros(orc.n))          # zero state: convection vanishes, pure diffusion blocks
    idx = numpy.arange(orc.n)
    rng = numpy.random.default_rng(0)
    for v in range(orc.dof):
        if v == cfg.p:
            continue
        iv = idx[idx % orc.dof == v]
        block = J[iv][:, iv].tocsc()
        r = rng.standard_normal(len(iv))
        u = fdm_apply(ops, v, nx, ny, nz, r.reshape(nz, ny, nx)).ravel()
        assert numpy.linalg.norm(block @ u - r) <= 1e-10 * numpy.linalg.norm(r), (name, v)
    # pressure: Lp = D M^-1 G, singular (Neumann); the FDM pseudo-inverse solves it on mean-free data
    vel = idx[idx % orc.dof < orc.dim]
    pp = idx[idx % orc.dof == orc.dim]
    mco, mj, _ = orc.mass_matrix()
    M = numpy.zeros(orc.n)
    M[mj] = mco
    import scipy.sparse as sp
    Lp = J[pp][:, vel] @ sp.diags(1 / M[vel]) @ J[vel][:, pp]
    r = rng.standard_normal(len(pp))
    r -= r.mean()
    q = fdm_apply(ops, cfg.p, nx, ny, nz, r.reshape(nz, ny, nx)).ravel()
    assert numpy.linalg.norm(Lp @ q - r) <= 1e-9 * numpy.linalg.norm(r), name


def test_bench_reference_arm_prints_the_contract_line():
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--grid', '16'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['cpu_baseline']['kind'] == 'port' and line['value'] > 0
