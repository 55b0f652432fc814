'''Coupled (w, T) line solve of the Rayleigh-Benard preconditioner (csrc/tfb_joint.h, compiled with g++):
the banded elimination must equal a dense solve of the per-mode system built from the host tables
(hostprep.joint_z_operators) and the vertical couplings of the oracle's Jacobian, and the resulting
approximate inverse of the (velocity, temperature) block must make GMRES on that block converge in a
few steps where the diffusion-only solve does not.  CPU only.'''
import ctypes
import os
import subprocess

import numpy

from oracle.tf_oracle import Oracle
from transiflow_b200 import hostprep, recipes
from test_fdm_cpu import fdm_apply

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
          'X-max': 10, 'Y-max': 10}


def _lib():
    so = os.path.join(HERE, 'cpu_harness', 'libtfjoint.so')
    src = os.path.join(HERE, 'cpu_harness', 'joint_harness.cpp')
    hdr = os.path.join(os.path.dirname(HERE), 'transiflow_b200', 'csrc', 'tfb_joint.h')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(['/usr/bin/g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-o', so, src])
    return ctypes.CDLL(so)


class JointModel:
    '''numpy mirror of joint_solve() in csrc/tfb_solver.cu.'''

    def __init__(self, params, nx, ny, nz, state=None):
        self.orc = orc = Oracle(dict(params), nx, ny, nz)
        self.nx, self.ny, self.nz, self.dof = nx, ny, nz, orc.dof
        problem = recipes.PROBLEM_IDS[params['Problem Type'].lower()]
        self.cfg = cfg = recipes.find_config(problem, orc.dim, nz, orc.dof)
        prm, _ = hostprep.make_params(cfg, problem, params, nx, ny, nz, orc.x, orc.y, orc.z)
        self.mets = mets = [hostprep.axis_metrics(v, m) for v, m in ((orc.x, nx), (orc.y, ny), (orc.z, nz))]
        self.ops = hostprep.fdm_operators(cfg, prm, mets, nx, ny, nz)
        self.cv, self.cT = prm.c_visc, prm.c_T
        tparts = sorted([o for o in self.ops if o[0] == 2], key=lambda o: o[1])     # w's horizontal basis
        self.Qx, self.Qy = tparts[0][3], tparts[1][3]
        self.lx, self.ly = tparts[0][4], tparts[1][4]
        self.zc = numpy.zeros((12, nz))
        self.zc[:8] = hostprep.joint_z_operators(cfg, prm, mets, nz)
        if state is None:   # conduction state: one Newton step from zero (the problem is linear at u = 0)
            from oracle.tf_oracle import direct_solve
            x0 = numpy.zeros(orc.n)
            state = x0 + direct_solve(orc.jacobian_csr(x0), -orc.rhs(x0), orc.dim, orc.dof)
        self.state = state
        self.J = orc.jacobian_csr(state).tocsr()
        self.couplings()

    def couplings(self):
        '''horizontal means of the vertical (w,T) and (T,w) entries, divided by the cell's horizontal area'''
        nx, ny, nz, dof, T = self.nx, self.ny, self.nz, self.dof, self.cfg.T
        hx, hy = self.mets[0][0], self.mets[1][0]
        co = self.J.tocoo()
        rc, rv, cc, cv = co.row // dof, co.row % dof, co.col // dof, co.col % dof
        ri, rj, rk = rc % nx, (rc // nx) % ny, rc // (nx * ny)
        ck = cc // (nx * ny)
        same = (rc % (nx * ny)) == (cc % (nx * ny))
        val = co.data / (hx[ri] * hy[rj]) / (nx * ny)
        for row, (rvar, cvar, dk) in {8: (2, T, 0), 9: (2, T, 1), 10: (T, 2, 0), 11: (T, 2, -1)}.items():
            m = same & (rv == rvar) & (cv == cvar) & (ck - rk == dk)
            if rvar == 2:
                m &= rk < nz - 1
            else:
                m &= ck < nz - 1
            numpy.add.at(self.zc[row], rk[m], val[m])

    def mode_matrix(self, mu):
        nz, z = self.nz, self.zc
        A = numpy.zeros((2 * nz - 1, 2 * nz - 1))
        for k in range(nz):
            i = 2 * k
            A[i, i] = self.cT * (mu * z[7, k] + z[5, k])
            if k > 0:
                A[i, i - 2] = self.cT * z[4, k]
                A[i, i - 1] = z[11, k]
            if k < nz - 1:
                A[i, i + 1] = z[10, k]
                A[i, i + 2] = self.cT * z[6, k]
                j = i + 1
                A[j, j] = self.cv * (mu * z[3, k] + z[1, k])
                A[j, j - 1] = z[8, k]
                A[j, j + 1] = z[9, k]
                if k > 0:
                    A[j, j - 2] = self.cv * z[0, k]
                if k < nz - 2:
                    A[j, j + 2] = self.cv * z[2, k]
        return A

    def solve(self, rw, rT):
        nx, ny, nz = self.nx, self.ny, self.nz
        both = numpy.concatenate([rw.reshape(nz, ny, nx), rT.reshape(nz, ny, nx)])
        t = numpy.einsum('kji,ia->kja', both, self.Qx)
        t = numpy.einsum('kja,jb->kba', t, self.Qy)
        w = numpy.ascontiguousarray(t[:nz].reshape(nz, ny * nx))
        T = numpy.ascontiguousarray(t[nz:].reshape(nz, ny * nx))
        mu = numpy.ascontiguousarray((self.ly[:, None] + self.lx[None, :]).ravel())
        al = numpy.empty((2 * nz, ny * nx))
        be = numpy.empty((2 * nz, ny * nx))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib().tfh_joint_lines(nz, P(self.zc), ny * nx, P(mu), ctypes.c_double(self.cv), ctypes.c_double(self.cT),
                               P(w), P(T), P(al), P(be))
        t = numpy.concatenate([w, T]).reshape(2 * nz, ny, nx)
        t = numpy.einsum('kba,jb->kja', t, self.Qy)
        t = numpy.einsum('kja,ia->kji', t, self.Qx)
        wout = t[:nz].copy()
        wout[nz - 1] = -rw.reshape(nz, ny, nx)[nz - 1]     # wall rows carry a -1 diagonal
        return wout.ravel(), t[nz:].ravel()


def test_hostprep_pencil_refactor_keeps_the_fdm_data():
    m = JointModel(PARAMS, 6, 5, 7)
    Kw, Mw = hostprep._pencil_km('own', m.mets[2], 7, 0.0, 0.0)
    assert numpy.allclose(numpy.diag(Kw), m.zc[1, :6]) and numpy.allclose(Mw, m.zc[3, :6])
    assert numpy.allclose(numpy.diag(Kw, 1), m.zc[2, :5]) and numpy.allclose(numpy.diag(Kw, -1), m.zc[0, 1:6])


def test_line_solve_equals_the_dense_mode_solve():
    m = JointModel(PARAMS, 6, 5, 9)
    nz = m.nz
    rng = numpy.random.default_rng(0)
    mus = numpy.ascontiguousarray(-rng.uniform(0.1, 30.0, 4))
    w = rng.standard_normal((nz, 4)); w[nz - 1] = 0
    T = rng.standard_normal((nz, 4))
    w0, T0 = w.copy(), T.copy()
    al = numpy.empty((2 * nz, 4)); be = numpy.empty((2 * nz, 4))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib().tfh_joint_lines(nz, P(m.zc), 4, P(mus), ctypes.c_double(m.cv), ctypes.c_double(m.cT), P(w), P(T), P(al), P(be))
    for q in range(4):
        A = m.mode_matrix(mus[q])
        r = numpy.empty(2 * nz - 1)
        r[0::2] = T0[:, q]
        r[1::2] = w0[:nz - 1, q]
        y = numpy.linalg.solve(A, r)
        assert numpy.allclose(T[:, q], y[0::2], rtol=1e-9, atol=1e-12)
        assert numpy.allclose(w[:nz - 1, q], y[1::2], rtol=1e-9, atol=1e-12)


def test_factor_then_substitute_equals_the_one_shot_line_solve():
    '''The device path factors every mode once per matrix and only substitutes per application.'''
    m = JointModel(PARAMS, 6, 5, 9)
    nz = m.nz
    rng = numpy.random.default_rng(3)
    mus = numpy.ascontiguousarray(-rng.uniform(0.1, 30.0, 5))
    w = rng.standard_normal((nz, 5)); w[nz - 1] = 0
    T = rng.standard_normal((nz, 5))
    w1, T1, w2, T2 = w.copy(), T.copy(), w.copy(), T.copy()
    al = numpy.empty((2 * nz, 5)); be = numpy.empty((2 * nz, 5))
    fac = numpy.zeros((4, 2 * nz - 1, 5))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L = _lib()
    L.tfh_joint_lines(nz, P(m.zc), 5, P(mus), ctypes.c_double(m.cv), ctypes.c_double(m.cT), P(w1), P(T1), P(al), P(be))
    L.tfh_joint_factor_substitute(nz, P(m.zc), 5, P(mus), ctypes.c_double(m.cv), ctypes.c_double(m.cT), P(w2), P(T2), P(fac))
    assert numpy.allclose(T2, T1, rtol=1e-12, atol=1e-14)
    assert numpy.allclose(w2[:nz - 1], w1[:nz - 1], rtol=1e-12, atol=1e-14)


def _gmres_steps(op, prec, b, k):
    '''relative residuals of right-preconditioned GMRES after 1..k steps'''
    beta = numpy.linalg.norm(b)
    V, Z, out = [b / beta], [], []
    H = numpy.zeros((k + 1, k))
    for j in range(k):
        Z.append(prec(V[j]))
        w = op(Z[j])
        for i in range(j + 1):
            H[i, j] = V[i] @ w
            w = w - H[i, j] * V[i]
        H[j + 1, j] = numpy.linalg.norm(w)
        e = numpy.zeros(j + 2); e[0] = beta
        y = numpy.linalg.lstsq(H[:j + 2, :j + 1], e, rcond=None)[0]
        out.append(numpy.linalg.norm(H[:j + 2, :j + 1] @ y - e) / beta)
        V.append(w / H[j + 1, j])
    return out


def test_joint_solve_preconditions_the_velocity_temperature_block():
    nx, ny, nz = 12, 12, 8
    m = JointModel(PARAMS, nx, ny, nz)
    dof, T = m.dof, m.cfg.T
    idx = numpy.arange(m.orc.n)
    sel = numpy.concatenate([idx[idx % dof < 3], idx[idx % dof == T]])
    nv = 3 * nx * ny * nz
    F = m.J[sel][:, sel]

    def fdm(v, r):
        return fdm_apply(m.ops, v, nx, ny, nz, r.reshape(nz, ny, nx)).ravel()

    def joint(r):
        out = numpy.empty_like(r)
        for v in range(2):
            out[v:nv:3] = fdm(v, r[v:nv:3])
        out[2:nv:3], out[nv:] = m.solve(r[2:nv:3], r[nv:])
        return out

    def diffusion_only(r):
        out = numpy.empty_like(r)
        for v in range(3):
            out[v:nv:3] = fdm(v, r[v:nv:3])
        out[nv:] = fdm(T, r[nv:])
        return out

    b = numpy.random.default_rng(1).standard_normal(len(sel))
    rj = _gmres_steps(lambda x: F @ x, joint, b, 6)
    rd = _gmres_steps(lambda x: F @ x, diffusion_only, b, 6)
    assert rj[-1] < 1e-2, rj
    assert rd[-1] > 10 * rj[-1], (rd, rj)
