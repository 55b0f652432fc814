'''CPU prototype (diagnostic, numpy/scipy on the oracle's matrices): GMRES iteration counts of the block preconditioner with
the scaled-mass Schur complement as a function of the scaling factor in gamma = factor * rho * |c_visc|, against the
least-squares commutator -- the experiment behind the default 2.5 (csrc/tfb_solver.cu: schur_gamma_refresh).  The velocity
sub-solve is the exact solve with the diffusion block (what the FDM solve computes).

    python tools/proto_scaled_mass.py [N] [Re]
'''
import os
import sys

import numpy
from scipy import sparse
from scipy.sparse import linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.tf_oracle import Oracle, direct_solve  # noqa: E402


def pinned(J, dim):
    A = sparse.csr_matrix(J).tolil(copy=True)
    A[dim, :] = 0
    A[:, dim] = 0
    A[dim, dim] = -1.0
    return sparse.csr_matrix(A)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    Re = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    params = {'Reynolds Number': Re}
    orc = Oracle(dict(params), N, N, N)
    dim, dof, n = orc.dim, orc.dof, orc.n
    x = numpy.zeros(n)
    for _ in range(3):
        x = x + direct_solve(orc.jacobian_csr(x), -orc.rhs(x), dim, dof)
    A = pinned(orc.jacobian_csr(x), dim)
    A0 = pinned(orc.jacobian_csr(numpy.zeros(n)), dim)          # no convection: the diffusion blocks the FDM solves invert
    b = -orc.rhs(x)
    b[dim] = 0
    var = numpy.arange(n) % dof
    iu, ip = numpy.nonzero(var < dim)[0], numpy.nonzero(var == dim)[0]
    Auu, Ah = A[iu][:, iu].tocsc(), A0[iu][:, iu].tocsc()
    G, D = A[iu][:, ip].tocsc(), A[ip][:, iu].tocsc()
    lu_h = linalg.splu(Ah)
    # cell volumes from the metric of a uniform grid (oracle default): h^3
    vol = numpy.full(len(ip), (1.0 / N) ** 3)
    cvisc = 1.0 / Re
    # rho(Auu Ah^-1) by power iteration
    v = numpy.random.default_rng(0).standard_normal(len(iu))
    for _ in range(30):
        w = Auu @ lu_h.solve(v)
        rho = numpy.linalg.norm(w) / numpy.linalg.norm(v)
        v = w / numpy.linalg.norm(w)
    print('N=%d Re=%g n=%d rho=%.3f' % (N, Re, n, rho))
    pin = numpy.nonzero(ip == dim)[0][0]

    def run(prec):
        its = [0]

        def cb(_):
            its[0] += 1
        M = linalg.LinearOperator((n, n), matvec=prec)
        y, info = linalg.gmres(A, b, M=M, rtol=1e-10, atol=0, restart=600, maxiter=600, callback=cb, callback_type='pr_norm')
        return its[0], numpy.linalg.norm(b - A @ y) / numpy.linalg.norm(b)

    for factor in (1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 6.0):
        gamma = factor * rho * cvisc

        def prec(r, gamma=gamma):
            z = numpy.zeros(n)
            dp = gamma * r[ip] / vol
            dp[pin] = -r[ip][pin]
            z[ip] = dp
            z[iu] = lu_h.solve(r[iu] - G @ dp)
            return z
        k, res = run(prec)
        print('  scaled mass factor %.1f: %4d iterations, relres %.1e' % (factor, k, res))


if __name__ == '__main__':
    main()
