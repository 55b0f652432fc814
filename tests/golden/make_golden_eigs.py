'''Golden for Interface.eigs: the configuration of the reference's own eigenvalue test
(tests/jada_fixtures.py:19-85, tests/test_jada.py:108-122): 2-D lid-driven cavity 6x6, continuation to
Re = 2000 with the unmodified reference SciPy backend, then the 10 generalized eigenvalues of
J v = lambda M v closest to zero from ARPACK shift-invert (sigma = 0.1, the reference fixture's call) and,
as a cross-check, from dense QZ.  Run here (reference importable):

    PYTHONPATH=/root/reference python tests/golden/make_golden_eigs.py
'''
import os
import sys

import numpy
import scipy.linalg
from scipy import sparse
from scipy.sparse import linalg as spla

sys.path.insert(0, '/root/reference')
from transiflow import Continuation                      # noqa: E402
from transiflow.interface.SciPy import Interface         # noqa: E402

nx = 6
num = 10
interface = Interface({}, nx, nx)
cont = Continuation(interface)
x0 = cont.newton(numpy.zeros(interface.dof * nx * nx))
x = cont.continuation(x0, 'Reynolds Number', 0, 2000, 100)[0]
x = cont.newton(x, 1e-12)
J = interface.jacobian(x)
M = interface.mass_matrix()
# pressure-pinned pencil (SciPy.py:212-216) -- the raw J - sigma M is exactly singular (constant pressure)
Jp = sparse.lil_matrix(J)
Jp[interface.dim, :] = 0
Jp[:, interface.dim] = 0
Jp[interface.dim, interface.dim] = -1
ev, v = spla.eigs(Jp.tocsc(), num, M, sigma=0.1, tol=1e-10)
ev = numpy.array(sorted(ev, key=lambda z: abs(z)))
# dense cross-check
lam = scipy.linalg.eig(Jp.toarray(), M.toarray(), right=False)
lam = lam[numpy.isfinite(lam)]
lam = numpy.array(sorted(lam, key=lambda z: abs(z)))[:num]
print('ARPACK', ev)
print('dense ', lam)
print('max |arpack - dense| (matched):', numpy.abs(ev[:, None] - lam[None, :]).min(axis=1).max())
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated', 'eigs_ldc2d_6_re2000.npz')
numpy.savez(out, x=x, eigs_arpack=ev, eigs_dense=lam, reynolds=numpy.array(interface.get_parameter('Reynolds Number')))
print('wrote', out)
