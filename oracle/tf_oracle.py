'''TEST INFRASTRUCTURE ONLY -- CPU oracle for the TransiFlow assembly + solve path.

Thin ctypes wrapper around ``oracle/libtforacle.so`` (built from ``tf_oracle.c``
by ``oracle/Makefile``) plus a numpy/scipy restatement of the SciPy backend's
direct solve.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package ``transiflow_b200`` never does.

Parity status: PINNED -- see ``tests/test_oracle_golden.py`` (reference golden
CSR/RHS files and vectors generated from the unmodified Python reference).

Citations are to ``/root/reference/transiflow/<file>:<line>``.
'''
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PROBLEMS = {
    'lid-driven cavity': 0,
    'rayleigh-benard': 1,
    'rayleigh-benard perturbation': 2,
    'differentially heated cavity': 3,
    'double gyre': 4,
    'amoc': 5,
}


class _Problem(ctypes.Structure):
    _fields_ = [
        ('nx', ctypes.c_int32), ('ny', ctypes.c_int32), ('nz', ctypes.c_int32),
        ('dim', ctypes.c_int32), ('dof', ctypes.c_int32),
        ('xper', ctypes.c_int32), ('yper', ctypes.c_int32), ('zper', ctypes.c_int32),
        ('problem', ctypes.c_int32), ('pad_', ctypes.c_int32),
        ('x', ctypes.c_void_p), ('y', ctypes.c_void_p), ('z', ctypes.c_void_p),
        ('Re_lin', ctypes.c_double), ('Re_nl', ctypes.c_double),
        ('Ra', ctypes.c_double), ('Pr', ctypes.c_double), ('Gr', ctypes.c_double), ('Le', ctypes.c_double),
        ('beta', ctypes.c_double), ('Bi', ctypes.c_double), ('asym', ctypes.c_double), ('lidv', ctypes.c_double),
        ('wind', ctypes.c_void_p), ('amoc_tval', ctypes.c_void_p), ('amoc_sval', ctypes.c_void_p),
    ]


def build(force=False):
    '''Compile the C restatement (idempotent).'''
    so = os.path.join(_HERE, 'libtforacle.so')
    src = os.path.join(_HERE, 'tf_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libtforacle.so'],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.tfo_rhs.restype = ctypes.c_int
        _LIB.tfo_jacobian.restype = ctypes.c_int
        _LIB.tfo_mass_matrix.restype = ctypes.c_int
        _LIB.tfo_num_threads.restype = ctypes.c_int
        _LIB.tfo_set_num_threads.argtypes = [ctypes.c_int]
        _LIB.tfo_set_num_threads.restype = None
        _LIB.tfo_free.argtypes = [ctypes.c_void_p]
    return _LIB


# ---- coordinate vectors: utils.py:175-267 ----

def uniform_coordinate_vector(start, end, nx):
    dx = (end - start) / nx
    x = start + numpy.arange(-1, nx + 2) * dx
    return numpy.roll(x, -2)


def _fix_ghosts(x, start, end):
    dx = x[0] - x[-1]
    if start == 0:
        x[-2] = x[-1] - dx
    if end == 1:
        x[-3] = x[-4] + dx
    return x


def tanh_coordinate_vector(start, end, nx, sigma):
    x = uniform_coordinate_vector(0, 1, nx)
    x = 0.5 * (1 + numpy.tanh(2 * sigma * (x - 0.5)) / numpy.tanh(sigma))
    x = start + x * (end - start)
    return _fix_ghosts(x, start, end)


def sin_coordinate_vector(start, end, nx, sigma):
    x = uniform_coordinate_vector(0, 1, nx)
    x = x - sigma * numpy.sin(2 * numpy.pi * x)
    x = start + x * (end - start)
    return _fix_ghosts(x, start, end)


def coordinate_vector(parameters, start, end, n):
    '''Discretization.get_coordinate_vector, Discretization.py:186-208'''
    if parameters.get('Grid Stretching', False) or 'Grid Stretching Factor' in parameters.keys():
        if parameters.get('Grid Stretching Method', 'tanh') == 'sin':
            return sin_coordinate_vector(start, end, n, parameters.get('Grid Stretching Factor', 0.1))
        return tanh_coordinate_vector(start, end, n, parameters.get('Grid Stretching Factor', 1.5))
    return uniform_coordinate_vector(start, end, n)


def _centers(vec):
    '''utils.compute_coordinate_vector_centers, utils.py:269-288'''
    x = numpy.zeros(len(vec) - 1)
    for i in range(-1, len(vec) - 2):
        x[i] = (vec[i] + vec[i - 1]) / 2
    return x


class Oracle:
    '''Mirror of ``transiflow.Discretization`` (constructor Discretization.py:106-143) on top
    of the C restatement. ``parameters`` is shared by reference and re-read on every call.'''

    def __init__(self, parameters, nx, ny, nz=1, dim=None, dof=None, x=None, y=None, z=None):
        self.parameters = parameters
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dim = dim if dim is not None else (3 if nz > 1 else 2)
        ptype = parameters.get('Problem Type', 'Lid-driven Cavity').lower()
        if ptype not in PROBLEMS:
            raise Exception('Invalid problem type %s' % parameters.get('Problem Type'))
        self.problem = PROBLEMS[ptype]
        if dof is None:  # set_dof, Discretization.py:559-573
            dof = self.dim + 1
            if self.problem in (1, 2, 3):
                dof = self.dim + 2
            elif self.problem == 5:
                dof = self.dim + 3
        self.dof = dof
        self.zper = nz == 1
        p = parameters
        self.x = coordinate_vector(p, p.get('X-min', 0.0), p.get('X-max', 1.0), nx) if x is None else x
        self.y = coordinate_vector(p, p.get('Y-min', 0.0), p.get('Y-max', 1.0), ny) if y is None else y
        self.z = coordinate_vector(p, p.get('Z-min', 0.0), p.get('Z-max', 1.0), nz) if z is None else z
        self.x = numpy.ascontiguousarray(self.x, dtype=numpy.float64)
        self.y = numpy.ascontiguousarray(self.y, dtype=numpy.float64)
        self.z = numpy.ascontiguousarray(self.z, dtype=numpy.float64)
        self.n = nx * ny * nz * self.dof

    def _get(self, name, default=0):
        return self.parameters.get(name, default) if name in self.parameters else default

    def _wind_stress(self):
        '''Discretization.wind_stress, Discretization.py:1101-1118'''
        alpha = self._get('Wind Stress Parameter')
        asym = self._get('Asymmetry Parameter')
        frc = numpy.zeros((self.nx, self.ny, self.nz, self.dof))
        for i, j, k in numpy.ndindex(self.nx - 1, self.ny, self.nz):
            dx = (self.x[i + 1] - self.x[i - 1]) / 2
            dy = self.y[j] - self.y[j - 1]
            dz = self.z[k] - self.z[k - 1]
            y = (self.y[j] + self.y[j - 1]) / 2
            frc[i, j, k, 0] = - (1 - asym) * numpy.cos(2 * numpy.pi * y) - asym * numpy.cos(numpy.pi * y)
            frc[i, j, k, 0] *= alpha / (2 * numpy.pi) * dx * dy * dz
        return numpy.ascontiguousarray(frc.transpose(2, 1, 0, 3)).ravel()

    def _problem(self):
        keep = []
        pr = _Problem()
        pr.nx, pr.ny, pr.nz, pr.dim, pr.dof = self.nx, self.ny, self.nz, self.dim, self.dof
        pr.xper, pr.yper, pr.zper = 0, 0, int(self.zper)
        pr.problem = self.problem
        pr.x, pr.y, pr.z = self.x.ctypes.data, self.y.ctypes.data, self.z.ctypes.data
        pr.Re_lin = self._get('Reynolds Number', 1.0)
        pr.Re_nl = self._get('Reynolds Number')
        Ra = self._get('Rayleigh Number', 1.0)
        Pr = self._get('Prandtl Number', 1.0)
        pr.Ra, pr.Pr = Ra, Pr
        pr.Gr = self._get('Grashof Number', Ra / Pr)
        pr.Le = self._get('Lewis Number', 1.0)
        pr.beta = self._get('Rossby Parameter')
        pr.Bi = self._get('Biot Number')
        pr.asym = self._get('Asymmetry Parameter')
        pr.lidv = self._get('Lid Velocity', 1)
        if self.problem == 4:
            w = self._wind_stress()
            keep.append(w)
            pr.wind = w.ctypes.data
        if self.problem == 5:  # Discretization._amoc, Discretization.py:670-686
            xc = _centers(self.x)
            theta = self._get('Temperature Forcing')
            asym = self._get('Asymmetry Parameter')
            A = self.parameters.get('X-max', 1.0)
            T_S = numpy.zeros((self.nx + 2, self.nz + 2))
            T_S[:, 0] = 1 / 2 * ((1 - asym) * numpy.cos(2 * numpy.pi * (xc / A - 1 / 2))
                                 + asym * numpy.cos(numpy.pi * xc / A) + 1)
            tval = numpy.ones((self.nx + 2, self.nz + 2)) * (2 * (theta * T_S))
            sigma = self._get('Freshwater Flux')
            pp = 2
            Q_S = numpy.zeros((self.nx + 2, self.nz + 2))
            Q_S[:, 0] = 3 * numpy.cos(pp * numpy.pi * (xc / A - 1 / 2)) - 6 / (pp * numpy.pi) * numpy.sin(pp * numpy.pi / 2)
            h = (self.y[self.ny] - self.y[self.ny - 2]) / 2
            sval = numpy.ones((self.nx + 2, self.nz + 2)) * (h * (sigma * Q_S))
            tval = numpy.ascontiguousarray(tval)
            sval = numpy.ascontiguousarray(sval)
            keep += [tval, sval]
            pr.amoc_tval, pr.amoc_sval = tval.ctypes.data, sval.ctypes.data
        return pr, keep

    def rhs(self, state):
        state = numpy.ascontiguousarray(state, dtype=numpy.float64)
        out = numpy.zeros(self.n)
        pr, keep = self._problem()
        lib().tfo_rhs(ctypes.byref(pr), ctypes.c_void_p(state.ctypes.data), ctypes.c_void_p(out.ctypes.data))
        return out

    def jacobian(self, state):
        '''Returns (coA, jcoA, begA) exactly like the reference's compressed CrsMatrix.'''
        state = numpy.ascontiguousarray(state, dtype=numpy.float64)
        begA = numpy.zeros(self.n + 1, dtype=numpy.int64)
        jp, cp = ctypes.c_void_p(), ctypes.c_void_p()
        pr, keep = self._problem()
        lib().tfo_jacobian(ctypes.byref(pr), ctypes.c_void_p(state.ctypes.data),
                           ctypes.c_void_p(begA.ctypes.data), ctypes.byref(jp), ctypes.byref(cp))
        nnz = int(begA[-1])
        jcoA = numpy.ctypeslib.as_array(ctypes.cast(jp, ctypes.POINTER(ctypes.c_int64)), (max(nnz, 1),))[:nnz].copy()
        coA = numpy.ctypeslib.as_array(ctypes.cast(cp, ctypes.POINTER(ctypes.c_double)), (max(nnz, 1),))[:nnz].copy()
        lib().tfo_free(jp)
        lib().tfo_free(cp)
        return coA, jcoA, begA

    def mass_matrix(self):
        begA = numpy.zeros(self.n + 1, dtype=numpy.int64)
        jcoA = numpy.zeros(self.n, dtype=numpy.int64)
        coA = numpy.zeros(self.n)
        pr, keep = self._problem()
        lib().tfo_mass_matrix(ctypes.byref(pr), ctypes.c_void_p(begA.ctypes.data),
                              ctypes.c_void_p(jcoA.ctypes.data), ctypes.c_void_p(coA.ctypes.data))
        nnz = int(begA[-1])
        return coA[:nnz], jcoA[:nnz], begA

    def jacobian_csr(self, state):
        from scipy import sparse
        coA, jcoA, begA = self.jacobian(state)
        return sparse.csr_matrix((coA, jcoA, begA), (self.n, self.n))


def direct_solve(jac_csr, rhs, dim, dof, pressure_row=None):
    '''Restatement of SciPy.Interface.direct_solve without border (SciPy.py:51-129,204-258):
    pin the pressure at row ``dim`` (row -> -1 on the diagonal, column dropped, rhs entry 0)
    and solve with SuperLU (scipy.sparse.linalg.splu, third-party, same package the
    reference calls at SciPy.py:151).'''
    from scipy import sparse
    from scipy.sparse import linalg
    A = sparse.csr_matrix(jac_csr).tolil(copy=True)
    x = numpy.array(rhs, dtype=numpy.float64, copy=True)
    if dof > dim:
        row = dim if pressure_row is None else pressure_row
        x[row] = 0
        A[row, :] = 0
        A[:, row] = 0
        A[row, row] = -1.0
    lu = linalg.splu(sparse.csc_matrix(A))
    return lu.solve(x)
