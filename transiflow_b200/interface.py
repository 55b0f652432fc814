'''``transiflow.interface`` backend for NVIDIA B200 (sm_100a).

``Interface`` keeps the API surface of the reference's backends (``vector, rhs, jacobian,
mass_matrix, solve, eigs`` plus the inherited parameter / save / load helpers of
``transiflow/interface/BaseInterface.py:30-215`` and the constructor signature of
``transiflow/interface/SciPy.py:25-26``), so ``Continuation`` and ``TimeIntegration`` drive
it unchanged, while assembly and the linear solve run as CUDA kernels behind the C ABI in
``include/tfb200.h``.  The host side is numpy + ctypes only.

Vectors are plain ``numpy.ndarray`` (like the SciPy backend, SciPy.py:37-38).  Matrices are
``DeviceMatrix`` handles whose values live in HBM on a fixed structural sparsity pattern.
'''
import ctypes
import json
import os

import numpy

from . import _lib, hostprep, recipes
from ._lib import as_f64, check, ptr


class ParameterEncoder(json.JSONEncoder):
    '''JSON form of a parameter set in the reference's file convention (BaseInterface.py:10-29): numpy scalars as
    Python numbers, arrays as lists, complex values as ``{'__complex__': true, 'real': .., 'imag': ..}``.'''

    def default(self, obj):
        if isinstance(obj, numpy.integer):
            return int(obj)
        if isinstance(obj, numpy.floating):
            return float(obj)
        if isinstance(obj, (complex, numpy.complexfloating)):
            return {'__complex__': True, 'real': float(obj.real), 'imag': float(obj.imag)}
        if isinstance(obj, numpy.ndarray):
            return obj.tolist()
        return super().default(obj)


def parameter_decoder(obj):
    '''object_hook that restores the complex values ParameterEncoder wrote.'''
    if '__complex__' in obj:
        return complex(obj['real'], obj['imag'])
    return obj


class _DiscretizationInfo:
    '''The attributes of ``interface.discretization`` that reference utilities read (utils.create_padded_state_mtx,
    utils.py:105-107; BaseInterface.py:58-63): grid, coordinate vectors and the periodicity flags of
    Discretization.py:121-128.  The discretization itself lives in the CUDA kernels.'''

    def __init__(self, interface):
        self.nx, self.ny, self.nz = interface.nx, interface.ny, interface.nz
        self.dim, self.dof = interface.dim, interface.dof
        self.x, self.y, self.z = interface.x, interface.y, interface.z
        self.x_periodic = False
        self.y_periodic = False
        self.z_periodic = interface.nz == 1
        self.parameters = interface.parameters


class DeviceMatrix:
    '''Jacobian on the device: CSR values on the Interface's fixed structural pattern.

    Behaves like the ``scipy.sparse`` matrix the SciPy backend returns for what the callers
    use (``J @ x``, ``shape``, ``dtype``; arithmetic such as ``J - M / s`` goes through a lazy
    ``tocsc()`` copy).  ``tocsr()/tocsc()`` drop entries with ``|v| <= 1e-14`` exactly like
    ``CrsMatrix.compress`` (CrsMatrix.py:52-71) so the result is identical to the reference's
    value-dependent pattern.'''

    def __init__(self, interface):
        self.interface = interface
        h = ctypes.c_void_p()
        check(_lib.lib().tfb_mat_create(interface._ctx, ctypes.byref(h)))
        self._h = h
        self.shape = (interface.n, interface.n)
        self.dtype = numpy.dtype(numpy.float64)
        self._host = None
        self._shift = 0.0                # the matrix is J + _shift * M (M: mass matrix), see _plus_diagonal

    @property
    def data(self):
        '''Values of the compressed matrix -- what ``scipy.sparse`` matrices expose and the reference's JaDa glue
        inspects for its dtype (``mat.data.dtype``, JaDa.py:27).'''
        return self.tocsr().data

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib._LIB is not None:
            _lib._LIB.tfb_mat_destroy(h)

    @classmethod
    def from_scipy(cls, interface, A):
        '''Upload a host matrix whose entries lie inside the structural pattern (e.g. the
        ``J - M / (theta * dt)`` of TimeIntegration.py:58 or a real shifted matrix
        ``beta * J - alpha * M`` of the eigen-solver glue) so that ``solve`` can run on the device.'''
        from scipy import sparse
        interface._require_whole_grid('DeviceMatrix.from_scipy')
        A = sparse.csr_matrix(A)
        if A.shape != (interface.n, interface.n):
            raise NotImplementedError('only matrices of the Interface size can be uploaded')
        if numpy.iscomplexobj(A.data):
            if not numpy.any(A.imag.data):
                A = A.real
            else:
                return ComplexDeviceMatrix(cls.from_scipy(interface, A.real), cls.from_scipy(interface, A.imag))
        A.sum_duplicates()
        A.sort_indices()
        row_ptr, col = interface.pattern()
        n = interface.n
        rows_s = numpy.repeat(numpy.arange(n, dtype=numpy.int64), numpy.diff(row_ptr))
        rows_a = numpy.repeat(numpy.arange(n, dtype=numpy.int64), numpy.diff(A.indptr))
        key_s = rows_s * n + col                       # sorted: rows ascending, columns ascending within a row
        key_a = rows_a * n + A.indices
        pos = numpy.searchsorted(key_s, key_a)
        ok = (pos < len(key_s)) & (key_s[numpy.minimum(pos, len(key_s) - 1)] == key_a)
        if not numpy.all(ok | (A.data == 0)):
            raise NotImplementedError('matrix has entries outside the structural pattern of this discretization')
        vals = numpy.zeros(interface.nnz)
        vals[pos[ok]] = A.data[ok]
        mat = cls(interface)
        check(_lib.lib().tfb_mat_set_values(mat._h, ptr(vals)))
        return mat

    def values(self):
        '''CSR values of the full structural pattern (explicit zeros included), D2H copy.'''
        out = numpy.empty(self.interface.nnz)
        check(_lib.lib().tfb_mat_get_values(self._h, ptr(out)))
        return out

    def structural_csr(self):
        from scipy import sparse
        row_ptr, col = self.interface.pattern()
        return sparse.csr_matrix((self.values(), col, row_ptr), self.shape)

    def tocsr(self):
        if self._host is None:
            from scipy import sparse
            row_ptr, col = self.interface.pattern()
            vals = self.values()
            keep = numpy.abs(vals) > 1e-14
            # kept entries per row through one cumulative sum (every structural row is non-empty)
            csum = numpy.concatenate(([0], numpy.cumsum(keep, dtype=numpy.int64)))
            self._host = sparse.csr_matrix((vals[keep], col[keep], csum[row_ptr]), self.shape)
        return self._host

    def tocsc(self):
        return self.tocsr().tocsc()

    def __matmul__(self, x):
        if numpy.iscomplexobj(x) and numpy.ndim(x) == 1:
            x = numpy.asarray(x)
            return (self @ numpy.ascontiguousarray(x.real)) + 1j * (self @ numpy.ascontiguousarray(x.imag))
        if numpy.ndim(x) != 1:
            return self.tocsr() @ x
        x = as_f64(x)
        y = numpy.empty_like(x)
        check(_lib.lib().tfb_spmv(self._h, ptr(x), ptr(y)))
        return y

    def __mul__(self, other):
        if numpy.isscalar(other):
            return self.tocsc() * other
        return self @ other

    def __rmul__(self, other):
        return self.tocsc() * other

    def __neg__(self):
        return -self.tocsc()

    def _plus_diagonal(self, other, sign):
        '''J +/- D for a diagonal host matrix D (e.g. ``mass / (theta * dt)``): stays on the device.'''
        from scipy import sparse
        if not sparse.issparse(other) or other.shape != self.shape or numpy.iscomplexobj(other):
            return None
        coo = sparse.coo_matrix(other)
        if coo.nnz and numpy.any(coo.row != coo.col):
            return None
        d = numpy.zeros(self.shape[0])
        numpy.add.at(d, coo.row, coo.data)
        out = DeviceMatrix(self.interface)
        check(_lib.lib().tfb_mat_add_diag(out._h, self._h, ctypes.c_double(sign), ptr(d)))
        # J + a M with M the mass matrix (time stepping: -1 / (theta dt); shifted eigenproblems: -sigma): tell the solver,
        # whose fast-diagonalisation basis is M-orthonormal, so that the preconditioner carries the shift exactly
        mass = self.interface._mass_diagonal()
        nz = mass != 0
        out._shift = self._shift
        if numpy.any(nz) and not numpy.any(d[~nz]):
            ratio = d[nz] / mass[nz]
            if numpy.ptp(ratio) <= 1e-12 * max(abs(ratio[0]), 1e-300):
                out._shift = self._shift + sign * float(ratio[0])
                check(_lib.lib().tfb_mat_set_shift(out._h, ctypes.c_double(out._shift)))
        return out

    def __add__(self, other):
        dev = self._plus_diagonal(other, 1.0)
        return dev if dev is not None else self.tocsc() + _host_matrix(other)

    def __radd__(self, other):
        return _host_matrix(other) + self.tocsc()

    def __sub__(self, other):
        dev = self._plus_diagonal(other, -1.0)
        return dev if dev is not None else self.tocsc() - _host_matrix(other)

    def __rsub__(self, other):
        return _host_matrix(other) - self.tocsc()

    def __truediv__(self, other):
        return self.tocsc() / other


def _host_matrix(m):
    return m.tocsc() if isinstance(m, DeviceMatrix) else m


class ComplexDeviceMatrix:
    '''A complex matrix on the structural pattern, e.g. the shifted matrix ``beta * J - alpha * M`` of the eigen-solver
    glue with a complex shift (JaDa.py:90,187): real and imaginary parts as two DeviceMatrix handles.  ``solve`` treats
    it with a complex Krylov iteration whose products and preconditioner applications run on the device.'''

    def __init__(self, real, imag):
        self.real, self.imag = real, imag
        self.shape = real.shape
        self.dtype = numpy.dtype(numpy.complex128)

    def __matmul__(self, x):
        x = numpy.asarray(x)
        xr, xi = numpy.ascontiguousarray(x.real, dtype=float), numpy.ascontiguousarray(x.imag, dtype=float)
        return (self.real @ xr - self.imag @ xi) + 1j * (self.imag @ xr + self.real @ xi)


class Interface:
    '''B200 backend.  Constructor arguments as ``transiflow.interface.SciPy.Interface``
    (SciPy.py:25-26) plus ``device``.'''

    AUTO_IDR_MIN_UNKNOWNS = 500000   # 'Method': 'auto' picks IDR(8) from this many unknowns on (3-D grids)

    def __init__(self, parameters, nx, ny, nz=1, dim=None, dof=None,
                 x=None, y=None, z=None, boundary_conditions=None, device=0, slab=None):
        self.parameters = parameters
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dim = dim if dim is not None else (3 if nz > 1 else 2)       # Discretization.py:115-117
        ptype = parameters.get('Problem Type', 'Lid-driven Cavity').lower()  # :550-557
        if ptype not in recipes.PROBLEM_IDS:
            raise Exception('Invalid problem type %s' % parameters.get('Problem Type'))  # :733
        self.problem = recipes.PROBLEM_IDS[ptype]
        if dof is None:                                                    # set_dof, :559-573
            dof = self.dim + 1
            if self.problem in (recipes.RB, recipes.RBP, recipes.DHC):
                dof = self.dim + 2
            elif self.problem == recipes.AMOC:
                dof = self.dim + 3
        self.dof = dof
        if boundary_conditions is not None:
            # Discretization.py:62-66,719: the callback is recorded once; it runs on the device if its sequence of ops has
            # the structure of a generated recipe (its constants are run-time arguments of the kernels), see recipes.py
            ops = recipes.record_boundary_conditions(boundary_conditions)
            self.config = recipes.match_recorded(ops, self.dim, nz, dof)
            if self.config is None:
                raise NotImplementedError(
                    'this boundary_conditions callback applies a sequence of ops for which no kernel family is generated '
                    '(dim=%d nz=%d dof=%d): %r' % (self.dim, nz, dof, [op[:4] for op in ops]))
        else:
            self.config = recipes.find_config(self.problem, self.dim, nz, dof)
        if self.config is None:
            raise NotImplementedError('no B200 kernel family for problem=%r dim=%d nz=%d dof=%d'
                                      % (ptype, self.dim, nz, dof))
        p = parameters
        cv = hostprep.coordinate_vector
        self.x = cv(p, p.get('X-min', 0.0), p.get('X-max', 1.0), nx) if x is None else x
        self.y = cv(p, p.get('Y-min', 0.0), p.get('Y-max', 1.0), ny) if y is None else y
        self.z = cv(p, p.get('Z-min', 0.0), p.get('Z-max', 1.0), nz) if z is None else z
        self.n = nx * ny * nz * dof
        self.pressure_row = self.dim                                       # SciPy.py:30
        self.border_scaling = 1e-3                                         # SciPy.py:33
        self.device = device
        self._subspaces = None
        self.discretization = _DiscretizationInfo(self)

        L = _lib.lib()
        if L.tfb_device_count() <= 0:
            raise RuntimeError('transiflow_b200 needs a CUDA device (no CPU fallback)')
        self._mets = [hostprep.axis_metrics(v, m) for v, m in ((self.x, nx), (self.y, ny), (self.z, nz))]
        self._cor = hostprep.coriolis_metrics(self.y, ny)
        d = _lib.TfbDesc()
        d.config, d.nx, d.ny, d.nz, d.dim, d.dof = self.config.cid, nx, ny, nz, self.dim, dof
        self.slab = (0, nz) if slab is None else (int(slab[0]), int(slab[1]))
        d.device, d.k0, d.k1 = device, self.slab[0], self.slab[1]
        for a in range(3):
            d.met[a] = self._mets[a].ctypes.data
        d.cor = self._cor.ctypes.data
        self._ctx = ctypes.c_void_p()
        check(L.tfb_create(ctypes.byref(d), ctypes.byref(self._ctx)))
        nnz, nloc, row0 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(L.tfb_sizes(self._ctx, ctypes.byref(nloc), ctypes.byref(nnz), None, ctypes.byref(row0)))
        self.nnz, self.n_local, self.row0 = nnz.value, nloc.value, row0.value
        self._pattern = None
        self._resident_key = None
        self._param_key = None
        self.last_solve = None
        # result vectors of rhs / jacobian_rhs / solve come from recycled page-locked buffers on large grids
        self._pool = _lib.PinnedPool(self.n_local) if self.n_local >= (1 << 20) else None

    def __del__(self):
        ctx, self._ctx = getattr(self, '_ctx', None), None
        pool, self._pool = getattr(self, '_pool', None), None
        if pool is not None:
            pool.close()
        if ctx and _lib._LIB is not None:
            _lib._LIB.tfb_destroy(ctx)

    def _require_whole_grid(self, what):
        '''The numpy layer above the C ABI is not distributed: on a z-slab only assembly, `jac @ x` free solves and
        the Krylov solve are.'''
        if self.slab != (0, self.nz):
            raise NotImplementedError('%s is not available on a z-slab Interface (slab=%r of %d planes)'
                                      % (what, self.slab, self.nz))

    def _result_vector(self):
        return self._pool.empty() if self._pool is not None else numpy.empty(self.n_local)

    # ---- parameters (BaseInterface.py:84-104; Discretization.py:145-184) ----
    def _debug_print(self, *args):
        if self.parameters.get('Verbose', False):
            print('Debug:', *args, flush=True)

    def _debug_print_residual(self, string, jac, x, rhs):
        # BaseInterface.py:36-43
        if self.parameters.get('Verbose', False):
            r = jac @ x - rhs
            self._debug_print(string, numpy.sqrt(abs(numpy.vdot(r, r))))

    def set_parameter(self, name, value):
        if name in self.parameters and self.get_parameter(name) == value:      # Discretization.py:158-161
            return
        self.parameters[name] = value

    def get_parameter(self, name, default=0):
        return self.parameters.get(name, default) if name in self.parameters else default

    def _sync_params(self):
        '''Re-evaluate the kernel scalars whenever the shared, mutable parameter dict changed.'''
        key = json.dumps(self.parameters, sort_keys=True, default=repr)
        if key == self._param_key:
            return
        prm, arrays = hostprep.make_params(self.config, self.problem, self.parameters,
                                           self.nx, self.ny, self.nz, self.x, self.y, self.z)
        fval = (ctypes.c_void_p * hostprep.TFB_MAX_FORCE)()
        fdir = (ctypes.c_int8 * hostprep.TFB_MAX_FORCE)()
        fi = 0
        for op in self.config.recipe:
            if op[0] == 'force':
                fdir[fi] = op[1]
                if fi in arrays:
                    fval[fi] = arrays[fi].ctypes.data
                fi += 1
        wind = None
        if self.problem == recipes.QG:
            wind = as_f64(hostprep.wind_stress(self.parameters, self.nx, self.ny, self.nz, self.dof,
                                               self.x, self.y, self.z))
        check(_lib.lib().tfb_set_params(self._ctx, ctypes.byref(prm), fval, fdir, ptr(wind)))
        self._param_key = key
        self._prm = prm
        self._fdm_key = None

    def _sync_solver(self):
        '''Upload the fast-diagonalisation data of the preconditioner (grid + parameter dependent).'''
        self._sync_params()
        if self._fdm_key == self._param_key:
            return
        # semi-2D grids (dim = 3, nz = 1): the z-offsets fold onto the cell itself, so the z parts of every diffusion
        # stencil cancel and the sub-solves are the 2-D fast-diagonalisation solves (w is one more scalar in x and y)
        pencils = []
        for v, a, m, Q, lam, coef in hostprep.fdm_operators(self.config, self._prm, self._mets, self.nx, self.ny, self.nz, pencils):
            check(_lib.lib().tfb_fdm_set(self._ctx, v, a, m, ptr(Q), ptr(lam), ctypes.c_double(coef)))
        for v, a, m, lower, diag, upper, mass in pencils:
            check(_lib.lib().tfb_fdm_set_pencil(self._ctx, v, a, m, ptr(lower), ptr(diag), ptr(upper), ptr(mass)))
        if self.problem == recipes.AMOC:
            check(_lib.lib().tfb_fdm_pin(self._ctx, self.config.S, ctypes.c_int64(0), ctypes.c_double(-1.0)))
        # Rayleigh-Benard on a true 3-D grid: w and T are solved together along z (csrc/tfb_joint.h)
        self._joint = self.problem in (recipes.RB, recipes.RBP) and self.dim == 3 and self.nz > 2 and self.dof == 5
        if self._joint:
            zops = hostprep.joint_z_operators(self.config, self._prm, self._mets, self.nz)
            check(_lib.lib().tfb_joint_set(self._ctx, self.dim - 1, self.config.T, self.nz, ptr(zops)))
        self._fdm_key = self._param_key

    # ---- vectors (SciPy.py:37-38; BaseInterface.py:84-92) ----
    def vector(self):
        return numpy.zeros(self.n_local)

    def vector_from_array(self, array):
        return array

    def array_from_vector(self, vector):
        return vector

    # ---- assembly ----
    def pattern(self):
        '''(row_ptr, col_idx) of the fixed structural pattern, int64 like CrsMatrix.begA/jcoA.'''
        if self._pattern is None:
            row_ptr = numpy.empty(self.n_local + 1, dtype=numpy.int64)
            col = numpy.empty(self.nnz, dtype=numpy.int64)
            check(_lib.lib().tfb_get_pattern(self._ctx, ptr(row_ptr), ptr(col)))
            self._pattern = (row_ptr, col)
        return self._pattern

    # ---- device-resident state: rhs(x; mu), rhs(x; mu + delta) and jacobian(x) of one corrector
    # iteration (Continuation.py:126,145-150) upload x once.  A host vector is recognised by its address, length and a
    # 64-bit checksum of all its words (any single changed entry changes it), so in-place edits are never missed.
    def _ensure_state(self, state):
        '''Make `state` the state resident in HBM; returns True if it had to be uploaded.'''
        L = _lib.lib()
        key = None
        if self.parameters.get('State Cache', True):
            cs = (ctypes.c_uint64 * 2)()
            check(L.tfb_host_checksum(ptr(state), ctypes.c_int64(state.size), cs))
            key = (state.ctypes.data, state.size, cs[0], cs[1])
            if key == self._resident_key:
                return False
        check(L.tfb_state_upload(self._ctx, ptr(state)))
        self._resident_key = key
        return True

    @property
    def state_uploads(self):
        '''Host -> device copies of the state so far (tfb_upload_count).'''
        return int(_lib.lib().tfb_upload_count(self._ctx))

    def rhs(self, state):
        '''F(x); replaces Discretization.rhs (Discretization.py:367-390).'''
        self._sync_params()
        state = as_f64(state)
        out = self._result_vector()
        L = _lib.lib()
        self._ensure_state(state)
        check(L.tfb_assemble_resident(self._ctx, None, 0, 1))
        check(L.tfb_rhs_download(self._ctx, ptr(out)))
        return out

    def jacobian(self, state):
        '''J(x) as a DeviceMatrix; replaces Discretization.jacobian (:392-415).'''
        self._sync_params()
        state = as_f64(state)
        mat = DeviceMatrix(self)
        L = _lib.lib()
        if self._resident_key is not None and not self._ensure_state(state):
            check(L.tfb_assemble_resident(self._ctx, mat._h, 1, 0))       # the state is already in HBM
            check(L.tfb_sync(self._ctx))
        else:
            check(L.tfb_jacobian(self._ctx, ptr(state), mat._h, None))    # (pipelined) upload + assembly
            self._resident_key = None                                     # uploaded by the library: checksum not taken
        return mat

    def jacobian_rhs(self, state):
        '''Fused J(x), F(x) in one kernel launch (the state is staged once).'''
        self._sync_params()
        state = as_f64(state)
        mat = DeviceMatrix(self)
        out = self._result_vector()
        check(_lib.lib().tfb_jacobian(self._ctx, ptr(state), mat._h, ptr(out)))
        self._resident_key = None
        return mat, out

    def jacobian_rhs_into(self, state, mat, out):
        '''Fused J(x), F(x) into an existing DeviceMatrix / host buffer (no allocations).'''
        self._sync_params()
        check(_lib.lib().tfb_jacobian(self._ctx, ptr(state), mat._h, ptr(out)))
        self._resident_key = None
        mat._host = None
        return mat, out

    def _mass_diagonal(self):
        if getattr(self, '_mass_diag', None) is None:
            diag = numpy.empty(self.n_local)
            check(_lib.lib().tfb_mass_diag(self._ctx, ptr(diag)))
            self._mass_diag = diag
        return self._mass_diag

    def mass_matrix(self):
        '''M as scipy csc (diagonal; pressure rows empty); replaces Discretization.mass_matrix
        (:417-437) + SciPy.Interface.mass_matrix (SciPy.py:47-49).'''
        from scipy import sparse
        self._require_whole_grid('mass_matrix')
        diag = numpy.empty(self.n)
        check(_lib.lib().tfb_mass_diag(self._ctx, ptr(diag)))
        keep = numpy.abs(diag) > 1e-14                      # Discretization.py:541
        rows = numpy.nonzero(keep)[0]
        begA = numpy.zeros(self.n + 1, dtype=numpy.int64)
        begA[1:] = numpy.cumsum(keep)
        return sparse.csr_matrix((diag[keep], rows, begA), (self.n, self.n)).tocsc()

    # ---- linear solve (SciPy.py:204-315) ----
    def solve(self, jac, rhs, rhs2=None, V=None, W=None, C=None):
        '''Solve ``J y = rhs`` (pressure pinned at row ``dim`` when dof > dim, SciPy.py:212-216)
        with the preconditioned Krylov solver on the device (IDR(s) / FGMRES / BiCGStab, see ``_solve_pinned``).  With a border (``rhs2, V, W, C``) the
        bordered system is reduced to two solves with J and a 1x1 Schur complement.'''
        if not isinstance(jac, (DeviceMatrix, ComplexDeviceMatrix)):
            # host matrices built from ours (TimeIntegration's J - M/(theta dt), real shifts):
            # re-upload onto the structural pattern, cached on the matrix object like `jac.lu`
            dev = getattr(jac, '_tfb_device', None)
            if dev is None:
                dev = DeviceMatrix.from_scipy(self, jac)
                try:
                    jac._tfb_device = dev
                except AttributeError:
                    pass
            jac = dev
        if V is not None:
            return self._bordered_solve(jac, rhs, rhs2, V, W, C)
        if isinstance(jac, ComplexDeviceMatrix):
            return self._solve_complex(jac, rhs)
        return self._solve1(jac, rhs)

    def _solve1(self, jac, rhs):
        if numpy.iscomplexobj(rhs):
            # real matrix, complex right-hand side (eigen-solver glue): the real and imaginary parts are
            # solved separately, exactly what SciPy.Interface._lu_solve does (SciPy.py:194-202)
            rhs = numpy.asarray(rhs)
            y = numpy.empty(rhs.shape, dtype=numpy.complex128)
            y.real = self._solve1(jac, numpy.ascontiguousarray(rhs.real))
            first = self.last_solve
            y.imag = self._solve1(jac, numpy.ascontiguousarray(rhs.imag))
            self.last_solve = dict(self.last_solve, iterations=first['iterations'] + self.last_solve['iterations'],
                                   converged=first['converged'] and self.last_solve['converged'],
                                   relres=max(first['relres'], self.last_solve['relres']),
                                   solve_ms=first['solve_ms'] + self.last_solve['solve_ms'])
            return y
        self._sync_solver()
        b = as_f64(rhs)
        prow, pin_local, pin_saved = -1, -1, 0.0
        if self.dof > self.dim:
            prow = self.pressure_row            # global row; lives on the slab that owns cell 0
            if self.row0 <= prow < self.row0 + self.n_local:
                # rhs[dim] = 0 (SciPy.py:216) without a copy of the vector: the caller's entry is put back after the solve
                pin_local = prow - self.row0
                if b is rhs and not b.flags.writeable:
                    b = b.copy()
                pin_saved = float(b[pin_local])
        try:
            if pin_local >= 0:
                b[pin_local] = 0
            return self._solve_pinned(jac, b, prow)
        finally:
            if pin_local >= 0:
                b[pin_local] = pin_saved

    def _solve_pinned(self, jac, b, prow):
        its = self.parameters.get('Iterative Solver', {})
        o = _lib.TfbSolveOpts()
        o.tol = its.get('Convergence Tolerance', 1e-10)
        o.maxit = its.get('Maximum Iterations', 1000)
        # 180 GB of HBM: keep the whole Krylov space whenever it fits (no restart), the basis
        # needs 16 bytes per unknown and iteration
        fit = max(20, int(100e9 // (16 * self.n_local)))
        o.restart = min(its.get('Restart', 500), fit)
        o.pressure_row = prow
        o.precond = its.get('Preconditioner Id', 0)
        o.verbose = int(bool(self.parameters.get('Verbose', False)))
        # 'Basis Precision': 'single' stores the Krylov basis in fp32 (compressed-basis GMRES, all
        # arithmetic fp64, cycles restart from the true residual); default fp64
        o.basis_fp32 = int(its.get('Basis Precision', 'double') == 'single')
        # 'Method': 'FGMRES' | 'IDR' | 'BiCGStab'.  Default ('auto'): IDR(8) for large true 3-D grids with a fixed
        # preconditioner (the orthogonalisation against an un-restarted basis is 45 % of an FGMRES solve at 128^3),
        # FGMRES otherwise and as the fallback whenever IDR does not reach the tolerance
        method = str(its.get('Method', 'auto')).lower()
        # 'Method': 'Direct' (2-D grids; the default there): block-tridiagonal elimination over the grid lines with pivoted
        # dense line inverses on the device, the counterpart of the reference's SuperLU solve.  Factors are cached on the
        # matrix, so the second solve of a corrector step only substitutes.  Falls back to the Krylov solver when the
        # line inverses do not fit or a line block is singular.
        line = self.dof * self.nx
        direct_ok = self.nz == 1 and self.slab == (0, self.nz) and 8.0 * line * line * (self.ny + 2) < 40e9
        if method == 'direct' and not direct_ok:
            raise ValueError("'Method': 'Direct' is for single-GPU 2-D grids whose line inverses fit in device memory")
        if method == 'direct' or (method == 'auto' and direct_ok):
            y = self._direct_solve(jac, b, prow)
            if y is not None:
                return y
            if method == 'direct':
                raise RuntimeError('direct solve failed: ' + str(self.last_solve))
            method = 'auto'
        o.method = _lib.METHOD_BICGSTAB if method == 'bicgstab' else _lib.METHOD_FGMRES
        o.stall_cycles = int(its.get('Stagnation Cycles', 0))
        # 'Scalar Coupling': 'joint' (default where available: 3-D Rayleigh-Benard) solves w and T together and
        # iterates on the (velocity, temperature) block; 'none' is the block-triangular preconditioner
        joint = getattr(self, '_joint', False) and str(its.get('Scalar Coupling', 'joint')).lower() != 'none'
        inner = int(its.get('Velocity Iterations', 8 if joint else 0))
        # 'Preconditioner Precision': 'double' | 'single' (fp32 FDM sub-solves) | 'tf32x3' (x/y transforms of the FDM
        # solves on the tensor cores with a 3xTF32 split -- fp32 accuracy -- and Thomas sweeps along z).  The
        # preconditioner only steers the iteration: convergence is always decided on the true fp64 residual.
        pprec = str(its.get('Preconditioner Precision', 'auto')).lower()
        if pprec not in ('auto', 'double', 'single', 'tf32x3'):
            raise ValueError("'Preconditioner Precision' must be 'double', 'single' or 'tf32x3'")
        tensor_ok = self.dim == 3 and self.nz > 1 and self.nx <= 128 and self.ny <= 128
        if pprec == 'tf32x3' and not tensor_ok:
            raise ValueError("'Preconditioner Precision': 'tf32x3' needs a 3-D grid with nx, ny <= 128")
        # 'Schur Complement': 'LSC' (least-squares commutator, two Poisson solves and a product with the velocity block)
        # or 'Scaled Mass' (dp = gamma r_p / cell volume, gamma estimated per matrix from the velocity block's spectrum)
        # Default ('auto'): the scaled mass matrix wherever IDR is chosen automatically for a problem without scalars
        # (measured, 128^3 cavity at Re = 100: 155 products / 275 ms against 199 / 505 ms with LSC and the same Newton
        # iterates to 6e-15; 64^3 at Re = 400: 520 products / 165 ms where IDR with LSC does not converge), LSC otherwise
        schur = str(its.get('Schur Complement', 'auto')).lower()
        if schur not in ('auto', 'lsc', 'scaled mass'):
            raise ValueError("'Schur Complement' must be 'LSC' or 'Scaled Mass'")
        auto = method == 'auto'
        big3d = self.dim == 3 and self.nz > 1 and self.n >= self.AUTO_IDR_MIN_UNKNOWNS   # global size: same choice on every rank
        if auto:
            method = 'idr' if (big3d and inner == 0 and 'Basis Precision' not in its and pprec in ('auto', 'double', 'tf32x3')) else 'fgmres'
        auto_schur = schur == 'auto'
        if auto_schur:
            schur = 'scaled mass' if (auto and method == 'idr' and self.dof == self.dim + 1) else 'lsc'
        if pprec == 'auto':
            # large 3-D grids: every FDM sub-solve on the tensor-core path (fp32 storage); small grids keep fp64
            pprec = 'tf32x3' if (big3d and tensor_ok) else 'double'
        flags = (_lib.PREC_FP32 if pprec == 'single' else 0) | (_lib.PREC_TENSOR if pprec == 'tf32x3' else 0) \
            | (0 if joint else _lib.PREC_NO_JOINT) | (_lib.PREC_SCALED_MASS if schur == 'scaled mass' else 0)
        o.precond_flags = flags
        o.inner_its = min(24, max(0, inner))
        if method.startswith('idr'):
            # IDR(s): short recurrences instead of a Krylov basis; needs a fixed preconditioner, so the variants with
            # inner iterations keep FGMRES
            if inner > 0:
                raise ValueError("'Method': 'IDR' cannot be combined with inner iterations ('Velocity Iterations', "
                                 "coupled scalar solve); use FGMRES or 'Scalar Coupling': 'none'")
            o.method = _lib.METHOD_IDR
            o.idr_s = max(1, min(16, int(its.get('IDR Dimension', 8))))
        info = _lib.TfbSolveInfo()
        y = self._result_vector()          # tfb_solve writes every entry (the solvers start from x = 0 on the device)
        rc = check(_lib.lib().tfb_solve(jac._h, ptr(b), ptr(y), ctypes.byref(o), ctypes.byref(info)))
        spent_its, spent_ms = 0, 0.0
        if rc != 0 and auto and method == 'idr':
            # IDR stagnated or broke down: the un-restarted FGMRES is the robust path
            self._debug_print('IDR: relres %.3e after %d products, falling back to FGMRES' % (info.relres, info.iters))
            spent_its, spent_ms = info.iters, info.solve_ms
            method, o.method = 'fgmres', _lib.METHOD_FGMRES
            if auto_schur:
                schur = 'lsc'
                o.precond_flags &= ~_lib.PREC_SCALED_MASS
            rc = check(_lib.lib().tfb_solve(jac._h, ptr(b), ptr(y), ctypes.byref(o), ctypes.byref(info)))
        self.last_solve = {'iterations': info.iters + spent_its, 'relres': info.relres, 'converged': rc == 0,
                           'setup_ms': info.setup_ms, 'solve_ms': info.solve_ms + spent_ms,
                           'method': 'IDR' if method.startswith('idr') else ('BiCGStab' if method == 'bicgstab' else 'FGMRES'),
                           'schur': 'Scaled Mass' if schur == 'scaled mass' else 'LSC', 'precond_precision': pprec}
        if rc != 0:
            # the reference's direct solve cannot fail silently; an unconverged Krylov solve must not either
            import warnings
            warnings.warn('B200 solve: %s stopped at relative residual %.3e after %d iterations (tolerance %.1e)'
                          % (self.last_solve['method'], info.relres, self.last_solve['iterations'], o.tol), RuntimeWarning)
        self._debug_print('%s: %d iterations, relres %.3e' % (self.last_solve['method'], info.iters, info.relres))
        return y

    def _solve_complex(self, mat, rhs):
        '''(A_r + i A_i) y = rhs with the pressure pinned like the real solve: right-preconditioned flexible GMRES in
        complex arithmetic on the host, every operator product (four real device SpMVs) and every preconditioner
        application (the block preconditioner of A_r on the real and imaginary parts) on the device.  The counterpart of
        SuperLU on a complex matrix in the reference (SciPy.py:131-162,194-202).'''
        self._sync_solver()
        its = self.parameters.get('Iterative Solver', {})
        tol = its.get('Convergence Tolerance', 1e-10)
        maxit = int(its.get('Maximum Iterations', 1000))
        restart = int(min(its.get('Restart', 100), 200))
        b = numpy.array(rhs, dtype=numpy.complex128)
        prow = self.pressure_row if self.dof > self.dim else -1
        if prow >= 0:
            b[prow] = 0
        o = _lib.TfbSolveOpts()
        o.pressure_row = prow
        o.precond_flags = 0 if getattr(self, '_joint', False) else _lib.PREC_NO_JOINT
        L = _lib.lib()

        def op(z):
            z = z.copy()
            zp = z[prow] if prow >= 0 else 0.0
            if prow >= 0:
                z[prow] = 0                     # dropped column
            y = mat @ z
            if prow >= 0:
                y[prow] = -zp                   # pinned row: -1 on the diagonal
            return y

        def prec(r):
            out = numpy.empty(self.n_local, dtype=numpy.complex128)
            for part in ('real', 'imag'):
                src = numpy.ascontiguousarray(getattr(r, part))
                dst = numpy.empty(self.n_local)
                check(L.tfb_precond_apply_opts(mat.real._h, ptr(src), ptr(dst), ctypes.byref(o)))
                setattr(out, part, dst)
            return out

        bnorm = numpy.linalg.norm(b)
        y = numpy.zeros(self.n_local, dtype=numpy.complex128)
        total, relres = 0, 1.0
        if bnorm == 0.0:
            relres = 0.0
        while total < maxit and relres > tol:
            r = b - op(y) if total else b.copy()
            beta = numpy.linalg.norm(r)
            relres = beta / bnorm
            if relres <= tol:
                break
            m = min(restart, maxit - total)
            V = numpy.zeros((m + 1, self.n_local), dtype=numpy.complex128)
            Z = numpy.zeros((m, self.n_local), dtype=numpy.complex128)
            H = numpy.zeros((m + 1, m), dtype=numpy.complex128)
            V[0] = r / beta
            k = 0
            for j in range(m):
                Z[j] = prec(V[j])
                w = op(Z[j])
                for _ in range(2):                      # classical Gram-Schmidt, two sweeps
                    h = V[:j + 1].conj() @ w
                    w = w - h @ V[:j + 1]
                    H[:j + 1, j] += h
                H[j + 1, j] = numpy.linalg.norm(w)
                total += 1
                k = j + 1
                e1 = numpy.zeros(k + 1, dtype=numpy.complex128)
                e1[0] = beta
                coef, res, _, _ = numpy.linalg.lstsq(H[:k + 1, :k], e1, rcond=None)
                est = numpy.linalg.norm(H[:k + 1, :k] @ coef - e1) / bnorm
                if est <= tol or H[j + 1, j] == 0 or total >= maxit:
                    break
                V[j + 1] = w / H[j + 1, j]
            y = y + coef @ Z[:k]
            relres = numpy.linalg.norm(b - op(y)) / bnorm
        self.last_solve = {'iterations': total, 'relres': float(relres), 'converged': bool(relres <= tol * 1.0001), 'setup_ms': 0.0,
                           'solve_ms': 0.0, 'method': 'complex FGMRES', 'schur': 'LSC', 'precond_precision': 'double'}
        if not self.last_solve['converged']:
            import warnings
            warnings.warn('B200 complex solve stopped at relative residual %.3e after %d iterations' % (relres, total), RuntimeWarning)
        return y

    def _direct_solve(self, jac, b, prow):
        info = _lib.TfbSolveInfo()
        y = self._result_vector()
        L = _lib.lib()
        rc = L.tfb_direct_solve(jac._h, ptr(b), ptr(y), ctypes.c_int(prow), ctypes.byref(info))
        if rc != 0:
            self.last_solve = {'iterations': 0, 'relres': float(info.relres) if rc > 0 else float('nan'), 'converged': False,
                               'setup_ms': 0.0, 'solve_ms': 0.0, 'method': 'Direct', 'schur': '-', 'precond_precision': '-',
                               'error': L.tfb_last_error().decode() if rc < 0 else 'residual %.2e' % info.relres}
            self._debug_print('direct solve not usable (%s), using the Krylov solver' % self.last_solve['error'])
            return None
        self.last_solve = {'iterations': 1, 'relres': info.relres, 'converged': True, 'setup_ms': info.setup_ms,
                           'solve_ms': info.solve_ms, 'method': 'Direct', 'schur': '-', 'precond_precision': '-'}
        self._debug_print('Direct: factor %.1f ms, solve %.1f ms, relres %.3e' % (info.setup_ms, info.solve_ms, info.relres))
        return y

    def _bordered_solve(self, jac, rhs, rhs2, V, W, C):
        # [J V; W^T C] [y1; y2] = [rhs; rhs2]  ->  block elimination with two solves with J
        self._require_whole_grid('the bordered solve')      # W @ a would need an all-reduce over the slabs
        if W is None:
            W = V
        if C is None:
            C = 0.0
        a = self._solve1(jac, rhs)
        b = self._solve1(jac, as_f64(V))
        y2 = (rhs2 - W @ a) / (C - W @ b)
        y1 = a - b * y2
        return y1, (y2.item() if numpy.ndim(y2) else y2)

    def eigs(self, state, return_eigenvectors=False, enable_recycling=False):
        '''Generalized eigenvalues of ``J(state) v = lambda M v`` closest to
        ``parameters['Eigenvalue Solver']['Target']`` (default 0), sorted by descending real part
        -- the contract of BaseInterface.eigs / _eigs (BaseInterface.py:294-386).  The reference
        runs jadapy's JDQZ there; this backend runs a shift-and-invert Arnoldi process whose
        operator ``(J - sigma M)^-1 M`` is one device Krylov solve per step (transiflow_b200/eigs.py).
        Complex targets use the complex solve of ``_solve_complex``.'''
        from .eigs import shift_invert_arnoldi
        self._require_whole_grid('eigs')
        prm = self.parameters.get('Eigenvalue Solver', {})
        target = prm.get('Target', 0.0)
        target = complex(target) if (numpy.iscomplexobj(target) and complex(target).imag != 0.0) else float(numpy.real(target))
        num = int(prm.get('Number of Eigenvalues', 5))
        tol = float(prm.get('Tolerance', 1e-7))
        max_dim = int(prm.get('Maximum Subspace Dimension', 60))
        recycle = prm.get('Recycle Subspaces', enable_recycling)
        jac = self.jacobian(state)
        mass = self.mass_matrix()
        if isinstance(target, complex):
            # complex shift: J - sigma M as real and imaginary parts on the device (ComplexDeviceMatrix)
            shifted = DeviceMatrix.from_scipy(self, jac.tocsc() - target * mass)
        else:
            shifted = jac if target == 0.0 else jac - target * mass
        failures = []

        def apply_op(v):
            y = self.solve(shifted, mass @ v)
            if not self.last_solve['converged']:
                failures.append(self.last_solve['relres'])
            return y

        v0 = self._eig_start if (recycle and getattr(self, '_eig_start', None) is not None) else None
        lam, vec, ok = shift_invert_arnoldi(apply_op, self.n_local, num=num, target=target, tol=tol,
                                            max_dim=max_dim, v0=v0)
        if failures or not ok:
            import warnings
            warnings.warn('eigs: %d inner solves did not converge; Arnoldi converged: %s' % (len(failures), ok))
        self._eig_start = numpy.real(vec.sum(axis=1)) if recycle else None
        if return_eigenvectors:
            return lam, vec
        return lam

    # ---- save / load: the file conventions of BaseInterface.py:118-215 ----
    def save_json(self, name, obj):
        with open(name, 'w') as f:
            json.dump(obj, f, cls=ParameterEncoder)

    def load_json(self, name):
        with open(name, 'r') as f:
            return json.load(f, object_hook=parameter_decoder)

    def save_parameters(self, name):
        '''Parameter set -> ``name + '.params'``.'''
        params_name = name + '.params'
        self.save_json(params_name, self.parameters)
        print('Wrote parameters to', params_name, flush=True)

    def load_parameters(self, name):
        '''``name + '.params'`` -> the (shared) parameter dict, reporting what changed.'''
        params_name = name + '.params'
        before = self.parameters.copy()
        self.parameters.update(self.load_json(params_name))
        print('Read parameters from', params_name, flush=True)
        for key, value in self.parameters.items():
            if value != before.get(key):
                print("Updated '{}' from {} to {}".format(key, before.get(key), value), flush=True)

    def save_state(self, name, x):
        self.save_parameters(name)
        if not name.endswith('.npy'):
            name += '.npy'
        numpy.save(name, x)
        print('Wrote state to', name, flush=True)

    def load_state(self, name):
        self.load_parameters(name)
        if not name.endswith('.npy'):
            name += '.npy'
        x = numpy.load(name)
        print('Read state from', name, flush=True)
        return x
