'''Host-side preparation for the B200 assembly kernels (numpy only).

Everything that involves transcendentals (grid stretching, wind / AMOC forcing profiles,
sqrt of the Grashof number) or the mutable parameter dictionary is evaluated here, on the
host, with the same numpy expressions the reference uses, and handed to the kernels as
1-D metric arrays and a small struct of scalars.  File:line citations are to
/root/reference/transiflow/.
'''
import ctypes

import numpy

from . import recipes

TFB_MAX_FORCE = 8
TFB_NMET = 8


class TfbParams(ctypes.Structure):
    '''Mirror of ``struct TfbParams`` (csrc/tfb_rows_common.h).'''
    _fields_ = [
        ('c_visc', ctypes.c_double), ('c_T', ctypes.c_double), ('c_S', ctypes.c_double),
        ('c_pert', ctypes.c_double), ('beta', ctypes.c_double),
        ('bc_cf', ctypes.c_double * TFB_MAX_FORCE), ('bc_ca', ctypes.c_double * TFB_MAX_FORCE),
        ('nl', ctypes.c_int), ('has_beta', ctypes.c_int), ('pert', ctypes.c_int), ('pad_', ctypes.c_int),
    ]


# ---------------------------------------------------------------------------------------
# coordinate vectors (utils.py:175-267): length n+3, x[i] = east face of cell i, the two
# trailing entries are the faces at and before the domain start (Python wrap-around).
# ---------------------------------------------------------------------------------------

def uniform_vector(start, end, n):
    h = (end - start) / n
    faces = start + numpy.arange(-1, n + 2) * h
    return numpy.roll(faces, -2)


def _mirror_ghost_cells(x, start, end):
    h = x[0] - x[-1]
    if start == 0:
        x[-2] = x[-1] - h
    if end == 1:
        x[-3] = x[-4] + h
    return x


def stretched_vector(start, end, n, sigma, method='tanh'):
    x = uniform_vector(0, 1, n)
    if method == 'sin':
        x = x - sigma * numpy.sin(2 * numpy.pi * x)
    else:
        x = 0.5 * (1 + numpy.tanh(2 * sigma * (x - 0.5)) / numpy.tanh(sigma))
    x = start + x * (end - start)
    return _mirror_ghost_cells(x, start, end)


def coordinate_vector(parameters, start, end, n):
    '''Discretization.get_coordinate_vector (Discretization.py:186-208).'''
    if parameters.get('Grid Stretching', False) or 'Grid Stretching Factor' in parameters:
        if parameters.get('Grid Stretching Method', 'tanh') == 'sin':
            return stretched_vector(start, end, n, parameters.get('Grid Stretching Factor', 0.1), 'sin')
        return stretched_vector(start, end, n, parameters.get('Grid Stretching Factor', 1.5), 'tanh')
    return uniform_vector(start, end, n)


def cell_centers(vec):
    '''utils.compute_coordinate_vector_centers (utils.py:269-288): centre of cell i at [i],
    ghost cell before the domain at [-1].'''
    n = len(vec) - 1
    idx = numpy.arange(-1, n - 1)
    out = numpy.zeros(n)
    out[idx] = (vec[idx] + vec[idx - 1]) / 2
    return out


def axis_metrics(X, n):
    '''The TFB_NMET 1-D arrays of one axis, shape (8, n), C order:
    hc, hu, 1/hc, 1/hp, 1/hm, 1/hu, wm, wp with Python wrap-around indexing of X.'''
    X = numpy.asarray(X, dtype=numpy.float64)
    i = numpy.arange(n)
    hc = X[i] - X[i - 1]                    # cell width                (Discretization.py:743)
    hp = X[i + 1] - X[i]                    # next cell width           (:745)
    hu = (X[i + 1] - X[i - 1]) / 2          # staggered width           (:784)
    hm = (X[i] - X[i - 2]) / 2              # centre distance i-1 -> i  (:780)
    wm = 1 / 2 * hc / hu                    # _weighted_average         (:1220)
    wp = 1 / 2 * hp / hu                    #                           (:1221)
    return numpy.ascontiguousarray(numpy.stack([hc, hu, 1 / hc, 1 / hp, 1 / hm, 1 / hu, wm, wp]))


def coriolis_metrics(Y, ny):
    j = numpy.arange(ny)
    Y = numpy.asarray(Y, dtype=numpy.float64)
    return numpy.ascontiguousarray(numpy.stack([Y[j] / 2, -(Y[j] + Y[j - 1]) / 4]))


def _get(parameters, name, default=0):
    '''Discretization.get_parameter (Discretization.py:164-184).'''
    return parameters.get(name, default) if name in parameters else default


def _robin_constants(X, m, far, Q, Bi):
    '''heat_flux_<face> (BoundaryConditions.py:339-421).'''
    if far:
        h = (X[m] - X[m - 2]) / 2
        return h * Q / (1 + h * Bi / 2), (1 - h * Bi / 2) / (1 + h * Bi / 2)
    h = (X[0] - X[-2]) / 2
    return -h * Q / (1 - h * Bi / 2), (1 + h * Bi / 2) / (1 - h * Bi / 2)


def wind_stress(parameters, nx, ny, nz, dof, x, y, z):
    '''Discretization.wind_stress (Discretization.py:1101-1118) as a state-ordered vector.'''
    alpha = _get(parameters, 'Wind Stress Parameter')
    asym = _get(parameters, 'Asymmetry Parameter')
    i, j, k = numpy.arange(nx - 1), numpy.arange(ny), numpy.arange(nz)
    dx = ((x[i + 1] - x[i - 1]) / 2)[:, None, None]
    dy = (y[j] - y[j - 1])[None, :, None]
    dz = (z[k] - z[k - 1])[None, None, :]
    yc = ((y[j] + y[j - 1]) / 2)[None, :, None]
    val = - (1 - asym) * numpy.cos(2 * numpy.pi * yc) - asym * numpy.cos(numpy.pi * yc)
    val = val * (alpha / (2 * numpy.pi) * dx * dy * dz)
    frc = numpy.zeros((nz, ny, nx, dof))
    frc[:, :, :nx - 1, 0] = numpy.transpose(val, (2, 1, 0))
    return frc.ravel()


def amoc_face_values(parameters, nx, ny, nz, x, y):
    '''Value arrays of temperature_north(theta*T_S) and salinity_flux_north(sigma*Q_S)
    (Discretization.py:670-686; BoundaryConditions.py:316,445,472), reduced to the in-plane
    centre entries the kernels need: shape (nz, nx), first in-plane axis (x) fastest.'''
    xc = cell_centers(x)
    theta = _get(parameters, 'Temperature Forcing')
    asym = _get(parameters, 'Asymmetry Parameter')
    A = parameters.get('X-max', 1.0)
    T_S = numpy.zeros((nx + 2, nz + 2))
    T_S[:, 0] = 1 / 2 * ((1 - asym) * numpy.cos(2 * numpy.pi * (xc / A - 1 / 2))
                         + asym * numpy.cos(numpy.pi * xc / A) + 1)
    tval = numpy.ones((nx + 2, nz + 2)) * (2 * (theta * T_S))
    sigma = _get(parameters, 'Freshwater Flux')
    p = 2
    Q_S = numpy.zeros((nx + 2, nz + 2))
    Q_S[:, 0] = 3 * numpy.cos(p * numpy.pi * (xc / A - 1 / 2)) - 6 / (p * numpy.pi) * numpy.sin(p * numpy.pi / 2)
    h = (y[ny] - y[ny - 2]) / 2
    sval = numpy.ones((nx + 2, nz + 2)) * (h * (sigma * Q_S))
    return (numpy.ascontiguousarray(tval[:nx, :nz].T), numpy.ascontiguousarray(sval[:nx, :nz].T))


def make_params(cfg, problem, parameters, nx, ny, nz, x, y, z):
    '''Evaluate the per-call scalars from the (mutable, shared) parameter dict.  Returns
    (TfbParams, {force op index: face value array}).'''
    prm = TfbParams()
    Re = _get(parameters, 'Reynolds Number', 1.0)       # Discretization.py:236,278
    Ra = _get(parameters, 'Rayleigh Number', 1.0)
    Pr = _get(parameters, 'Prandtl Number', 1.0)
    Gr = _get(parameters, 'Grashof Number', Ra / Pr)
    Le = _get(parameters, 'Lewis Number', 1.0)
    if Re == 0:
        Re = 1
    if Gr == 0:
        Gr = 1 / Pr
    prm.c_visc = 1 / (Re * numpy.sqrt(Gr))
    prm.c_T = 1 / (Pr * numpy.sqrt(Gr))
    prm.c_S = 1 / (Le * Pr * numpy.sqrt(Gr))
    Bi = _get(parameters, 'Biot Number')
    prm.pert = int(problem == recipes.RBP)
    prm.c_pert = Bi / (Bi + 1) if prm.pert else 0.0
    beta = _get(parameters, 'Rossby Parameter')
    prm.beta = beta
    prm.has_beta = int(bool(beta) and cfg.dim == 2)
    Re_nl = _get(parameters, 'Reynolds Number')         # default 0 here, Discretization.py:333
    prm.nl = int(not (Re_nl == 0 and not cfg.dof > cfg.dim + 1))
    X = (x, y, z)
    m = (nx, ny, nz)
    arrays = {}
    fidx = 0
    for op in cfg.recipe:
        if op[0] != 'force':
            continue
        _, axis, far, var, kind, arg = op
        if kind == 'lid':
            cf, ca = 2 * _get(parameters, 'Lid Velocity', 1), -1
        elif kind == 'lidv':                                  # a user callback's moving lid: literal velocity
            cf, ca = 2 * arg, -1
        elif kind == 'temp':
            Tb = (1 if problem == recipes.RB else 0) if arg == 'bottom' else arg
            cf, ca = 2 * Tb, -1
        elif kind == 'hflux':
            Q = _get(parameters, 'Asymmetry Parameter') if arg[0] == 'asym' else arg[0]
            b = Bi if arg[1] == 'Bi' else float(arg[1])
            cf, ca = _robin_constants(X[axis], m[axis], far, Q, b)
        elif kind == 'sflux':
            Xa = X[axis]
            h = (Xa[m[axis]] - Xa[m[axis] - 2]) / 2 if far else (Xa[0] - Xa[-2]) / 2
            cf, ca = (h * arg, 1) if far else (-h * arg, 1)
        elif kind in ('tarr', 'sarr'):
            if 'amoc' not in arrays:
                arrays['amoc'] = amoc_face_values(parameters, nx, ny, nz, x, y)
            arrays[fidx] = arrays['amoc'][0 if kind == 'tarr' else 1]
            cf, ca = 0.0, (-1 if kind == 'tarr' else 1)
        else:
            raise ValueError(kind)
        prm.bc_cf[fidx], prm.bc_ca[fidx] = cf, ca
        fidx += 1
    arrays.pop('amoc', None)
    return prm, arrays


# ---------------------------------------------------------------------------------------
# Fast-diagonalisation (FDM) data for the block preconditioner of the linear solver.
# The diffusion part of every variable is a sum of Kronecker products of 1-D stencils
# (Discretization.py:740-865 with the wall folds of BoundaryConditions.py:55-233,471-544):
#     Op_v = coef_v * sum_a  K_a (x) prod_{b != a} M_b
# with K_a symmetric tridiagonal and M_b diagonal.  With K q = lambda M q, Q^T M Q = I the
# inverse is three dense transforms, a scaling by 1/(coef*(lx+ly+lz)) and three transforms back.
# ---------------------------------------------------------------------------------------

def fold_coefficients(cfg, prm):
    '''(variable, axis, far) -> coefficient with which the recipe folds the ghost value into the
    wall cell: -1 no-slip / Dirichlet, +1 free-slip / zero flux, Robin constants for heat flux.'''
    out = {}
    fidx = 0
    for op in cfg.recipe:
        if op[0] == 'wall':
            _, axis, far, sign = op
            for v in range(cfg.dof):
                if v != cfg.p:
                    out.setdefault((v, axis, far), float(sign))
        elif op[0] == 'force':
            _, axis, far, var, kind, arg = op
            out[(cfg.var(var), axis, far)] = float(prm.bc_ca[fidx])
            fidx += 1
    return out


def _pencil_km(kind, met, n, s_near, s_far):
    '''Dense symmetric tridiagonal stiffness K and diagonal mass M of one variable along one axis.'''
    hc, hu, rhc, rhp, rhm, rhu = met[0], met[1], met[2], met[3], met[4], met[5]
    if kind == 'own':       # velocity along its own axis (_u_xx): faces 0..n-2, the wall face is not an unknown
        m = n - 1
        lower, upper, M = rhc[:m], rhp[:m], hu[:m]
        diag = -(lower + upper)
    else:                   # staggered direction of a velocity (_u_yy/_u_zz) or a cell-centred scalar (_C_xx)
        m = n
        lower, upper, M = rhm[:m], rhu[:m], hc[:m]
        diag = -(lower + upper)
        diag[0] += s_near * lower[0]
        diag[m - 1] += s_far * upper[m - 1]
    K = numpy.diag(diag)
    if m > 1:
        K += numpy.diag(upper[:m - 1], 1) + numpy.diag(lower[1:], -1)
    K = (K + K.T) / 2       # symmetric by construction (1/hp[i] == 1/hc[i+1], 1/hu[i] == 1/hm[i+1])
    return K, numpy.array(M, dtype=numpy.float64)


def _pencil(kind, met, n, s_near, s_far):
    K, M = _pencil_km(kind, met, n, s_near, s_far)
    ms = 1 / numpy.sqrt(M)
    lam, Y = numpy.linalg.eigh(ms[:, None] * K * ms[None, :])
    return numpy.ascontiguousarray(ms[:, None] * Y), numpy.ascontiguousarray(lam)


def joint_z_operators(cfg, prm, mets, nz):
    '''(8, nz) table for tfb_joint_set: the tridiagonal z-stencils (lower, diagonal, upper) and masses of
    the vertical velocity (nz-1 faces, zero padded) and of the temperature, rows ordered as the
    TFB_JZ_* enum of csrc/tfb_joint.h.  The coupled (w, T) line solve of the Rayleigh-Benard
    preconditioner is built from these and from the vertical couplings read off the Jacobian.'''
    folds = fold_coefficients(cfg, prm)
    w = cfg.dim - 1
    Kw, Mw = _pencil_km('own', mets[2], nz, 0.0, 0.0)
    KT, MT = _pencil_km('cen', mets[2], nz, folds.get((cfg.T, 2, 0), 0.0), folds.get((cfg.T, 2, 1), 0.0))
    out = numpy.zeros((8, nz))
    m = nz - 1
    out[0, 1:m] = numpy.diag(Kw, -1)
    out[1, :m] = numpy.diag(Kw)
    out[2, :m - 1] = numpy.diag(Kw, 1)
    out[3, :m] = Mw
    out[4, 1:] = numpy.diag(KT, -1)
    out[5] = numpy.diag(KT)
    out[6, :nz - 1] = numpy.diag(KT, 1)
    out[7] = MT
    assert w == 2
    return numpy.ascontiguousarray(out)


def fdm_operators(cfg, prm, mets, nx, ny, nz, pencils=None):
    '''[(var, axis, m, Q, lam, coef)] for tfb_fdm_set.  If `pencils` is a list it receives
    (var, axis, m, lower, diag, upper, mass) of the same 1-D stencils for tfb_fdm_set_pencil.'''
    folds = fold_coefficients(cfg, prm)
    n = (nx, ny, nz)
    ndir = 3 if (cfg.dim == 3 and nz > 1) else 2
    out = []
    for v in range(cfg.dof):
        if v == cfg.p:
            coef = -1.0     # Lp = D M^-1 G = -(Neumann Laplacian)
        elif v < cfg.dim:
            coef = prm.c_visc
        elif v == cfg.T:
            coef = prm.c_T
        else:
            coef = prm.c_S
        for a in range(ndir):
            if v == cfg.p:
                args = ('cen', mets[a], n[a], 1.0, 1.0)
            else:
                kind = 'own' if (v < cfg.dim and a == v) else 'cen'
                args = (kind, mets[a], n[a], folds.get((v, a, 0), 0.0), folds.get((v, a, 1), 0.0))
            Q, lam = _pencil(*args)
            out.append((v, a, Q.shape[0], Q, lam, float(coef)))
            if pencils is not None:
                K, M = _pencil_km(*args)
                m = K.shape[0]
                lower, upper = numpy.zeros(m), numpy.zeros(m)
                lower[1:] = numpy.diag(K, -1)
                upper[:m - 1] = numpy.diag(K, 1)
                pencils.append((v, a, m, lower, numpy.ascontiguousarray(numpy.diag(K)), upper, numpy.ascontiguousarray(M)))
    return out
