'''Problem configurations of the B200 assembly path: which equations exist, and the ordered
boundary-condition recipe of every problem type.

This table is the single source for (a) the kernel generator (``codegen/gen.py``), which
unrolls each recipe into predicated straight-line code, and (b) the host shim
(``interface.py``), which evaluates the per-op constants from the parameter dict on every
call.  It restates the recipes of the reference's ``Discretization._lid_driven_cavity`` ...
``_amoc`` (/root/reference/transiflow/Discretization.py:577-702) as data.

Op tuples
---------
``('wall', dir, far, sign)``
    no-slip (sign -1) / free-slip (sign +1) wall; BoundaryConditions.py:55-233.
``('force', dir, far, var, kind, arg)``
    ``_constant_forcing_<face>`` (BoundaryConditions.py:471-544) on variable ``var``
    ('u','v','T','S'); ``kind`` selects how the host computes ``(forcing_constant,
    atom_constant)``: 'lid' (2V,-1), 'temp' (2Tb,-1), 'hflux' (Robin), 'sflux' (h*Q, 1),
    'tarr'/'sarr' (per-face value array, AMOC).  ``arg`` names the parameter(s).
``('pin_s',)``
    AMOC: zero the column of the first salinity unknown, -1 on its diagonal
    (Discretization.py:690-701).
dir: 0=x (west/east), 1=y (south/north), 2=z (bottom/top); far: 1 = east/north/top.
'''

E, W, N, S, T, B = (0, 1), (0, 0), (1, 1), (1, 0), (2, 1), (2, 0)


def wall(face, sign):
    return ('wall', face[0], face[1], sign)


def force(face, var, kind, arg=None):
    return ('force', face[0], face[1], var, kind, arg)


NOSLIP, FREESLIP = -1, +1


def _ldc(flat):
    ops = [wall(E, NOSLIP), wall(W, NOSLIP), wall(S, NOSLIP)]
    if flat:
        return ops + [force(N, 'u', 'lid'), wall(N, NOSLIP)]
    return ops + [wall(N, NOSLIP), wall(B, NOSLIP), force(T, 'u', 'lid'), wall(T, NOSLIP)]


def _rb(flat):
    ops = [force(E, 'T', 'hflux', ('asym', 0)), force(W, 'T', 'hflux', (0, 0)), wall(E, NOSLIP), wall(W, NOSLIP)]
    if flat:
        return ops + [force(N, 'T', 'hflux', (0, 'Bi')), force(S, 'T', 'temp', 'bottom'),
                      wall(N, FREESLIP), wall(S, NOSLIP)]
    return ops + [force(N, 'T', 'hflux', (0, 0)), force(S, 'T', 'hflux', (0, 0)), wall(N, NOSLIP), wall(S, NOSLIP),
                  force(T, 'T', 'hflux', (0, 'Bi')), force(B, 'T', 'temp', 'bottom'),
                  wall(T, FREESLIP), wall(B, NOSLIP)]


def _dhc(flat):
    ops = [force(E, 'T', 'temp', -1 / 2), force(W, 'T', 'temp', 1 / 2), wall(E, NOSLIP), wall(W, NOSLIP),
           force(N, 'T', 'hflux', (0, 0)), force(S, 'T', 'hflux', (0, 0)), wall(N, NOSLIP), wall(S, NOSLIP)]
    if flat:
        return ops
    return ops + [force(T, 'T', 'hflux', (0, 0)), force(B, 'T', 'hflux', (0, 0)), wall(T, NOSLIP), wall(B, NOSLIP)]


def _qg(flat):
    assert flat
    return [wall(E, NOSLIP), wall(W, NOSLIP), wall(N, FREESLIP), wall(S, FREESLIP)]


def _amoc(flat):
    assert flat
    return [force(E, 'T', 'hflux', (0, 0)), force(W, 'T', 'hflux', (0, 0)),
            force(E, 'S', 'sflux', 0), force(W, 'S', 'sflux', 0),
            wall(E, FREESLIP), wall(W, FREESLIP),
            force(S, 'T', 'hflux', (0, 0)), force(S, 'S', 'sflux', 0), wall(S, FREESLIP),
            force(N, 'T', 'tarr'), force(N, 'S', 'sarr'), wall(N, FREESLIP),
            ('pin_s',)]


class Config:
    '''One generated kernel family.  ``flat`` = the grid has nz == 1 (dim == 2, or dim == 3
    "semi-2D" where the three z-offsets fold onto one column, Discretization.py:127-128).'''

    def __init__(self, name, cid, problem, dim, flat, dof, recipe_fn):
        self.name, self.cid, self.problem, self.dim, self.flat, self.dof = name, cid, problem, dim, flat, dof
        self.recipe = recipe_fn(flat)
        self.fold = dim == 3 and flat
        self.has_T = dof > dim + 1
        self.has_S = dof > dim + 2
        self.p = dim
        self.T = dim + 1
        self.S = dim + 2
        self.nforce = sum(1 for op in self.recipe if op[0] == 'force')

    def var(self, v):
        return {'u': 0, 'v': 1, 'w': 2, 'p': self.p, 'T': self.T, 'S': self.S}[v]


# problem ids are shared with the C ABI (include/tfb200.h: TFB_PROBLEM_*)
LDC, RB, RBP, DHC, QG, AMOC = range(6)

CONFIGS = [
    Config('ldc2d', 0, LDC, 2, True, 3, _ldc),
    Config('ldc3d', 1, LDC, 3, False, 4, _ldc),
    Config('rb2d', 2, RB, 2, True, 4, _rb),
    Config('rb3d', 3, RB, 3, False, 5, _rb),
    Config('dhc2d', 4, DHC, 2, True, 4, _dhc),
    Config('dhc3d', 5, DHC, 3, False, 5, _dhc),
    Config('qg2d', 6, QG, 2, True, 3, _qg),
    Config('amoc2d', 7, AMOC, 2, True, 5, _amoc),
    Config('ldc3d_flat', 8, LDC, 3, True, 4, _ldc),
    Config('rb3d_flat', 9, RB, 3, True, 5, _rb),
    Config('dhc3d_flat', 10, DHC, 3, True, 5, _dhc),
]

PROBLEM_IDS = {
    'lid-driven cavity': LDC,
    'rayleigh-benard': RB,
    'rayleigh-benard perturbation': RBP,
    'differentially heated cavity': DHC,
    'double gyre': QG,
    'amoc': AMOC,
}


def find_config(problem, dim, nz, dof):
    '''Config for (problem id, dim, nz, dof) or None if the combination is not generated.
    Rayleigh-Benard Perturbation shares the RB kernels (runtime flag).'''
    base = RB if problem == RBP else problem
    flat = dim == 2 or nz <= 1
    for c in CONFIGS:
        if c.problem == base and c.dim == dim and c.flat == flat and c.dof == dof:
            return c
    return None


# ---------------------------------------------------------------------------------------
# User-supplied boundary conditions (Discretization.py:62-66,719: ``boundary_conditions(bc, atom)``).
# The callback is run once against a recorder that has the public methods of the reference's
# BoundaryConditions class (BoundaryConditions.py:55-469) and notes which op each call is -- every
# one of them is a wall fold or a `_constant_forcing_<face>` with two constants.  The generated
# kernels depend on the ORDER and KIND of the ops (face, variable, no-slip / free-slip) only; the
# constants are run-time arguments.  So a callback whose op sequence has the structure of one of the
# generated recipes runs on the device with its own constants; anything else raises.
# ---------------------------------------------------------------------------------------
_FACES = {'east': E, 'west': W, 'north': N, 'south': S, 'top': T, 'bottom': B}


class BoundaryRecorder:
    '''Stand-in for ``transiflow.BoundaryConditions`` that records the ops a callback applies.'''

    def __init__(self):
        self.ops = []

    def get_forcing(self):
        return None


def _recorder_method(kind, face):
    axis, far = _FACES[face]
    if kind == 'no_slip':
        return lambda self, atom: self.ops.append(('wall', axis, far, NOSLIP))
    if kind == 'free_slip':
        return lambda self, atom: self.ops.append(('wall', axis, far, FREESLIP))
    if kind == 'moving_lid':
        # the tangential velocity that is prescribed: v on the x faces, u elsewhere (BoundaryConditions.py:235-295);
        # the reference applies the no-slip fold right after the forcing
        var = 'v' if axis == 0 else 'u'

        def lid(self, atom, velocity):
            self.ops.append(('force', axis, far, var, 'lidv', velocity))
            self.ops.append(('wall', axis, far, NOSLIP))
        return lid
    if kind == 'temperature':
        return lambda self, atom, temperature: self.ops.append(('force', axis, far, 'T', 'temp', temperature))
    if kind == 'heat_flux':
        return lambda self, atom, heat_flux, biot=0.0: self.ops.append(('force', axis, far, 'T', 'hflux', (heat_flux, biot)))
    if kind == 'salinity_flux':
        return lambda self, atom, salinity_flux: self.ops.append(('force', axis, far, 'S', 'sflux', salinity_flux))
    raise ValueError(kind)


for _kind in ('no_slip', 'free_slip', 'moving_lid', 'temperature', 'heat_flux', 'salinity_flux'):
    for _face in _FACES:
        setattr(BoundaryRecorder, '%s_%s' % (_kind, _face), _recorder_method(_kind, _face))


def _structure(op):
    if op[0] == 'wall':
        return op
    if op[0] == 'force':
        return op[:4] + (op[4] in ('tarr', 'sarr'),)     # constants vs per-face value arrays
    return op


def record_boundary_conditions(callback):
    '''Ops (in our tuple format, constants as literals) that ``callback(bc, atom)`` applies.'''
    rec = BoundaryRecorder()
    callback(rec, None)
    return rec.ops


def match_recorded(ops, dim, nz, dof):
    '''A copy of the generated Config whose recipe has the structure of ``ops`` (same grid family: dim, flat, dof),
    carrying ``ops`` -- i.e. the callback's own constants -- as its recipe; None if no kernel family has that structure.'''
    import copy
    flat = dim == 2 or nz <= 1
    want = [_structure(op) for op in ops]
    for c in CONFIGS:
        if c.dim == dim and c.flat == flat and c.dof == dof and [_structure(op) for op in c.recipe] == want:
            c2 = copy.copy(c)
            c2.recipe = list(ops)
            return c2
    return None
