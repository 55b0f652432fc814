#!/usr/bin/env python3
'''Kernel generator for the B200 assembly path.

For every problem configuration in ``recipes.CONFIGS`` this script emits one header
``csrc/gen/rows_<config>.h`` with, per equation (matrix row type) of a grid cell, a
straight-line ``__host__ __device__`` function that computes

  * the structurally non-zero Jacobian entries of that row ("slots", in CSR column order),
  * a bit mask of the slots that exist at this cell (boundary rows are shorter), and
  * the right-hand-side entry F(x) of that row,

directly from the 3x3x3 state neighbourhood and a handful of 1-D grid metrics -- the
reference's dense ``(nx,ny,nz,dof,dof,3,3,3)`` atoms are never materialised.

Every floating-point operation is emitted in the order the reference evaluates it
(/root/reference/transiflow/Discretization.py:229-365, :740-1467 and
BoundaryConditions.py:55-544), with explicit parentheses, so that compiling without FMA
contraction gives bit-identical CSR values and RHS entries.  Boundary conditions are the
reference's ordered in-place edits, unrolled over the slots of a row and predicated on the
face flags of the cell.

The same headers compile with g++ (tests/cpu_harness) which is how the arithmetic is
verified bit-for-bit against the oracle without a GPU.

    python -m transiflow_b200.codegen.gen      # regenerates csrc/gen/*.h
'''
import os
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from transiflow_b200.recipes import CONFIGS  # noqa: E402

AX = 'xyz'
OFF = {-1: 'm', 0: 'o', 1: 'p'}


def others(axis):
    return [a for a in range(3) if a != axis]


class Row:
    '''Code builder for one row function.'''

    def __init__(self, cfg, d1):
        self.cfg, self.d1 = cfg, d1
        self.lin = OrderedDict()   # always-evaluated temporaries: name -> expr
        self.nl = OrderedDict()    # temporaries inside `if (prm.nl)`
        self.nl_out = []           # nonlinear term values visible outside the block

    # ---- temporaries ----
    def let(self, name, expr, nl=False):
        d = self.nl if nl else self.lin
        if name not in self.lin and name not in self.nl:
            d[name] = expr
        return name

    def P(self, d, off):
        '''padded-state value of variable d at unpadded offset off=(ox,oy,oz)'''
        name = 's%d_%s%s%s' % (d, OFF[off[0]], OFF[off[1]], OFF[off[2]])
        return self.let(name, 'P(%d, %d, %d, %d)' % (d, off[0], off[1], off[2]))

    def m(self, name, axis):
        return 'c.%s%s' % (name, AX[axis])

    def avg(self, d, axis, hi, oth=(0, 0)):
        '''average_{x,y,z}: 1/2 P[a] + 1/2 P[a+1] (Discretization.py:1178-1209); hi=0 -> faces
        (-1,0), hi=1 -> faces (0,+1) along axis; oth = offsets on the two other axes.'''
        o1, o2 = others(axis)
        def off(v):
            o = [0, 0, 0]
            o[axis], o[o1], o[o2] = v, oth[0], oth[1]
            return tuple(o)
        a, b = self.P(d, off(hi - 1)), self.P(d, off(hi))
        name = 'av%d%s%d_%s%s' % (d, AX[axis], hi, OFF[oth[0]], OFF[oth[1]])
        return self.let(name, '(0.5 * %s) + (0.5 * %s)' % (a, b), nl=True)

    def wavg(self, d, axis, oth):
        '''weighted_average_{x,y,z}: wm*P[0] + wp*P[+1] along axis (Discretization.py:1211-1266)'''
        o1, o2 = others(axis)
        def off(v):
            o = [0, 0, 0]
            o[axis], o[o1], o[o2] = v, oth[0], oth[1]
            return tuple(o)
        a, b = self.P(d, off(0)), self.P(d, off(1))
        name = 'wa%d%s_%s%s' % (d, AX[axis], OFF[oth[0]], OFF[oth[1]])
        return self.let(name, '(%s * %s) + (%s * %s)' % (self.m('wm', axis), a, self.m('wp', axis), b), nl=True)


def key_dir(d2, axis, o, p1=1, p2=1):
    '''atom key (d2,x,y,z) with offset index o (0,1,2) along axis and p1,p2 on the others'''
    k = [1, 1, 1]
    o1, o2 = others(axis)
    k[axis], k[o1], k[o2] = o, p1, p2
    return (d2, k[0], k[1], k[2])


class CellModel:
    '''Symbolic raw atoms (linear L, convective F, Jacobian-only G) of every row of one cell.'''

    def __init__(self, cfg):
        self.cfg = cfg
        dof = cfg.dof
        self.rows = [Row(cfg, d1) for d1 in range(dof)]
        self.L = [OrderedDict() for _ in range(dof)]   # key -> expr
        self.F = [OrderedDict() for _ in range(dof)]   # key -> [term names] (subtracted in order)
        self.G = [OrderedDict() for _ in range(dof)]
        self._linear()
        self._nonlinear()

    # ---------------- linear part: Discretization.py:229-317 ----------------
    def _lap(self, d, axis, kind):
        '''3-point Laplacian of variable d along axis.  kind selects the metric recipe:
        'own'  (_u_xx: the variable's own direction),  'stag' (_u_yy/_u_zz: a velocity
        differentiated along another axis) or 'cen' (_C_xx: cell-centred scalar).
        Returns (a0, a1, a2) expression names (before the viscous coefficient).'''
        r = self.rows[d]
        o1, o2 = others(axis)
        if kind == 'own':
            # _u_xx(atom, i,j,k, x,y,z): 1/dx*dy*dz with the two other axes in the CALL's order.
            # u_xx: (y,z); v_yy: called (j,i,k,y,x,z) -> (x,z); w_zz: called (k,j,i,z,y,x) -> (y,x)
            b, c = {0: (1, 2), 1: (0, 2), 2: (1, 0)}[axis]
            a0 = '(%s * %s) * %s' % (r.m('rhc', axis), r.m('hc', b), r.m('hc', c))
            a2 = '(%s * %s) * %s' % (r.m('rhp', axis), r.m('hc', b), r.m('hc', c))
        elif kind == 'stag':
            # _u_yy / _u_zz: 1/d(axis)*hu(own)*hc(third); own = the velocity's own axis d
            third = [a for a in range(3) if a != axis and a != d][0]
            a0 = '(%s * %s) * %s' % (r.m('rhm', axis), r.m('hu', d), r.m('hc', third))
            a2 = '(%s * %s) * %s' % (r.m('rhu', axis), r.m('hu', d), r.m('hc', third))
        else:
            # _C_xx(i,j,k,x,y,z): C_xx (y,z); C_yy called (j,i,k,y,x,z) -> (x,z); C_zz (k,j,i,z,y,x) -> (y,x)
            b, c = {0: (1, 2), 1: (0, 2), 2: (1, 0)}[axis]
            a0 = '(%s * %s) * %s' % (r.m('rhm', axis), r.m('hc', b), r.m('hc', c))
            a2 = '(%s * %s) * %s' % (r.m('rhu', axis), r.m('hc', b), r.m('hc', c))
        n0 = r.let('lap%d%s0' % (d, AX[axis]), a0)
        n2 = r.let('lap%d%s2' % (d, AX[axis]), a2)
        n1 = r.let('lap%d%s1' % (d, AX[axis]), '(-%s) - %s' % (n0, n2))
        return n0, n1, n2

    def _area(self, r, b, c, stag_b=False):
        return '%s * %s' % (r.m('hu' if stag_b else 'hc', b), r.m('hc', c))

    def _linear(self):
        cfg = self.cfg
        dim = cfg.dim
        vel = list(range(dim))
        # viscous Laplacians; the centre entry is c*((a1x + a1y) + a1z) in the order u_xx,u_yy,u_zz
        for d in vel:
            r = self.rows[d]
            centre = None
            for axis in range(dim):
                n0, n1, n2 = self._lap(d, axis, 'own' if axis == d else 'stag')
                self.L[d][key_dir(d, axis, 0)] = 'prm.c_visc * %s' % n0
                self.L[d][key_dir(d, axis, 2)] = 'prm.c_visc * %s' % n2
                centre = n1 if centre is None else '(%s + %s)' % (centre, n1)
            self.L[d][(d, 1, 1, 1)] = 'prm.c_visc * %s' % centre
        # pressure gradient (after "- (p_x + p_y + p_z)") and divergence
        prod = {0: (1, 2), 1: (0, 2), 2: (1, 0)}
        for d in vel:
            r = self.rows[d]
            b, c = prod[d]
            A = r.let('Aface%d' % d, self._area(r, b, c))
            self.L[d][key_dir(cfg.p, d, 1)] = A            # 0 - (-A)
            self.L[d][key_dir(cfg.p, d, 2)] = '-%s' % A    # 0 - A
        rp = self.rows[cfg.p]
        for d in vel:
            b, c = prod[d]
            A = rp.let('Aface%d' % d, self._area(rp, b, c))
            self.L[cfg.p][key_dir(d, d, 1)] = A
            self.L[cfg.p][key_dir(d, d, 0)] = '-%s' % A
        # Coriolis (2D only), runtime-gated on beta != 0: atom -= beta * coriolis()
        if dim == 2:
            r = self.rows[1]
            fa = r.let('cor_a', '((%s * %s) * %s) * 0.5' % (r.m('hu', 1), r.m('hc', 0), r.m('hc', 2)))
            va = r.let('cor_av', '%s * c.cor1' % fa)
            for xo in (0, 1):
                for yo in (1, 2):
                    self.L[1][(0, xo, yo, 1)] = ('GATE_BETA', '-(prm.beta * %s)' % va)
            r = self.rows[0]
            fb = r.let('cor_b', '((%s * %s) * %s) * 0.5' % (r.m('hu', 0), r.m('hc', 1), r.m('hc', 2)))
            vb = r.let('cor_bv', '%s * c.cor2' % fb)
            for yo in (0, 1):
                for xo in (1, 2):
                    self.L[0][(1, xo, yo, 1)] = ('GATE_BETA', '-(prm.beta * %s)' % vb)
        # scalars
        baxis = 1 if cfg.flat else 2   # buoyancy acts on v when nz == 1 (Discretization.py:298-308)
        for v, coef, sign in ((cfg.T, 'prm.c_T', '+'), (cfg.S, 'prm.c_S', '-')):
            if v >= cfg.dof:
                continue
            r = self.rows[v]
            centre = None
            for axis in range(dim):
                n0, n1, n2 = self._lap(v, axis, 'cen')
                self.L[v][key_dir(v, axis, 0)] = '%s * %s' % (coef, n0)
                self.L[v][key_dir(v, axis, 2)] = '%s * %s' % (coef, n2)
                centre = n1 if centre is None else '(%s + %s)' % (centre, n1)
            self.L[v][(v, 1, 1, 1)] = '%s * %s' % (coef, centre)
            # forward_average_C_{y,z}: _forward_average_x called with the buoyancy axis first
            rb = self.rows[baxis]
            b, c = prod[baxis]
            fav = rb.let('buoy', '((%s * %s) * %s) * 0.5' % (rb.m('hu', baxis), rb.m('hc', b), rb.m('hc', c)))
            for o in (1, 2):
                self.L[baxis][key_dir(v, baxis, o)] = fav if sign == '+' else '-%s' % fav
        # Rayleigh-Benard perturbation source, runtime-gated: atom += Bi/(Bi+1) * backward_average
        if cfg.has_T and cfg.problem == 1:
            r = self.rows[cfg.T]
            b, c = prod[baxis]
            bav = r.let('pert', '((%s * %s) * %s) * 0.5' % (r.m('hu', baxis), r.m('hc', b), r.m('hc', c)))
            for o in (0, 1):
                self.L[cfg.T][key_dir(baxis, baxis, o)] = ('GATE_PERT', 'prm.c_pert * %s' % bav)

    # ---------------- nonlinear part: Discretization.py:319-365, :1268-1467 ----------------
    def _sub_pair(self, tab, d1, d2, axis, hi, term):
        for o in (hi, hi + 1):
            tab[d1].setdefault(key_dir(d2, axis, o), []).append(term)

    def _conv_self(self, d, A_expr_bc):
        '''u_u_x / v_v_y / w_w_z: F and an identical copy in G'''
        r = self.rows[d]
        b, c = A_expr_bc
        A = r.let('Aface%d' % d, '%s * %s' % (r.m('hc', b), r.m('hc', c)))
        tlo = r.let('t%d%d_lo' % (d, d), '((-%s) * %s) * 0.5' % (A, r.avg(d, d, 0)), nl=True)
        thi = r.let('t%d%d_hi' % (d, d), '(%s * %s) * 0.5' % (A, r.avg(d, d, 1)), nl=True)
        r.nl_out += [tlo, thi]
        for tab in (self.F, self.G):
            self._sub_pair(tab, d, d, d, 0, tlo)
            self._sub_pair(tab, d, d, d, 1, thi)

    def _conv_cross(self, adv, d, area):
        '''advecting velocity `adv` (flux direction = axis adv) transporting velocity d:
        u_v_x (adv=0,d=1), u_w_x (0,2), v_u_y (1,0), v_w_y (1,2), w_u_z (2,0), w_w... etc.
        area = (staggered axis, centred axis) of the face area hu*hc in the reference's order.'''
        r = self.rows[d]
        A = r.let('Ac%d%d' % (adv, d), '%s * %s' % (r.m('hu', area[0]), r.m('hc', area[1])))
        # F: weighted average of the advecting velocity along axis d, at flux-axis offsets -1 / 0
        o1, o2 = others(d)
        def oth(v):
            return (v, 0) if o1 == adv else (0, v)
        tlo = r.let('t%d%d_lo' % (adv, d), '((-%s) * %s) * 0.5' % (A, r.wavg(adv, d, oth(-1))), nl=True)
        thi = r.let('t%d%d_hi' % (adv, d), '(%s * %s) * 0.5' % (A, r.wavg(adv, d, oth(0))), nl=True)
        r.nl_out += [tlo, thi]
        self._sub_pair(self.F, d, d, adv, 0, tlo)
        self._sub_pair(self.F, d, d, adv, 1, thi)
        # G[d, adv]: flux-axis offset index 0/1, spread over offsets 1,2 of axis d with weights
        for hi, sgn in ((0, '(-%s)' % A), (1, A)):
            base = r.let('gb%d%d_%d' % (adv, d, hi), '%s * %s' % (sgn, r.avg(d, adv, hi)), nl=True)
            for o, w in ((1, 'wm'), (2, 'wp')):
                g = r.let('g%d%d_%d%d' % (adv, d, hi, o), '%s * %s' % (base, r.m(w, d)), nl=True)
                r.nl_out.append(g)
                k = [1, 1, 1]
                k[adv], k[d] = hi, o
                self.G[d].setdefault((adv, k[0], k[1], k[2]), []).append(g)

    def _conv_scalar(self, adv, v):
        '''u_C_x / v_C_y / w_C_z'''
        r = self.rows[v]
        b, c = {0: (1, 2), 1: (0, 2), 2: (1, 0)}[adv]
        A = r.let('Aface%d' % adv, '%s * %s' % (r.m('hc', b), r.m('hc', c)))
        def off(val):
            o = [0, 0, 0]
            o[adv] = val
            return tuple(o)
        tlo = r.let('t%d%d_lo' % (adv, v), '((-%s) * %s) * 0.5' % (A, r.P(adv, off(-1))), nl=True)
        thi = r.let('t%d%d_hi' % (adv, v), '(%s * %s) * 0.5' % (A, r.P(adv, off(0))), nl=True)
        r.nl_out += [tlo, thi]
        self._sub_pair(self.F, v, v, adv, 0, tlo)
        self._sub_pair(self.F, v, v, adv, 1, thi)
        for hi, sgn in ((0, '(-%s)' % A), (1, A)):
            g = r.let('g%d%d_%d' % (adv, v, hi), '%s * %s' % (sgn, r.avg(v, adv, hi)), nl=True)
            r.nl_out.append(g)
            self.G[v].setdefault(key_dir(adv, adv, hi), []).append(g)

    def _nonlinear(self):
        cfg = self.cfg
        self._conv_self(0, (1, 2))                 # u_u_x :1268
        self._conv_cross(0, 1, (1, 2))             # u_v_x :1281  _backward_u_y(j,i,k,y,x,z): hu_y*hc_z
        self._conv_cross(1, 0, (0, 2))             # v_u_y :1335  _backward_u_y(i,j,k,x,y,z): hu_x*hc_z
        self._conv_self(1, (0, 2))                 # v_v_y :1351
        if cfg.dim > 2:
            self._conv_cross(0, 2, (2, 1))         # u_w_x :1297  _backward_u_z(k,j,i,z,y,x): hu_z*hc_y
            self._conv_cross(1, 2, (2, 0))         # v_w_y :1364  _backward_u_y(k,j,i,z,y,x): hu_z*hc_x
            self._conv_cross(2, 0, (0, 1))         # w_u_z :1402  _backward_u_z(i,j,k,x,y,z): hu_x*hc_y
            self._conv_cross(2, 1, (1, 0))         # w_v_z :1418  _backward_u_z(j,i,k,y,x,z): hu_y*hc_x
            self._conv_self(2, (1, 0))             # w_w_z :1434
        for v in (cfg.T, cfg.S):
            if v < cfg.dof:
                for adv in range(cfg.dim):
                    self._conv_scalar(adv, v)      # u_C_x :1313, v_C_y :1380, w_C_z :1447


def cname(prefix, key):
    return '%s%d_%d%d%d' % (prefix, key[0], key[1], key[2], key[3])


class RowEmitter:
    def __init__(self, cfg, model, d1):
        self.cfg, self.model, self.d1 = cfg, model, d1
        self.row = model.rows[d1]
        L, F, G = model.L[d1], model.F[d1], model.G[d1]
        fkeys = set(L) | set(F)
        jkeys = fkeys | set(G)
        # CSR column order: (z, y, x, d2); with the z-fold the key order inside one column is z = 0,1,2
        self.jkeys = sorted(jkeys, key=lambda k: (k[3], k[2], k[1], k[0]))
        self.fkeys = sorted(fkeys, key=lambda k: (k[3], k[2], k[1], k[0]))
        if cfg.fold:
            cols = OrderedDict()
            for k in sorted(jkeys, key=lambda k: (k[2], k[1], k[0], k[3])):
                cols.setdefault((k[0], k[1], k[2]), []).append(k)
            self.cols = list(cols.items())   # [((d2,x,y), [keys z-ordered])]
        else:
            self.cols = [((k[0], k[1], k[2], k[3]), [k]) for k in self.jkeys]
        assert len(self.cols) <= 32
        self.L, self.F, self.G = L, F, G

    # ---- BC op unrolling on one slot set ----
    def _wall(self, pre, keys, op, lines, with_frc):
        _, axis, far, sign = op
        d1, q = self.d1, axis
        ks = set(keys)
        n_out = 2 if far else 0
        sg = '-' if sign < 0 else '+'
        flag = 'c.%s[%d]' % ('far' if far else 'near', axis)
        body = []
        def fold():
            for k in keys:
                if k[1 + axis] == 1:
                    kk = list(k)
                    kk[1 + axis] = n_out
                    kk = tuple(kk)
                    if kk in ks:
                        body.append('%s = %s %s %s;' % (cname(pre, k), cname(pre, k), sg, cname(pre, kk)))
        def zero(cond):
            for k in keys:
                if cond(k):
                    body.append('%s = 0.0;' % cname(pre, k))
        for k in keys:   # closure check: an outside entry always has its centre partner
            if k[1 + axis] != 1:
                kk = list(k)
                kk[1 + axis] = 1
                assert tuple(kk) in ks, (self.cfg.name, d1, k)
        if far:
            fold()
            zero(lambda k: k[0] == q and k[1 + axis] == 1)
            if d1 == q:
                zero(lambda k: True)
            zero(lambda k: k[1 + axis] == 2)
            if d1 == q:
                body.append('%s = -1.0;' % cname(pre, (q, 1, 1, 1)))
            if with_frc and d1 == (1 if axis == 2 else q):   # top zeroes frc[...,1] (sic)
                body.append('frc = 0.0;')
        else:
            zero(lambda k: k[0] == q and k[1 + axis] == 0)
            fold()
            zero(lambda k: k[1 + axis] == 0)
        if body:
            lines.append('if (%s) {' % flag)
            lines += ['    ' + b for b in body]
            lines.append('}')
        if far and d1 == q:
            b2 = ['%s = 0.0;' % cname(pre, k) for k in keys if k[0] == q and k[1 + axis] == 2]
            if b2:
                lines.append('if (c.far2[%d]) {' % axis)
                lines += ['    ' + b for b in b2]
                lines.append('}')

    def _force(self, pre, keys, op, fidx, lines, with_frc):
        _, axis, far, var, kind, arg = op
        v = self.cfg.var(var)
        d1 = self.d1
        ks = set(keys)
        n_out = 2 if far else 0
        flag = 'c.%s[%d]' % ('far' if far else 'near', axis)
        body = []
        if with_frc and d1 == v:
            terms = [k for k in keys if k[0] == v and k[1 + axis] == n_out]
            # in-plane order: second in-plane axis outer, first inner (BoundaryConditions.py:475)
            o1, o2 = others(axis)
            terms.sort(key=lambda k: (k[1 + o2], k[1 + o1]))
            if terms:
                for k in terms:
                    assert k[1 + o1] == 1 and k[1 + o2] == 1, 'in-plane forcing offsets unsupported'
                val = 'c.fval[%d]' % fidx if kind in ('tarr', 'sarr') else 'prm.bc_cf[%d]' % fidx
                expr = ' + '.join('(%s * %s)' % (cname(pre, k), val) for k in terms)
                body.append('frc = frc + (%s);' % expr)
        for k in keys:
            if k[0] == v and k[1 + axis] == 1:
                kk = list(k)
                kk[1 + axis] = n_out
                kk = tuple(kk)
                if kk in ks:
                    body.append('%s = %s + (prm.bc_ca[%d] * %s);' % (cname(pre, k), cname(pre, k), fidx, cname(pre, kk)))
        for k in keys:
            if k[0] == v and k[1 + axis] == n_out:
                body.append('%s = 0.0;' % cname(pre, k))
        if body:
            lines.append('if (%s) {' % flag)
            lines += ['    ' + b for b in body]
            lines.append('}')

    def _pin(self, pre, keys, lines, with_frc):
        cfg, d1 = self.cfg, self.d1
        for k in keys:
            if k[0] == cfg.S:
                lines.append('if (c.pin(%d, %d, %d)) %s = 0.0;' % (k[1] - 1, k[2] - 1, k[3] - 1, cname(pre, k)))
        if d1 == cfg.S:
            lines.append('if (c.cell0) {')
            lines.append('    %s = -1.0;' % cname(pre, (cfg.S, 1, 1, 1)))
            if with_frc:
                lines.append('    frc = 0.0;')
            lines.append('}')

    def bc_lines(self, pre, keys, with_frc):
        '''BCM template parameter: 0 = no boundary code (interior cells), 1 = only the x-face ops
        (cells that are interior in y and z: the warp-uniform fast path keeps x-edge lanes without
        a second pass), 2 = the full recipe.'''
        lines = []
        fidx = 0
        for op in self.cfg.recipe:
            sub = []
            if op[0] == 'wall':
                self._wall(pre, keys, op, sub, with_frc)
                level = 1 if op[1] == 0 else 2
            elif op[0] == 'force':
                self._force(pre, keys, op, fidx, sub, with_frc)
                fidx += 1
                level = 1 if op[1] == 0 else 2
            else:
                self._pin(pre, keys, sub, with_frc)
                level = 2
            if sub:
                lines.append('if (BCM >= %d) {' % level)
                lines += ['    ' + ln for ln in sub]
                lines.append('}')
        return lines

    # ---- structural mask (which slots exist at this cell) ----
    def mask_lines(self):
        '''Boolean shadow of the BC ops on the Jacobian slot set.  Returns code that computes
        `unsigned m` (bit = output column) from the face flags only.'''
        keys = self.jkeys
        ks = set(keys)
        lines = []
        for k in keys:
            lines.append('bool %s = true;' % cname('b', k))
        for op in self.cfg.recipe:
            if op[0] == 'wall':
                _, axis, far, sign = op
                q, n_out = axis, (2 if far else 0)
                body = []
                def fold():
                    for k in keys:
                        if k[1 + axis] == 1:
                            kk = list(k); kk[1 + axis] = n_out; kk = tuple(kk)
                            if kk in ks:
                                body.append('%s = %s || %s;' % (cname('b', k), cname('b', k), cname('b', kk)))
                def zero(cond):
                    for k in keys:
                        if cond(k):
                            body.append('%s = false;' % cname('b', k))
                if far:
                    fold()
                    zero(lambda k: k[0] == q and k[1 + axis] == 1)
                    if self.d1 == q:
                        zero(lambda k: True)
                    zero(lambda k: k[1 + axis] == 2)
                    if self.d1 == q:
                        body.append('%s = true;' % cname('b', (q, 1, 1, 1)))
                else:
                    zero(lambda k: k[0] == q and k[1 + axis] == 0)
                    fold()
                    zero(lambda k: k[1 + axis] == 0)
                if body:
                    lines.append('if (c.%s[%d]) { %s }' % ('far' if far else 'near', axis, ' '.join(body)))
                if far and self.d1 == q:
                    b2 = ['%s = false;' % cname('b', k) for k in keys if k[0] == q and k[1 + axis] == 2]
                    if b2:
                        lines.append('if (c.far2[%d]) { %s }' % (axis, ' '.join(b2)))
            elif op[0] == 'force':
                _, axis, far, var, kind, arg = op
                v, n_out = self.cfg.var(var), (2 if far else 0)
                body = []
                for k in keys:
                    if k[0] == v and k[1 + axis] == 1:
                        kk = list(k); kk[1 + axis] = n_out; kk = tuple(kk)
                        if kk in ks:
                            body.append('%s = %s || %s;' % (cname('b', k), cname('b', k), cname('b', kk)))
                for k in keys:
                    if k[0] == v and k[1 + axis] == n_out:
                        body.append('%s = false;' % cname('b', k))
                if body:
                    lines.append('if (c.%s[%d]) { %s }' % ('far' if far else 'near', axis, ' '.join(body)))
            else:
                for k in keys:
                    if k[0] == self.cfg.S:
                        lines.append('if (c.pin(%d, %d, %d)) %s = false;' % (k[1] - 1, k[2] - 1, k[3] - 1, cname('b', k)))
                if self.d1 == self.cfg.S:
                    lines.append('if (c.cell0) %s = true;' % cname('b', (self.cfg.S, 1, 1, 1)))
        lines.append('unsigned m = 0u;')
        for ci, (col, ckeys) in enumerate(self.cols):
            lines.append('if (%s) m |= %du;' % (' || '.join(cname('b', k) for k in ckeys), 1 << ci))
        return lines

    # ---- the row function ----
    def emit(self):
        cfg, d1, row = self.cfg, self.d1, self.row
        L, F, G = self.L, self.F, self.G
        out = []
        w = out.append
        name = '%s_row%d' % (cfg.name, d1)
        w('template <bool DO_J, bool DO_F, int BCM, class Cell, class State, class Sink>')
        w('TFB_HD void %s(const TfbParams& prm, const Cell& c, const State& P, Sink& Jout, double& rhs_out) {' % name)
        # make sure every state value the RHS product needs is loaded
        for k in self.fkeys:
            row.P(k[0], (k[1] - 1, k[2] - 1, k[3] - 1))
        for n, e in row.lin.items():
            w('    const double %s = %s;' % (n, e))
        outs = list(OrderedDict.fromkeys(row.nl_out))
        if outs:
            w('    double %s;' % ', '.join('%s = 0.0' % o for o in outs))
            w('    if (prm.nl) {')
            for n, e in row.nl.items():
                if n in outs:
                    w('        %s = %s;' % (n, e))
                else:
                    w('        const double %s = %s;' % (n, e))
            w('    }')
        # chains
        def chain(terms):
            e = '(-%s)' % terms[0]
            for t in terms[1:]:
                e = '(%s - %s)' % (e, t)
            return e
        def lin_expr(k):
            v = L[k]
            if isinstance(v, tuple):
                gate = {'GATE_BETA': 'prm.has_beta', 'GATE_PERT': 'prm.pert'}[v[0]]
                return '(%s ? (%s) : 0.0)' % (gate, v[1])
            return '(%s)' % v
        for k in self.jkeys:
            if k in F:
                w('    const double %s = %s;' % (cname('f', k), chain(F[k])))
            if k in L:
                w('    const double %s = %s;' % (cname('l', k), lin_expr(k)))
        w('    double frc = 0.0;')
        w('    (void)frc;')
        # F atom (rhs)
        w('    if (DO_F) {')
        for k in self.fkeys:
            if k in F and k in L:
                e = '%s + %s' % (cname('f', k), cname('l', k))
            elif k in F:
                e = cname('f', k)
            else:
                e = cname('l', k)
            w('        double %s = %s;' % (cname('F', k), e))
        for ln in self.bc_lines('F', self.fkeys, True):
            w('        ' + ln)
        terms = ['(%s * %s)' % (cname('F', k), row.P(k[0], (k[1] - 1, k[2] - 1, k[3] - 1))) for k in self.fkeys]
        e = terms[0]
        for t in terms[1:]:
            e = '(%s + %s)' % (e, t)
        w('        rhs_out = %s + frc;' % e)
        w('    }')
        # J atom
        w('    if (DO_J) {')
        for k in self.jkeys:
            parts = None
            if k in G:
                parts = chain(G[k])
            if k in F:
                parts = cname('f', k) if parts is None else '(%s + %s)' % (parts, cname('f', k))
            if k in L:
                parts = cname('l', k) if parts is None else '%s + %s' % (parts, cname('l', k))
            w('        double %s = %s;' % (cname('J', k), parts))
        for ln in self.bc_lines('J', self.jkeys, False):
            w('        ' + ln)
        for ci, (col, ckeys) in enumerate(self.cols):
            if len(ckeys) == 1:
                w('        Jout.put(%d, %s);' % (ci, cname('J', ckeys[0])))
            else:
                # z-fold: CrsMatrix.compress merges duplicates in emission order (z = 0,1,2) after
                # assemble_jacobian dropped entries with |a| <= 1e-14 (Discretization.py:515)
                e = '0.0'
                for k in ckeys:
                    e = '(%s + tfb_keep(%s))' % (e, cname('J', k))
                w('        Jout.put(%d, %s);' % (ci, e))
        w('    }')
        w('}')
        w('')
        # mask function
        w('template <class Cell>')
        w('TFB_HD unsigned %s_mask%d(const Cell& c) {' % (cfg.name, d1))
        for ln in self.mask_lines():
            w('    ' + ln)
        w('    return m;')
        w('}')
        w('')
        return out


def emit_config(cfg):
    model = CellModel(cfg)
    ems = [RowEmitter(cfg, model, d1) for d1 in range(cfg.dof)]
    out = []
    w = out.append
    w('// GENERATED by transiflow_b200/codegen/gen.py -- do not edit.')
    w('// Config %s: dim=%d flat=%d dof=%d problem=%d' % (cfg.name, cfg.dim, cfg.flat, cfg.dof, cfg.problem))
    w('#pragma once')
    w('#include "../tfb_rows_common.h"')
    w('')
    for em in ems:
        out += em.emit()
    maxs = max(len(em.cols) for em in ems)
    w('struct Cfg_%s {' % cfg.name)
    w('    static constexpr int ID = %d, DIM = %d, DOF = %d, FLAT = %d, FOLD = %d, MAXSLOT = %d, NFORCE = %d;' % (
        cfg.cid, cfg.dim, cfg.dof, int(cfg.flat), int(cfg.fold), maxs, cfg.nforce))
    w('    static constexpr int CELL_SLOTS = %d;' % sum(len(em.cols) for em in ems))
    w('    static constexpr const char* NAME = "%s";' % cfg.name)
    w('    TFB_HD static constexpr int nslot(int d1) {')
    w('        switch (d1) { %s default: return 0; }' % ' '.join('case %d: return %d;' % (d, len(em.cols)) for d, em in enumerate(ems)))
    w('    }')
    # slot table: for each row and output column: d2, dx, dy, dz (dz = 0 for folded columns)
    w('    TFB_HD static void slot(int d1, int s, int& d2, int& dx, int& dy, int& dz) {')
    w('        int code = 0;')
    w('        switch (d1) {')
    for d, em in enumerate(ems):
        codes = []
        for col, ckeys in em.cols:
            k = ckeys[0]
            z = 1 if cfg.fold else k[3]
            codes.append(k[0] | (k[1] << 4) | (k[2] << 6) | (z << 8))
        w('        case %d: { const short t[%d] = {%s}; code = t[s]; break; }' % (d, len(codes), ', '.join(map(str, codes))))
    w('        }')
    w('        d2 = code & 15; dx = ((code >> 4) & 3) - 1; dy = ((code >> 6) & 3) - 1; dz = ((code >> 8) & 3) - 1;')
    w('    }')
    w('    template <bool DO_J, bool DO_F, int BCM, class Cell, class State, class Sink>')
    w('    TFB_HD static void row(int d1, const TfbParams& prm, const Cell& c, const State& P, Sink& Jout, double& rhs_out) {')
    w('        switch (d1) {')
    for d in range(cfg.dof):
        w('        case %d: %s_row%d<DO_J, DO_F, BCM>(prm, c, P, Jout, rhs_out); break;' % (d, cfg.name, d))
    w('        }')
    w('    }')
    w('    template <class Cell>')
    w('    TFB_HD static unsigned mask(int d1, const Cell& c) {')
    w('        switch (d1) {')
    for d in range(cfg.dof):
        w('        case %d: return %s_mask%d(c);' % (d, cfg.name, d))
    w('        }')
    w('        return 0u;')
    w('    }')
    w('};')
    return '\n'.join(out) + '\n'


def main():
    outdir = os.path.join(os.path.dirname(HERE), 'csrc', 'gen')
    os.makedirs(outdir, exist_ok=True)
    names = []
    for cfg in CONFIGS:
        src = emit_config(cfg)
        with open(os.path.join(outdir, 'rows_%s.h' % cfg.name), 'w') as f:
            f.write(src)
        names.append(cfg.name)
        print('rows_%s.h: %d lines' % (cfg.name, src.count('\n')))
    with open(os.path.join(outdir, 'all_configs.h'), 'w') as f:
        f.write('// GENERATED by transiflow_b200/codegen/gen.py -- do not edit.\n#pragma once\n')
        for n in names:
            f.write('#include "rows_%s.h"\n' % n)
        f.write('#define TFB_FOR_EACH_CONFIG(X) %s\n' % ' '.join('X(Cfg_%s)' % n for n in names))
        f.write('#define TFB_NUM_CONFIGS %d\n' % len(names))


if __name__ == '__main__':
    main()
