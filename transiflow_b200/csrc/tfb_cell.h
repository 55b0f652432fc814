// Per-cell context shared by the CUDA assembly kernels and the CPU verification harness.
#pragma once
#include "tfb_rows_common.h"

#define TFB_NMET 8  // per-axis metric arrays: hc, hu, rhc, rhp, rhm, rhu, wm, wp

// Read-only view of the grid geometry (device pointers on the GPU, host pointers in the harness).
struct TfbGrid {
    int nx, ny, nz, dim, dof;
    int zfold;                     // nz == 1: the z direction is periodic and folds onto one plane
    const double* met[3];          // met[a][m * n_a + idx], m in 0..TFB_NMET-1
    const double* cor;             // [2][ny]: y[j]/2 and -(y[j]+y[j-1])/4 (Discretization.py:1093,1098)
    const double* fval[TFB_MAX_FORCE];  // per 'force' op: face value array (in-plane, first axis fastest) or null
    signed char fdir[TFB_MAX_FORCE];    // axis of the face of each 'force' op (for fval indexing)
};

struct TfbCell {
    double hcx, hux, rhcx, rhpx, rhmx, rhux, wmx, wpx;
    double hcy, huy, rhcy, rhpy, rhmy, rhuy, wmy, wpy;
    double hcz, huz, rhcz, rhpz, rhmz, rhuz, wmz, wpz;
    double cor1, cor2;
    double fval[TFB_MAX_FORCE];
    int i, j, k, nx, ny, nz;
    bool near[3], far[3], far2[3];
    bool cell0;
    // AMOC: is the cell at offset (dx,dy,dz), with the reference's modulo wrap, cell 0?
    // (Discretization.py:692-696)
    TFB_HD bool pin(int dx, int dy, int dz) const {
        int a = (i + dx) % nx, b = (j + dy) % ny, cc = (k + dz) % nz;
        return a == 0 && b == 0 && cc == 0;   // a negative remainder (-1) is never cell 0
    }
};

#define TFB_LOADMET(ax, n, idx, C)                                   \
    C.hc##ax = g.met[A][0 * (n) + (idx)]; C.hu##ax = g.met[A][1 * (n) + (idx)];    \
    C.rhc##ax = g.met[A][2 * (n) + (idx)]; C.rhp##ax = g.met[A][3 * (n) + (idx)];  \
    C.rhm##ax = g.met[A][4 * (n) + (idx)]; C.rhu##ax = g.met[A][5 * (n) + (idx)];  \
    C.wm##ax = g.met[A][6 * (n) + (idx)]; C.wp##ax = g.met[A][7 * (n) + (idx)];

TFB_HD int tfb_far2_index(int m) { return m >= 2 ? m - 2 : 2 * m - 2; }  // Python index m-2 with wrap

// everything of a TfbCell except the metrics: position, face flags, per-face forcing values
template <int NFORCE>
TFB_HD void tfb_cell_flags(const TfbGrid& g, int i, int j, int k, TfbCell& c) {
    c.i = i; c.j = j; c.k = k; c.nx = g.nx; c.ny = g.ny; c.nz = g.nz;
    c.near[0] = i == 0; c.far[0] = i == g.nx - 1; c.far2[0] = i == tfb_far2_index(g.nx);
    c.near[1] = j == 0; c.far[1] = j == g.ny - 1; c.far2[1] = j == tfb_far2_index(g.ny);
    c.near[2] = k == 0; c.far[2] = k == g.nz - 1; c.far2[2] = k == tfb_far2_index(g.nz);
    c.cell0 = (i == 0 && j == 0 && k == 0);
#pragma unroll
    for (int f = 0; f < NFORCE; f++) {
        double v = 0.0;
        if (g.fval[f]) {
            int a = g.fdir[f];
            int i1 = a == 0 ? j : i, i2 = a == 2 ? j : k, n1 = a == 0 ? g.ny : g.nx;
            v = g.fval[f][i1 + n1 * i2];
        }
        c.fval[f] = v;
    }
}

template <int NFORCE>
TFB_HD void tfb_make_cell(const TfbGrid& g, int i, int j, int k, TfbCell& c) {
    { constexpr int A = 0; TFB_LOADMET(x, g.nx, i, c) }
    { constexpr int A = 1; TFB_LOADMET(y, g.ny, j, c) }
    { constexpr int A = 2; TFB_LOADMET(z, g.nz, k, c) }
    c.cor1 = g.cor[j];
    c.cor2 = g.cor[g.ny + j];
    tfb_cell_flags<NFORCE>(g, i, j, k, c);
}

// Padded-state semantics of utils.create_padded_state_mtx (utils.py:62-133) for the
// non-periodic x/y directions and the (periodic iff nz == 1) z direction: zero outside the
// domain, wall-normal velocity on the far walls forced to zero, z-fold onto the single plane.
// `kofs` maps a global k to the plane index of `state` (slab-local storage has ghost planes).
TFB_HD double tfb_padded_load(const TfbGrid& g, const double* __restrict__ state, int kofs,
                                     int ii, int jj, int kk, int d) {
    if (g.zfold) kk = 0;
    else if (kk < 0 || kk >= g.nz) return 0.0;
    if (ii < 0 || ii >= g.nx || jj < 0 || jj >= g.ny) return 0.0;
    if ((ii == g.nx - 1 && d == 0) || (jj == g.ny - 1 && d == 1) || (!g.zfold && kk == g.nz - 1 && d == 2)) return 0.0;
    return state[(((long long)(kk + kofs) * g.ny + jj) * g.nx + ii) * g.dof + d];
}

// Global column of slot (d2, dx, dy, dz) of cell (i,j,k), Discretization.py:516-518.
TFB_HD long long tfb_column(const TfbGrid& g, int i, int j, int k, int d2, int dx, int dy, int dz) {
    int a = (i + dx + g.nx) % g.nx, b = (j + dy + g.ny) % g.ny, cc = (k + dz + g.nz) % g.nz;
    return (((long long)cc * g.ny + b) * g.nx + a) * g.dof + d2;
}
