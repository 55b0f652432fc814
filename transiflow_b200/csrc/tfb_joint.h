// Coupled (vertical velocity, scalar) line solve of the Rayleigh-Benard preconditioner.
//
// The buoyancy block B (scalar -> vertical momentum, Discretization.py:1016-1057) and the
// advection of the scalar's vertical profile C (vertical velocity -> scalar equation,
// Discretization.py:1447-1467) couple w and T so strongly at O(10^3) Rayleigh numbers that a
// block-triangular preconditioner that ignores one of them stalls.  Both blocks are two-point
// averages along z and (for a horizontally uniform profile) diagonal in the horizontal
// directions, so after the x/y transforms of the fast-diagonalisation solve every horizontal
// mode (a, b) is left with ONE banded system along z:
//
//   unknowns  y[2k] = T_k (k = 0..nz-1),  y[2k+1] = w_k (k = 0..nz-2)       N = 2 nz - 1
//   row 2k   :  cT*KT_lo[k] T_{k-1} + Cz_m[k] w_{k-1} + cT*(mu*MT[k] + KT_d[k]) T_k + Cz_0[k] w_k + cT*KT_up[k] T_{k+1}
//   row 2k+1 :  cv*Kw_lo[k] w_{k-1} + Bz_0[k] T_k     + cv*(mu*Mw[k] + Kw_d[k]) w_k + Bz_p[k] T_{k+1} + cv*Kw_up[k] w_{k+1}
//
// (mu = lambda_x[a] + lambda_y[b]); pentadiagonal in this ordering.  tfb_joint_line() factors and
// solves it without pivoting, one call per mode; `stride` is the distance between consecutive
// planes so that neighbouring modes are neighbouring threads (coalesced).  The same function is
// compiled by g++ in tests/cpu_harness, which is how it is checked in the GPU-less container.
#pragma once

#ifndef TFB_HD
#ifdef __CUDACC__
#define TFB_HD __host__ __device__ __forceinline__
#else
#define TFB_HD inline
#endif
#endif

// rows of the 12 x nz coefficient table (tfb_joint_set uploads 0..7, the matrix refresh fills 8..11)
enum {
    TFB_JZ_KW_LO = 0, TFB_JZ_KW_D, TFB_JZ_KW_UP, TFB_JZ_MW,
    TFB_JZ_KT_LO, TFB_JZ_KT_D, TFB_JZ_KT_UP, TFB_JZ_MT,
    TFB_JZ_B0, TFB_JZ_BP, TFB_JZ_C0, TFB_JZ_CM,
    TFB_JZ_ROWS
};

// w, T: right-hand sides on entry, solution on exit (plane k at [k * stride]); al, be: scratch for
// the two upper bands of the factor, N entries each (entry i at [i * stride]).
// VT: storage type of the right-hand sides / solutions (double, or float on the tensor-core path); arithmetic is fp64.
template <class VT>
TFB_HD void tfb_joint_line(int nz, const double* __restrict__ zc, double mu, double cv, double cT, long long stride,
                           VT* __restrict__ w, VT* __restrict__ T, double* __restrict__ al, double* __restrict__ be) {
    const double* KW_LO = zc + TFB_JZ_KW_LO * nz; const double* KW_D = zc + TFB_JZ_KW_D * nz;
    const double* KW_UP = zc + TFB_JZ_KW_UP * nz; const double* MW = zc + TFB_JZ_MW * nz;
    const double* KT_LO = zc + TFB_JZ_KT_LO * nz; const double* KT_D = zc + TFB_JZ_KT_D * nz;
    const double* KT_UP = zc + TFB_JZ_KT_UP * nz; const double* MT = zc + TFB_JZ_MT * nz;
    const double* B0 = zc + TFB_JZ_B0 * nz; const double* BP = zc + TFB_JZ_BP * nz;
    const double* C0 = zc + TFB_JZ_C0 * nz; const double* CM = zc + TFB_JZ_CM * nz;
    // forward elimination; rows i-2 and i-1 are kept as  y_i + a y_{i+1} + b y_{i+2} = g
    double a2p = 0.0, b2p = 0.0, g2p = 0.0;   // row i-2
    double a1p = 0.0, b1p = 0.0, g1p = 0.0;   // row i-1
    for (int k = 0; k < nz; k++) {
        {   // row 2k : scalar equation of plane k
            const double s2 = k > 0 ? cT * KT_LO[k] : 0.0;
            double s1 = k > 0 ? CM[k] : 0.0;
            double d = cT * (mu * MT[k] + KT_D[k]);
            double c1 = k < nz - 1 ? C0[k] : 0.0;
            const double c2 = k < nz - 1 ? cT * KT_UP[k] : 0.0;
            double r = (double)T[k * stride];
            s1 -= s2 * a2p; d -= s2 * b2p; r -= s2 * g2p;
            d -= s1 * a1p; c1 -= s1 * b1p; r -= s1 * g1p;
            const double inv = 1.0 / d;
            a2p = a1p; b2p = b1p; g2p = g1p;
            a1p = c1 * inv; b1p = c2 * inv; g1p = r * inv;
            al[(2LL * k) * stride] = a1p; be[(2LL * k) * stride] = b1p; T[k * stride] = (VT)g1p;
        }
        if (k < nz - 1) {   // row 2k+1 : vertical momentum at the face above plane k
            const double s2 = k > 0 ? cv * KW_LO[k] : 0.0;
            double s1 = B0[k];
            double d = cv * (mu * MW[k] + KW_D[k]);
            double c1 = BP[k];
            const double c2 = k < nz - 2 ? cv * KW_UP[k] : 0.0;
            double r = (double)w[k * stride];
            s1 -= s2 * a2p; d -= s2 * b2p; r -= s2 * g2p;
            d -= s1 * a1p; c1 -= s1 * b1p; r -= s1 * g1p;
            const double inv = 1.0 / d;
            a2p = a1p; b2p = b1p; g2p = g1p;
            a1p = c1 * inv; b1p = c2 * inv; g1p = r * inv;
            al[(2LL * k + 1) * stride] = a1p; be[(2LL * k + 1) * stride] = b1p; w[k * stride] = (VT)g1p;
        }
    }
    // back substitution  y_i = g_i - a_i y_{i+1} - b_i y_{i+2}
    double y1 = 0.0, y2 = 0.0;                // y_{i+1}, y_{i+2}
    for (int k = nz - 1; k >= 0; k--) {
        if (k < nz - 1) {
            const double y = (double)w[k * stride] - al[(2LL * k + 1) * stride] * y1 - be[(2LL * k + 1) * stride] * y2;
            w[k * stride] = (VT)y;
            y2 = y1; y1 = y;
        }
        const double y = (double)T[k * stride] - al[(2LL * k) * stride] * y1 - be[(2LL * k) * stride] * y2;
        T[k * stride] = (VT)y;
        y2 = y1; y1 = y;
    }
}

// ---- the same solve split into a factorisation (once per matrix) and a substitution (every application) ----
// The pivots and the eliminated bands depend on (mode, row) but not on the right-hand side.  tfb_joint_factor_line
// stores, for row i of the interleaved system, the modified first sub-diagonal s1', 1 / pivot and the two upper bands
// (FT: float on the tensor-core path); tfb_joint_substitute_line then needs two fused multiply-adds per band and unknown
// and no division.  Entry i of a factor array sits at [i * stride].
template <class FT>
TFB_HD void tfb_joint_factor_line(int nz, const double* __restrict__ zc, double mu, double cv, double cT, long long stride,
                                  FT* __restrict__ s1p, FT* __restrict__ piv, FT* __restrict__ al, FT* __restrict__ be) {
    const double* KW_LO = zc + TFB_JZ_KW_LO * nz; const double* KW_D = zc + TFB_JZ_KW_D * nz;
    const double* KW_UP = zc + TFB_JZ_KW_UP * nz; const double* MW = zc + TFB_JZ_MW * nz;
    const double* KT_LO = zc + TFB_JZ_KT_LO * nz; const double* KT_D = zc + TFB_JZ_KT_D * nz;
    const double* KT_UP = zc + TFB_JZ_KT_UP * nz; const double* MT = zc + TFB_JZ_MT * nz;
    const double* B0 = zc + TFB_JZ_B0 * nz; const double* BP = zc + TFB_JZ_BP * nz;
    const double* C0 = zc + TFB_JZ_C0 * nz; const double* CM = zc + TFB_JZ_CM * nz;
    double a2p = 0.0, b2p = 0.0, a1p = 0.0, b1p = 0.0;
    for (int k = 0; k < nz; k++) {
        {
            const double s2 = k > 0 ? cT * KT_LO[k] : 0.0;
            double s1 = k > 0 ? CM[k] : 0.0;
            double d = cT * (mu * MT[k] + KT_D[k]);
            double c1 = k < nz - 1 ? C0[k] : 0.0;
            const double c2 = k < nz - 1 ? cT * KT_UP[k] : 0.0;
            s1 -= s2 * a2p; d -= s2 * b2p;
            d -= s1 * a1p; c1 -= s1 * b1p;
            const double inv = 1.0 / d;
            a2p = a1p; b2p = b1p;
            a1p = c1 * inv; b1p = c2 * inv;
            const long long o = (2LL * k) * stride;
            s1p[o] = (FT)s1; piv[o] = (FT)inv; al[o] = (FT)a1p; be[o] = (FT)b1p;
        }
        if (k < nz - 1) {
            const double s2 = k > 0 ? cv * KW_LO[k] : 0.0;
            double s1 = B0[k];
            double d = cv * (mu * MW[k] + KW_D[k]);
            double c1 = BP[k];
            const double c2 = k < nz - 2 ? cv * KW_UP[k] : 0.0;
            s1 -= s2 * a2p; d -= s2 * b2p;
            d -= s1 * a1p; c1 -= s1 * b1p;
            const double inv = 1.0 / d;
            a2p = a1p; b2p = b1p;
            a1p = c1 * inv; b1p = c2 * inv;
            const long long o = (2LL * k + 1) * stride;
            s1p[o] = (FT)s1; piv[o] = (FT)inv; al[o] = (FT)a1p; be[o] = (FT)b1p;
        }
    }
}

// w, T: right-hand sides on entry, solution on exit (plane k at [k * stride]); arithmetic in AT (float or double)
template <class AT, class VT, class FT>
TFB_HD void tfb_joint_substitute_line(int nz, const double* __restrict__ zc, double cv, double cT, long long stride,
                                      VT* __restrict__ w, VT* __restrict__ T, const FT* __restrict__ s1p,
                                      const FT* __restrict__ piv, const FT* __restrict__ al, const FT* __restrict__ be) {
    const double* KW_LO = zc + TFB_JZ_KW_LO * nz;
    const double* KT_LO = zc + TFB_JZ_KT_LO * nz;
    AT g1 = 0, g2 = 0;        // g_{i-1}, g_{i-2}
    for (int k = 0; k < nz; k++) {
        {
            const long long o = (2LL * k) * stride;
            const AT s2 = k > 0 ? (AT)(cT * KT_LO[k]) : (AT)0;
            const AT g = ((AT)T[k * stride] - s2 * g2 - (AT)s1p[o] * g1) * (AT)piv[o];
            T[k * stride] = (VT)g;
            g2 = g1; g1 = g;
        }
        if (k < nz - 1) {
            const long long o = (2LL * k + 1) * stride;
            const AT s2 = k > 0 ? (AT)(cv * KW_LO[k]) : (AT)0;
            const AT g = ((AT)w[k * stride] - s2 * g2 - (AT)s1p[o] * g1) * (AT)piv[o];
            w[k * stride] = (VT)g;
            g2 = g1; g1 = g;
        }
    }
    AT y1 = 0, y2 = 0;
    for (int k = nz - 1; k >= 0; k--) {
        if (k < nz - 1) {
            const long long o = (2LL * k + 1) * stride;
            const AT y = (AT)w[k * stride] - (AT)al[o] * y1 - (AT)be[o] * y2;
            w[k * stride] = (VT)y;
            y2 = y1; y1 = y;
        }
        const long long o = (2LL * k) * stride;
        const AT y = (AT)T[k * stride] - (AT)al[o] * y1 - (AT)be[o] * y2;
        T[k * stride] = (VT)y;
        y2 = y1; y1 = y;
    }
}
