// Shared definitions for the generated row functions (csrc/gen/rows_*.h).
// Plain C++ so the same code compiles for sm_100a (nvcc) and for the CPU verification
// harness (g++, tests/cpu_harness) -- the arithmetic is identical by construction.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TFB_HD __host__ __device__ __forceinline__
#else
#define TFB_HD inline
#endif

#define TFB_MAX_FORCE 8

// Per-call scalars, evaluated on the host exactly as the reference does
// (Discretization.py:236-246, 278-288; BoundaryConditions.py:339-469).
struct TfbParams {
    double c_visc;   // 1 / (Re * sqrt(Gr))
    double c_T;      // 1 / (Pr * sqrt(Gr))
    double c_S;      // 1 / (Le * Pr * sqrt(Gr))
    double c_pert;   // Bi / (Bi + 1)
    double beta;     // Rossby parameter
    double bc_cf[TFB_MAX_FORCE];  // forcing_constant of the i-th 'force' op of the recipe
    double bc_ca[TFB_MAX_FORCE];  // atom_constant of the i-th 'force' op
    int nl;          // convective terms on? (Discretization.py:333-335)
    int has_beta;    // beta != 0
    int pert;        // problem is 'Rayleigh-Benard Perturbation'
    int pad_;
};

// assemble_jacobian only emits |a| > 1e-14 (Discretization.py:515); needed where duplicate
// columns are merged afterwards (z-fold of semi-2D grids).
TFB_HD double tfb_keep(double a) { return fabs(a) > 1e-14 ? a : 0.0; }

// Where a row function delivers its Jacobian slot values (slot index s is a compile-time
// constant at every call site after inlining).
struct TfbArraySink {
    double* v;
    TFB_HD void put(int s, double x) { v[s] = x; }
};
