/*
 * tfb200.h -- C ABI of the B200-native TransiFlow backend (libtfb200.so).
 *
 * The reference (BIMAU/transiflow) is pure Python and has no FFI; this ABI is the boundary a
 * `transiflow/interface/<Backend>.py` module binds with ctypes.  Each entry
 * point names the reference method it replaces (paths relative to /root/reference/transiflow).
 *
 * Conventions: every function returns 0 on success, <0 on error (message via
 * tfb_last_error()), >0 for "did not converge" style soft statuses.  All vectors are
 * contiguous fp64 in the reference's state ordering [u,v,(w),p,(T),(S)] per cell, cells
 * i-fastest then j, k (Discretization.py:20-26).  With z-slab partitioning every rank passes
 * only the planes it owns; halos are exchanged inside the library.
 */
#ifndef TFB200_H
#define TFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFB_MAX_FORCE 8
#define TFB_NMET 8

typedef struct tfb_ctx tfb_ctx; /* one discretised problem (or one z-slab of it) on one GPU */
typedef struct tfb_mat tfb_mat; /* CSR values on the ctx's fixed sparsity pattern */

/* Per-call scalars; layout identical to struct TfbParams in csrc/tfb_rows_common.h.
 * Evaluated by the host exactly like Discretization._linear_part_2D/3D (Discretization.py:229-317)
 * and BoundaryConditions.heat_flux_* etc. (BoundaryConditions.py:339-469). */
typedef struct {
    double c_visc, c_T, c_S, c_pert, beta;
    double bc_cf[TFB_MAX_FORCE], bc_ca[TFB_MAX_FORCE];
    int32_t nl, has_beta, pert, pad_;
} tfb_params;

/* Problem descriptor: replaces the constructor arguments of Discretization
 * (Discretization.py:106-143).  Coordinate vectors enter as the per-axis metric arrays of
 * hostprep.axis_metrics (shape (TFB_NMET, n_axis), host memory, copied). */
typedef struct {
    int32_t config;          /* generated kernel family, recipes.Config.cid */
    int32_t nx, ny, nz, dim, dof;
    int32_t device;          /* CUDA device ordinal */
    int32_t k0, k1;          /* owned z-planes [k0,k1); single GPU: 0, nz */
    const double* met[3];
    const double* cor;       /* (2, ny) Coriolis metrics */
} tfb_desc;

int tfb_device_count(void);
const char* tfb_last_error(void);
const char* tfb_config_name(int config);

int tfb_create(const tfb_desc* desc, tfb_ctx** out);
void tfb_destroy(tfb_ctx* ctx);

/* Discretization.set_parameter / the shared parameter dict: scalars for the next calls.
 * fval[f]: optional host array of face values for 'force' op f (AMOC), in-plane order;
 * frc_static: optional host vector (local rows) added to the RHS (wind stress). */
int tfb_set_params(tfb_ctx* ctx, const tfb_params* prm, const double* const* fval,
                   const int8_t* fdir, const double* frc_static);

/* sizes of the local slab: rows, structural non-zeros, first global row */
int tfb_sizes(tfb_ctx* ctx, int64_t* n_local, int64_t* nnz_local, int64_t* n_global, int64_t* row0);
/* CrsMatrix.begA / jcoA of the fixed structural pattern (int64 like the reference), host out.
 * row_ptr has n_local+1 entries starting at 0; col_idx holds GLOBAL column indices. */
int tfb_get_pattern(tfb_ctx* ctx, int64_t* row_ptr, int64_t* col_idx);

int tfb_mat_create(tfb_ctx* ctx, tfb_mat** out);
void tfb_mat_destroy(tfb_mat* mat);
int tfb_mat_get_values(tfb_mat* mat, double* vals_out);        /* D2H, nnz_local doubles */
int tfb_mat_set_values(tfb_mat* mat, const double* vals_in);   /* H2D */

/* dst = src + alpha * diag(d): the matrix arithmetic TimeIntegration needs on the backend's matrix type
 * (`jacobian(x) - mass / (theta * dt)`, TimeIntegration.py:58); d has one entry per local row.  Every row
 * with d != 0 must have a structural diagonal (all rows with mass do). dst may equal src. */
int tfb_mat_add_diag(tfb_mat* dst, const tfb_mat* src, double alpha, const double* d);
/* Tell the solver that `mat` is J + shift * M with M the mass matrix (what tfb_mat_add_diag produced for TimeIntegration's
 * J - M / (theta dt) or a shifted eigenproblem J - sigma M): the fast-diagonalisation basis is M-orthonormal, so the block
 * preconditioner solves the shifted diffusion operators exactly instead of ignoring the shift. */
int tfb_mat_set_shift(tfb_mat* mat, double shift);

/* Interface.rhs -> Discretization.rhs (Discretization.py:367-390); host in, host out. */
int tfb_rhs(tfb_ctx* ctx, const double* state, double* out);
/* Interface.jacobian -> Discretization.jacobian (:392-415) into `mat`; if rhs_out != NULL the
 * same launch also produces F(x) (fused Jacobian+RHS). */
int tfb_jacobian(tfb_ctx* ctx, const double* state, tfb_mat* mat, double* rhs_out);
/* Interface.mass_matrix -> Discretization.mass_matrix (:417-437): the diagonal, one value per
 * local row (0 for pressure rows). */
int tfb_mass_diag(tfb_ctx* ctx, double* diag_out);

/* Device-resident variants used for kernel-only timing: state already uploaded. */
int tfb_state_upload(tfb_ctx* ctx, const double* state);
int tfb_assemble_resident(tfb_ctx* ctx, tfb_mat* mat, int do_jacobian, int do_rhs);
int tfb_rhs_download(tfb_ctx* ctx, double* out);
/* Device-resident continuation fast path: a corrector iteration of Continuation.newton evaluates
 * rhs(x; mu), rhs(x; mu + delta) and jacobian(x) for the SAME state (Continuation.py:126,145-150).  The host shim keeps
 * the state resident between those calls: tfb_host_checksum (64-bit wrapping sum and xor of the words, multi-threaded --
 * any single changed entry changes it) tells it whether a host vector is the one already in HBM, tfb_upload_count counts
 * the state uploads of a context. */
int tfb_host_checksum(const double* p, int64_t n, uint64_t out[2]);
int64_t tfb_upload_count(tfb_ctx* ctx);
int tfb_sync(tfb_ctx* ctx);

/* CUDA-event timers on the ctx's stream. slot in [0,1040): a timing loop can record a (start, stop) pair per iteration
 * and read them all after the loop, so that no host synchronisation sits between the iterations. */
int tfb_event_record(tfb_ctx* ctx, int slot);
int tfb_event_elapsed_ms(tfb_ctx* ctx, int slot_a, int slot_b, float* ms);
/* Plane boundaries 0 = b[0] < b[1] < ... = nzl of the z-pieces over which tfb_jacobian pipelines upload, assembly and
 * download of a slab of nzl planes (small pieces at both ends, 32-plane pieces in the middle). Returns the number of
 * boundaries written to `out` (< 0: bad arguments). Host-only: no device is touched. */
int tfb_pipe_pieces_of(int nzl, int* out, int cap);
/* write `bytes` of device memory (> L2) to evict the L2 between timed iterations */
int tfb_flush_l2(tfb_ctx* ctx);
/* page-locked host buffers for the host<->device legs of the e2e path */
int tfb_pinned_alloc(size_t bytes, void** out);
int tfb_pinned_free(void* p);
/* number of kernel launches issued by this library since load */
int64_t tfb_launch_count(void);

/* y = A x with host vectors (CrsMatrix.matvec / `jac @ x` of the SciPy backend) */
int tfb_spmv(tfb_mat* mat, const double* x, double* y);

/* average device time (ms) of `reps` launches of y = J x (masked != 0: the velocity-velocity block
 * used by the preconditioner) with resident operands; measurement helper for bench.py */
int tfb_spmv_bench(tfb_mat* mat, int reps, int masked, float* ms_out);

/* Interface.solve (interface/SciPy.py:204-315): solve mat * x = b with the pressure pinned at
 * global row `pressure_row` (<0: no pin) by a preconditioned Krylov method to ||r||/||b|| <= tol.
 * Returns 0 converged, 1 not converged (x holds the best iterate). */
enum { TFB_METHOD_FGMRES = 0, TFB_METHOD_BICGSTAB = 1, TFB_METHOD_IDR = 2 };
/* tfb_solve_opts.precond_flags */
#define TFB_PREC_FP32 1         /* FDM sub-solves of the preconditioner in fp32 (SIMT transforms) */
#define TFB_PREC_NO_JOINT 2     /* do not use the coupled (w, scalar) solve even if tfb_joint_set was called */
#define TFB_PREC_SCALED_MASS 8  /* scaled-mass Schur complement instead of the least-squares commutator */
#define TFB_PREC_TENSOR 16      /* x/y transforms of the FDM solves on the tensor cores (tcgen05, 3xTF32 split, fp32
                                   storage) and Thomas sweeps along z; needs tfb_fdm_set_pencil for the z axis */
typedef struct {
    double tol;
    int32_t maxit, restart;
    int32_t pressure_row;
    int32_t precond;         /* reserved, 0 */
    int32_t verbose;
    int32_t basis_fp32;      /* 1: fp32 storage of the GMRES basis (all arithmetic fp64) */
    int32_t method;          /* TFB_METHOD_* */
    int32_t idr_s;           /* dimension of the IDR(s) shadow space, 0 -> 8 (fixed preconditioner only) */
    int32_t precond_flags;   /* TFB_PREC_* bits */
    int32_t inner_its;       /* inner GMRES steps of the velocity / (velocity, scalar) sub-solve, 0..24 */
    int32_t stall_cycles;    /* restart cycles without a 2x residual reduction before giving up, 0 -> 3 */
} tfb_solve_opts;
typedef struct {
    int32_t iters, converged;
    double relres;
    float setup_ms, solve_ms;
} tfb_solve_info;
int tfb_solve(tfb_mat* mat, const double* b, double* x, const tfb_solve_opts* opts, tfb_solve_info* info);

/* Interface.solve for 2-D grids by a direct method, the counterpart of SciPy.Interface.direct_solve (SciPy.py:204-258:
 * SuperLU, factors cached on the matrix): block-tridiagonal elimination over the grid lines with one dense, pivoted
 * inverse per line (csrc/tfb_direct.cu), fp64, one step of iterative refinement.  The factors stay with `mat` until its
 * values change, so the second solve of a corrector step only substitutes.  info->setup_ms: factorisation time of this
 * call (0 when reused).  Returns 0, 1 (residual above 1e-8: use tfb_solve), <0 on errors (e.g. a singular line block). */
int tfb_direct_solve(tfb_mat* mat, const double* b, double* x, int pressure_row, tfb_solve_info* info);

/* Fast-diagonalisation data of the block preconditioner: for variable `var` and axis `axis`
 * the m x m M-orthonormal eigenvector matrix Q (row-major) and eigenvalues lam of the 1-D pencil
 * (K, M) of that variable's diffusion stencil incl. wall folds; coef = the operator's scalar
 * (c_visc, c_T, c_S; -1 for the pressure Poisson operator D M^-1 G).  Computed by the host
 * (hostprep.fdm_operators) whenever grid or parameters change. */
int tfb_fdm_set(tfb_ctx* ctx, int var, int axis, int m, const double* Q, const double* lam, double coef);
/* The 1-D stencils behind those eigen-decompositions (hostprep._pencil_km): K = tridiag(lower, diag, upper), mass M
 * diagonal, m entries each.  With them the tensor-core path (TFB_PREC_TENSOR) solves the z direction by Thomas sweeps
 * per horizontal mode instead of two dense transforms. */
int tfb_fdm_set_pencil(tfb_ctx* ctx, int var, int axis, int m, const double* lower, const double* diag,
                       const double* upper, const double* mass);
/* A scalar whose diffusion operator is singular (zero-flux on every wall) is pinned at `cell`
 * with diagonal `sign`, like the reference's "fix one salinity value" (Discretization.py:690-701). */
int tfb_fdm_pin(tfb_ctx* ctx, int var, int64_t cell, double sign);
/* Coupled (vertical velocity, scalar) solve of the preconditioner for buoyancy-driven 3-D problems whose
 * scalar is stratified along z (Rayleigh-Benard): replaces "scalar first, velocities after" by one solve of
 * the (velocity, scalar) block in which w and the scalar are solved together, line by line along z, after
 * the x/y fast-diagonalisation transforms.  zops = 8 x nz table (row-major): lower / diagonal / upper / mass
 * of the w stencil along z (nz-1 faces, zero padded), then the same for the scalar
 * (hostprep.joint_z_operators); the vertical couplings are read off each Jacobian.  There is no
 * counterpart in the reference (SciPy.py:131-162 factorises the whole matrix with SuperLU). */
int tfb_joint_set(tfb_ctx* ctx, int wvar, int svar, int nz, const double* zops);
/* diagnostics: the coupled solve alone with host vectors (other variables' rows come back zero);
 * table_out (optional, 12 x nz): the coefficient table after reading the couplings from `mat` */
int tfb_joint_apply(tfb_mat* mat, const double* r, double* z, double* table_out);
/* z = P^-1 r with host vectors (diagnostics / tests of the preconditioner alone) */
int tfb_precond_apply(tfb_mat* mat, const double* r, double* z, int pressure_row);
/* the same with the preconditioner variant chosen by opts (precond_flags, inner_its, pressure_row) */
int tfb_precond_apply_opts(tfb_mat* mat, const double* r, double* z, const tfb_solve_opts* opts);

/* NCCL plumbing for z-slab runs (one process per GPU). */
int tfb_nccl_unique_id(uint8_t id[128]);
int tfb_comm_init(tfb_ctx* ctx, int nranks, int rank, const uint8_t id[128]);

#ifdef __cplusplus
}
#endif
#endif
