// Newton-step linear solver on the device: flexible GMRES with a block preconditioner for the
// velocity-pressure saddle point.  Replaces SciPy.Interface.solve / direct_solve
// (interface/SciPy.py:204-315: SuperLU) for matrices that live in HBM.
//
//   J = [ A  G  B ]   u : velocities        A : convection-diffusion (+ Newton terms)
//       [ D  0  0 ]   p : pressure          G, D : gradient / divergence
//       [ C  0  At]   s : scalars (T, S)    B : buoyancy, C : scalar advection wrt velocity
//
// Preconditioner (right, block upper triangular, "least-squares commutator" Schur complement):
//   s  = At^-1 r_s                              At ~ c_T * Laplacian            (FDM)
//   dp = -Lp^-1 (D M^-1 A M^-1 G) Lp^-1 r_p     Lp = D M^-1 G = Neumann Poisson (FDM, twice)
//   u  = Ah^-1 (r_u - B s - G dp)               Ah ~ c_visc * vector Laplacian  (FDM)
// Every sub-solve is a fast-diagonalisation (FDM) solve: the operators are sums of Kronecker
// products of 1-D tridiagonal stencils on the tensor-product grid, so with the generalized
// eigen-decompositions of the 1-D pencils (host, numpy, once per grid) a solve is three dense
// transforms along x,y,z, a pointwise scaling and three transforms back -- exact for the
// diffusion operators incl. the wall folds, stretched grids included, no inner iteration.
// The pressure is pinned at the reference's pressure row (SciPy.py:95-106,212-216): row -> -1
// on the diagonal, column dropped; this is applied on the fly inside the SpMV kernels.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "tfb_internal.h"

int tfb_allreduce_sum(tfb_ctx* c, double* d_buf, int count);
int tfb_alltoallv_bytes(tfb_ctx* c, const void* send, const long long* scount, const long long* sdispl, void* recv,
                        const long long* rcount, const long long* rdispl, int elem_bytes);

#include "tfb_joint.h"
#include "tfb_fdm_tc.cuh"
#define TFB_MAXVAR 6
// tensor-core plane transforms (tfb_fdm_tc.cuh): K-chunk / pipeline depth of the kernel instantiation in use
#define TFB_TC_KC 32
#define TFB_TC_STAGES 3

struct FdmVar {
    bool present = false;
    int m[3] = {0, 0, 0};         // active extent per axis (n-1 along the velocity's own axis)
    double* Q[3] = {nullptr, nullptr, nullptr};   // m x m, row-major, M-orthonormal eigenvectors
    double* lam[3] = {nullptr, nullptr, nullptr}; // m generalized eigenvalues
    float* Qf[3] = {nullptr, nullptr, nullptr};   // fp32 copies for the single-precision preconditioner
    float* lamf[3] = {nullptr, nullptr, nullptr};
    double coef = 1.0;
    double maxden = 0.0;
    long long pin_cell = -1;      // singular (all-Neumann) scalar operators are pinned at one cell
    double pin_sign = 1.0;        // diagonal of the pinned row (+1 identity, -1 for the AMOC salinity pin)
    // tensor-core path (tfb_fdm_tc.cuh): pre-split, pre-swizzled copies of Q^T (forward) and Q (backward) for the
    // x and y axes, the 1-D stencils (lower, diagonal, upper of K, then M; 4 x n doubles) of every axis, and the
    // Thomas factors of the z direction for the current (eigenvalues, coefficient, shift)
    float* tcA[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [axis][0 forward | 1 backward]
    double* pencil[3] = {nullptr, nullptr, nullptr};
    int pencil_n[3] = {0, 0, 0};
    float* th_inv = nullptr;
    float* th_cp = nullptr;
    long long th_cap = 0;         // floats allocated per array
    bool th_dirty = true;
    double shift = 0.0;           // operator + shift * mass (time stepping, shifted eigenproblems)
    std::vector<double> pencil_h[3];   // host copies of the pencil tables
    // z-slabs: scaled spikes of the local block, this rank's / everybody's reduced-system coefficients, inverse rows
    float* sp_v = nullptr;
    float* sp_w = nullptr;
    float* sp_coef = nullptr;     // [4][modes]
    float* sp_coef_all = nullptr; // [G][4][modes]
    float* sp_weights = nullptr;  // [2][2G][modes]
    int mz_local = 0;
};

// Compact copy of the entries of J with (row variable, column variable) in given masks: the
// gradient G, divergence D and buoyancy B blocks hold 2..6 entries per row, so multiplying with
// them through the full 17-entries-per-row matrix would waste ~90 % of the traffic.
struct SubCsr {
    unsigned rowmask = 0, colmask = 0;
    int* row_ptr = nullptr;   // n_local + 1
    int* col = nullptr;
    int* src = nullptr;       // position in the parent values array
    double* vals = nullptr;
    int nnz = 0;
    uint64_t version = ~0ull; // matrix version the values were copied from
    const tfb_mat* owner = nullptr;
};

struct tfb_solver_state {
    FdmVar var[TFB_MAXVAR];
    SubCsr subG, subD, subB, subC;
    // coupled (vertical velocity, scalar) solve of the Rayleigh-Benard preconditioner (tfb_joint.h)
    bool joint_ready = false;     // tfb_joint_set was called
    bool joint_on = false;        // this solve uses it
    int joint_w = -1, joint_s = -1;
    double* d_jz = nullptr;       // TFB_JZ_ROWS x nz coefficient table
    double* jbuf[2] = {nullptr, nullptr};   // 2 x ncell each: [w planes | T planes]
    double* jab = nullptr;        // 4 x ncell: the two upper bands of the line factorisations
    float* jfac = nullptr;        // tensor-core path: 4 x (2 nz - 1) x modes factors of every mode (tfb_joint_factor_line)
    const tfb_mat* jfac_owner = nullptr;
    uint64_t jfac_version = ~0ull;
    const tfb_mat* jz_owner = nullptr;
    uint64_t jz_version = ~0ull;
    int sub_prow = -2;
    // 'Schur Complement': 'Scaled Mass' -- dp = gamma * r_p / (cell volume); gamma is re-estimated for every matrix
    bool schur_mass = false;
    double gamma = 0.0, gamma_rho = 0.0;
    const tfb_mat* gam_owner = nullptr;
    uint64_t gam_version = ~0ull;
    double* d_mass = nullptr;     // velocity mass diagonal (LSC scaling), n_local
    double* comp[3] = {};         // SoA work arrays, ncell each
    double* vec[8] = {};          // interleaved work vectors, n_local each, one plane of slack before and after
    double* vec_base[8] = {};
    long long dv_stride = 0;      // > 0: d_V holds dv_count vectors of this stride, each preceded by a plane of slack (idr_run)
    int dv_count = 0;
    double* d_scal = nullptr;     // small device scalars
    // z-slab runs: ghosted copy of a vector, pencil-layout work arrays, all-to-all staging
    double* xg = nullptr;
    double* pen[2] = {nullptr, nullptr};
    double* sbuf = nullptr;
    double* rbuf = nullptr;
    int j0s[TFB_MAX_RANKS + 1] = {};   // y-chunk of every rank in the pencil layout
    long long a2a_cnt_slab[TFB_MAX_RANKS] = {}, a2a_dsp_slab[TFB_MAX_RANKS] = {};   // slab side (packed by y-chunk)
    long long a2a_cnt_pen[TFB_MAX_RANKS] = {}, a2a_dsp_pen[TFB_MAX_RANKS] = {};     // pencil side (planes of each rank)
    bool dist_ready = false;
    bool precond_single = false;  // apply the FDM sub-solves in fp32 (FGMRES keeps the outer iteration exact)
    bool precond_tc = false;      // x/y transforms on the tensor cores (3xTF32), Thomas sweeps along z
    float* sp_send = nullptr;     // z-slabs: interface values of the local Thomas solutions, [TFB_MAXVAR][2][modes]
    float* sp_recv = nullptr;     // [G] x that
    float* tc32[2] = {nullptr, nullptr};   // fp32 SoA work arrays, TFB_MAXVAR x ncell (pencil-sized with z-slabs) each
    long long tc32_cap = 0;
    // fused head / tail of the scaled-mass preconditioner: two-slot copy of the gradient block, pressure update
    double* d_ih[3] = {nullptr, nullptr, nullptr};   // reciprocal cell widths per axis
    float* gell = nullptr;        // [dim][2][ncell]
    float* dp32 = nullptr;        // ncell + one plane
    int* gell_misfit = nullptr;
    bool gell_ok = false;
    const tfb_mat* gell_owner = nullptr;
    uint64_t gell_version = ~0ull;
    // velocity sub-solve: inner GMRES on the convection-diffusion block, preconditioned by the FDM solve
    int inner_its = 0;            // 0: one FDM (diffusion-only) solve
    double inner_tol = 1e-2;
    double* d_Vi = nullptr;       // (inner_cap + 1) x n
    double* d_Zi = nullptr;       // inner_cap x n
    int inner_cap = 0;
    long long inner_total = 0;    // inner iterations of the current solve (diagnostic)
    double* d_V = nullptr;        // Krylov basis  (m+1) x n  (fp64, or fp32 when basis_single)
    bool basis_single = false;
    double* d_Z = nullptr;        // preconditioned basis  m x n
    double* d_h = nullptr;        // dot products
    int cap = 0;                  // allocated Krylov dimension
};

void tfb_solver_free(tfb_solver_state* s) {
    if (!s) return;
    for (auto& v : s->var) {
        for (int a = 0; a < 3; a++) { cudaFree(v.Q[a]); cudaFree(v.lam[a]); cudaFree(v.Qf[a]); cudaFree(v.lamf[a]); cudaFree(v.pencil[a]); }
        for (int a = 0; a < 2; a++) { cudaFree(v.tcA[a][0]); cudaFree(v.tcA[a][1]); }
        cudaFree(v.th_inv); cudaFree(v.th_cp);
        cudaFree(v.sp_v); cudaFree(v.sp_w); cudaFree(v.sp_coef); cudaFree(v.sp_coef_all); cudaFree(v.sp_weights);
    }
    cudaFree(s->sp_send); cudaFree(s->sp_recv);
    for (auto& p : s->tc32) cudaFree(p);
    cudaFree(s->gell); cudaFree(s->dp32); cudaFree(s->gell_misfit);
    for (auto p : s->d_ih) cudaFree(p);
    for (SubCsr* q : {&s->subG, &s->subD, &s->subB, &s->subC}) { cudaFree(q->row_ptr); cudaFree(q->col); cudaFree(q->src); cudaFree(q->vals); }
    cudaFree(s->d_jz); cudaFree(s->jbuf[0]); cudaFree(s->jbuf[1]); cudaFree(s->jab); cudaFree(s->jfac);
    cudaFree(s->d_mass);
    for (auto p : s->comp) cudaFree(p);
    for (auto p : s->vec_base) cudaFree(p);
    cudaFree(s->xg); cudaFree(s->pen[0]); cudaFree(s->pen[1]); cudaFree(s->sbuf); cudaFree(s->rbuf);
    cudaFree(s->d_scal); cudaFree(s->d_V); cudaFree(s->d_Z); cudaFree(s->d_h); cudaFree(s->d_Vi); cudaFree(s->d_Zi);
    delete s;
}

static tfb_solver_state* solver_of(tfb_ctx* c) {
    if (!c->solver) c->solver = new tfb_solver_state();
    return c->solver;
}

// ------------------------------------------------------------------------------------
// SpMV on the fixed pattern.  8 lanes per row (rows hold 4..23 entries), shuffle reduction.
// rowmask/colmask select variables (bit v = variable v of a cell); prow = pinned pressure row:
// the row reads -x[prow] (only in the full operator), its column is skipped everywhere.
// ------------------------------------------------------------------------------------
template <bool MASKED>
__global__ void __launch_bounds__(256)
tfb_spmv_kernel(long long nrows, int dof, const int* __restrict__ row_ptr, const int* __restrict__ col,
                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                int prow_local, int prow, unsigned rowmask, unsigned colmask, const double* __restrict__ rowscale) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gt >> 3;
    const int lane = threadIdx.x & 7;
    if (row >= nrows) return;
    double s = 0.0;
    const int rv = (int)(row % dof);
    const bool active = !MASKED || ((rowmask >> rv) & 1u);
    if (active && row != prow_local) {
        const int e0 = row_ptr[row], e1 = row_ptr[row + 1];
        for (int e = e0 + lane; e < e1; e += 8) {
            const int cidx = col[e];
            bool take = cidx != prow;
            if (MASKED) take = take && ((colmask >> (cidx % dof)) & 1u);
            if (take) s += vals[e] * x[cidx];
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (lane == 0) {
        if (!MASKED && row == prow_local) s = -x[prow];
        if (MASKED && rowscale) s = active ? s / rowscale[row] : 0.0;
        y[row] = s;
    }
}

// local row index of the (global) pinned row, -1 if another rank owns it
static inline int local_prow(const tfb_ctx* c, int prow) {
    if (prow < 0) return -1;
    const long long l = (long long)prow - c->row0;
    return (l >= 0 && l < c->n_local) ? (int)l : -1;
}

// Column indices are GLOBAL.  Single GPU: x itself.  z-slabs: x is copied between two ghost
// planes, the halos are exchanged (NCCL) and the kernels index a pointer shifted by the first
// owned row, so that global columns address the ghosted copy.
static int dist_setup(tfb_ctx* c);
// Work vectors are allocated with one plane of slack on either side (ensure_buffers, idr_run), so on z-slabs the halo
// planes of x are received next to it and no ghosted copy of the vector is made.
static bool ghost_capable(const tfb_ctx* c, const double* x) {
    const tfb_solver_state* s = c->solver;
    for (const double* v : s->vec) if (x == v) return true;
    if (s->dv_stride > 0 && s->d_V) {
        const double* first = s->d_V + c->plane_rows;
        if (x >= first && x < first + (size_t)s->dv_stride * s->dv_count && (size_t)(x - first) % (size_t)s->dv_stride == 0) return true;
    }
    return false;
}
static int ghosted(tfb_ctx* c, const double* x, const double** xs) {
    if (c->nranks == 1) { *xs = x; return 0; }
    if (dist_setup(c)) return -1;
    tfb_solver_state* s = c->solver;
    if (ghost_capable(c, x)) {
        if (tfb_halo_exchange(c, const_cast<double*>(x) - c->plane_rows)) return -1;
        *xs = x - c->row0;
        return 0;
    }
    TFB_CUDA(cudaMemcpyAsync(s->xg + c->plane_rows, x, sizeof(double) * c->n_local, cudaMemcpyDeviceToDevice, c->stream));
    if (tfb_halo_exchange(c, s->xg)) return -1;
    *xs = s->xg + c->plane_rows - c->row0;
    return 0;
}

static int spmv(tfb_ctx* c, tfb_mat* m, const double* x, double* y, int prow, unsigned rowmask = 0, unsigned colmask = 0,
                const double* rowscale = nullptr) {
    const long long threads = c->n_local * 8;
    const unsigned nb = (unsigned)((threads + 255) / 256);
    static int use_structured = -1;
    if (use_structured < 0) { const char* e = getenv("TFB_SPMV_CSR"); use_structured = !(e && e[0] == '1'); }
    // planes of x that exist: the slab plus one halo plane on either side, clipped to the domain
    const int kv0 = std::max(0, c->desc.k0 - 1), kv1 = std::min(c->desc.nz, c->desc.k1 + 1);
    const bool marching = use_structured && c->desc.dim == 3 && c->desc.nz > 1;
    if (c->nranks > 1 && marching && c->nzl >= 4 && ghost_capable(c, x) && tfb_overlap_enabled(1)) {
        // z-slabs: the halo planes of x travel on a side stream while the interior planes are multiplied; the two planes
        // next to the halo follow when it has landed
        if (dist_setup(c) || tfb_comm_stream(c)) return -1;
        TFB_CUDA(cudaEventRecord(c->ev_comm[0], c->stream));
        TFB_CUDA(cudaStreamWaitEvent(c->s_comm, c->ev_comm[0], 0));
        if (tfb_halo_exchange_on(c, const_cast<double*>(x) - c->plane_rows, c->s_comm)) return -1;
        TFB_CUDA(cudaEventRecord(c->ev_comm[1], c->s_comm));
        const double* xs = x - c->row0;
        c->win0 = 1; c->win1 = c->nzl - 1;
        int rc = tfb_spmv_structured(c, m, xs, kv0, kv1, y, prow, rowmask, colmask, rowscale);
        TFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm[1], 0));
        c->win0 = 0; c->win1 = 1;
        if (rc == 0) rc = tfb_spmv_structured(c, m, xs, kv0, kv1, y, prow, rowmask, colmask, rowscale);
        c->win0 = c->nzl - 1; c->win1 = c->nzl;
        if (rc == 0) rc = tfb_spmv_structured(c, m, xs, kv0, kv1, y, prow, rowmask, colmask, rowscale);
        c->win0 = 0; c->win1 = -1;
        if (rc <= 0) return rc;
    }
    const double* xs = nullptr;
    if (ghosted(c, x, &xs)) return -1;
    if (use_structured) {   // true 3-D grids: structured kernel that never reads the column indices
        const int rc = tfb_spmv_structured(c, m, xs, kv0, kv1, y, prow, rowmask, colmask, rowscale);
        if (rc <= 0) return rc;
    }
    const int pl = local_prow(c, prow);
    if (rowmask)
        tfb_spmv_kernel<true><<<nb, 256, 0, c->stream>>>(c->n_local, c->desc.dof, c->d_row_ptr, c->d_col, m->d_vals, xs, y,
                                                        pl, prow, rowmask, colmask, rowscale);
    else
        tfb_spmv_kernel<false><<<nb, 256, 0, c->stream>>>(c->n_local, c->desc.dof, c->d_row_ptr, c->d_col, m->d_vals, xs, y,
                                                         pl, prow, 0u, 0u, nullptr);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

static int sub_build(tfb_ctx* c, SubCsr& S, int prow, unsigned rowmask, unsigned colmask);
// ---- sub-matrix extraction (structure once per pattern and pin; values once per Jacobian) ----
__global__ void k_sub_count(long long nrows, int dof, const int* __restrict__ row_ptr, const int* __restrict__ col,
                            int prow_local, int prow, unsigned rowmask, unsigned colmask, int* __restrict__ counts) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int cnt = 0;
    if (((rowmask >> (row % dof)) & 1u) && row != prow_local)
        for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) {
            const int cidx = col[e];
            if (cidx != prow && ((colmask >> (cidx % dof)) & 1u)) cnt++;
        }
    counts[row] = cnt;
}
__global__ void k_sub_fill(long long nrows, int dof, const int* __restrict__ row_ptr, const int* __restrict__ col,
                           int prow_local, int prow, unsigned rowmask, unsigned colmask, const int* __restrict__ sub_ptr,
                           int* __restrict__ sub_col, int* __restrict__ sub_src) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    if (!(((rowmask >> (row % dof)) & 1u) && row != prow_local)) return;
    int pos = sub_ptr[row];
    for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) {
        const int cidx = col[e];
        if (cidx != prow && ((colmask >> (cidx % dof)) & 1u)) { sub_col[pos] = cidx; sub_src[pos] = e; pos++; }
    }
}
__global__ void k_sub_gather(int nnz, const int* __restrict__ src, const double* __restrict__ vals, double* __restrict__ out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) out[e] = vals[src[e]];
}
// y = S x (optionally / rowscale) : one thread per row, rows are 0..6 entries long
__global__ void k_sub_spmv(long long nrows, const int* __restrict__ row_ptr, const int* __restrict__ col,
                           const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                           const double* __restrict__ rowscale) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double s = 0.0;
    for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) s += vals[e] * x[col[e]];
    if (rowscale) s /= rowscale[row];
    y[row] = s;
}

// ------------------------------------------------------------------------------------
// vector kernels
// ------------------------------------------------------------------------------------
__global__ void k_axpy(long long n, double a, const double* __restrict__ x, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void k_scale_to(long long n, const double* __restrict__ scal, int idx, const double* __restrict__ x, double* __restrict__ y) {
    const double a = 1.0 / scal[idx];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = a * x[i];
}
__global__ void k_rowdiv(long long n, const double* __restrict__ d, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] /= d[i];
}
__global__ void k_sub(long long n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = a[i] - b[i];
}


// ---- single-pass orthogonalisation kernels: every basis vector and w are read exactly once ----
// h[v] += sum_i V[v][i] * w[i] for all v < nv.  A CTA owns a tile of rows, keeps its w values in
// registers and walks over the basis vectors; per-warp partial sums are parked in shared memory.
#define TFB_ORTH_ROWS 4
template <class BT>
__global__ void __launch_bounds__(256) k_all_dots(long long n, const BT* __restrict__ V, long long ld, int nv,
                                                  const double* __restrict__ w, double* __restrict__ out) {
    extern __shared__ double part[];   // [8 warps][nv]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int v = threadIdx.x; v < 8 * nv; v += 256) part[v] = 0.0;
    __syncthreads();
    const long long tile = (long long)256 * TFB_ORTH_ROWS;
    for (long long base = (long long)blockIdx.x * tile; base < n; base += (long long)gridDim.x * tile) {
        double wr[TFB_ORTH_ROWS];
#pragma unroll
        for (int r = 0; r < TFB_ORTH_ROWS; r++) {
            const long long i = base + r * 256 + threadIdx.x;
            wr[r] = i < n ? w[i] : 0.0;
        }
        for (int v = 0; v < nv; v++) {
            const BT* vp = V + (long long)v * ld;
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < TFB_ORTH_ROWS; r++) {
                const long long i = base + r * 256 + threadIdx.x;
                if (i < n) a += (double)vp[i] * wr[r];
            }
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) part[warp * nv + v] += a;
        }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += 256) {
        double a = 0.0;
        for (int wv = 0; wv < 8; wv++) a += part[wv * nv + v];
        atomicAdd(&out[v], a);
    }
}
// w += sign * sum_v h[v] V[v]; optionally accumulates |w_new|^2 into *nrm2
template <class BT>
__global__ void __launch_bounds__(256) k_all_axpy(long long n, const BT* __restrict__ V, long long ld, int nv,
                                                  const double* __restrict__ h, double sign, double* __restrict__ w,
                                                  double* __restrict__ nrm2, const double* __restrict__ base = nullptr,
                                                  double bscale = 1.0, double wscale = 1.0) {
    extern __shared__ double hs[];
    for (int v = threadIdx.x; v < nv; v += 256) hs[v] = h[v];
    __syncthreads();
    double loc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int v = 0; v < nv; v++) a += hs[v] * (double)V[(long long)v * ld + i];
        // general form w = wscale * w + bscale * base + sign * V h  (wscale == 0: w is write-only)
        double wn = sign * a;
        if (wscale != 0.0) wn += wscale * w[i];
        if (base) wn += bscale * base[i];
        w[i] = wn;
        loc += wn * wn;
    }
    if (nrm2) {
        for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
        __shared__ double red[8];
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int wv = 0; wv < 8; wv++) t += red[wv];
            atomicAdd(nrm2, t);
        }
    }
}

// y (basis type) = x / scal[idx];  x64 = (double) basis vector
template <class BT>
__global__ void k_store_scaled(long long n, const double* __restrict__ scal, int idx, const double* __restrict__ x, BT* __restrict__ y) {
    const double a = 1.0 / scal[idx];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = (BT)(a * x[i]);
}
__global__ void k_widen(long long n, const float* __restrict__ x, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = (double)x[i];
}

static inline unsigned vec_blocks(long long n);
static int sub_build(tfb_ctx* c, SubCsr& S, int prow, unsigned rowmask, unsigned colmask) {
    const long long n = c->n_local;
    S.rowmask = rowmask; S.colmask = colmask;
    cudaFree(S.row_ptr); cudaFree(S.col); cudaFree(S.src); cudaFree(S.vals);
    S.row_ptr = S.col = S.src = nullptr; S.vals = nullptr;
    int* counts = nullptr;
    TFB_CUDA(cudaMalloc(&counts, sizeof(int) * (n + 1)));
    TFB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (n + 1), c->stream));
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_sub_count<<<nb, 256, 0, c->stream>>>(n, c->desc.dof, c->d_row_ptr, c->d_col, local_prow(c, prow), prow, rowmask, colmask, counts);
    TFB_LAUNCHED();
    // exclusive scan on the host side of a small int array is avoided: reuse CUB through a tiny kernel-free path
    std::vector<int> h(n + 1);
    TFB_CUDA(cudaMemcpyAsync(h.data(), counts, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    long long acc = 0;
    for (long long r = 0; r <= n; r++) { const int v = h[r]; h[r] = (int)acc; acc += v; }
    S.nnz = (int)acc;
    TFB_CUDA(cudaMalloc(&S.row_ptr, sizeof(int) * (n + 1)));
    TFB_CUDA(cudaMemcpyAsync(S.row_ptr, h.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, c->stream));
    const size_t cap = (size_t)std::max(S.nnz, 1);
    TFB_CUDA(cudaMalloc(&S.col, sizeof(int) * cap));
    TFB_CUDA(cudaMalloc(&S.src, sizeof(int) * cap));
    TFB_CUDA(cudaMalloc(&S.vals, sizeof(double) * cap));
    k_sub_fill<<<nb, 256, 0, c->stream>>>(n, c->desc.dof, c->d_row_ptr, c->d_col, local_prow(c, prow), prow, rowmask, colmask, S.row_ptr, S.col, S.src);
    TFB_LAUNCHED();
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(counts);
    S.version = ~0ull;
    S.owner = nullptr;
    return 0;
}
static int joint_refresh(tfb_ctx* c, tfb_mat* m);
static int schur_gamma_refresh(tfb_ctx* c, tfb_mat* m, int prow);
static int gell_refresh(tfb_ctx* c, tfb_mat* m);
static int sub_refresh(tfb_ctx* c, tfb_mat* m, int prow) {
    tfb_solver_state* s = c->solver;
    const int dof = c->desc.dof, dim = c->desc.dim;
    // J + shift * M: the diffusion sub-solves of every variable with mass carry the shift (exact in the M-orthonormal basis)
    for (int v = 0; v < dof; v++) {
        if (v == dim) continue;
        FdmVar& f = s->var[v];
        if (f.shift != m->shift) { f.shift = m->shift; f.th_dirty = true; }
    }
    const unsigned velmask = (1u << dim) - 1u, pmask = 1u << dim, smask = ((1u << dof) - 1u) & ~(velmask | pmask);
    if (s->sub_prow != prow || !s->subG.row_ptr) {
        if (sub_build(c, s->subG, prow, velmask, pmask)) return -1;
        if (sub_build(c, s->subD, prow, pmask, velmask)) return -1;
        if (smask && sub_build(c, s->subB, prow, velmask, smask)) return -1;
        if (smask && s->joint_ready && sub_build(c, s->subC, prow, smask, velmask)) return -1;
        s->sub_prow = prow;
    }
    if (smask && s->joint_ready && !s->subC.row_ptr && sub_build(c, s->subC, prow, smask, velmask)) return -1;
    for (SubCsr* q : {&s->subG, &s->subD, &s->subB, &s->subC}) {
        if (!q->row_ptr || q->nnz == 0) continue;
        if (q->owner == m && q->version == m->version) continue;
        k_sub_gather<<<vec_blocks(q->nnz), 256, 0, c->stream>>>(q->nnz, q->src, m->d_vals, q->vals);
        TFB_LAUNCHED();
        q->owner = m; q->version = m->version;
    }
    if (s->joint_ready && joint_refresh(c, m)) return -1;
    if (s->schur_mass && s->precond_tc && gell_refresh(c, m)) return -1;
    if (s->schur_mass && schur_gamma_refresh(c, m, prow)) return -1;
    TFB_CUDA(cudaGetLastError());
    return 0;
}
static int sub_spmv(tfb_ctx* c, const SubCsr& S, const double* x, double* y, const double* rowscale = nullptr) {
    const double* xs = nullptr;
    if (ghosted(c, x, &xs)) return -1;
    k_sub_spmv<<<(unsigned)((c->n_local + 255) / 256), 256, 0, c->stream>>>(c->n_local, S.row_ptr, S.col, S.vals, xs, y, rowscale);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

static inline unsigned vec_blocks(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148 * 16); }

// ------------------------------------------------------------------------------------
// FDM building blocks (SoA arrays of one variable, dims (nz, ny, nx), x fastest)
// ------------------------------------------------------------------------------------
template <class FT>
__global__ void k_deinterleave(long long ncell, int dof, int v, const double* __restrict__ x, FT* __restrict__ c) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) c[i] = (FT)x[i * dof + v];
}
template <class FT>
__global__ void k_interleave(long long ncell, int dof, int v, const FT* __restrict__ c, double* __restrict__ x, double sign) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) x[i * dof + v] = sign * (double)c[i];
}

// C(b, m, n) = sum_k A(b, m, k) * Q(k, n)   [TRANS: Q(n, k)],   A/C element (b,m,k) at b*sb + m*sm + k*sk.
// 64 x 64 output tile per CTA, 16-deep k-slabs in shared memory, 4 x 4 outputs per thread, fp64 FMA.
template <bool TRANS, class FT>
__global__ void __launch_bounds__(256)
k_axis_gemm(const FT* __restrict__ A, FT* __restrict__ C, const FT* __restrict__ Q, int ldq,
            int M, int K, int N, long long sm, long long sk, long long sb) {
    __shared__ FT As[16][64 + 1];
    __shared__ FT Qs[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const FT* Ab = A + (long long)blockIdx.z * sb;
    FT* Cb = C + (long long)blockIdx.z * sb;
    FT acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = (FT)0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        // A tile: 64 (m) x 16 (k); pick the thread->element map that follows the unit stride
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            int mm, kk;
            if (sk == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < K) ? Ab[gm * sm + gk * sk] : (FT)0;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            int kk, nn;
            if (TRANS) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
            const int gk = k0 + kk, gn = n0 + nn;
            FT q = (FT)0;
            if (gk < K && gn < N) q = TRANS ? Q[(long long)gn * ldq + gk] : Q[(long long)gk * ldq + gn];
            Qs[kk][nn] = q;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            FT a[4], q[4];
#pragma unroll
            for (int r = 0; r < 4; r++) a[r] = As[kk][ty * 4 + r];
#pragma unroll
            for (int r = 0; r < 4; r++) q[r] = Qs[kk][tx * 4 + r];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int s = 0; s < 4; s++) acc[r][s] += a[r] * q[s];
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const int gm = m0 + ty * 4 + r, gn = n0 + tx * 4 + s;
            if (gm < M && gn < N) Cb[gm * sm + gn * sk] = acc[r][s];
        }
}

// t /= coef*(lx[i]+ly[j]+lz[k]) on an array of extents (ex, ey, ez) that is a window of the grid
// starting at global (0, jofs, kofs); m* = active global extents
template <class FT>
__global__ void k_fdm_scale(int ex, int ey, int ez, int jofs, int kofs, int mx, int my, int mz, const double* __restrict__ lx,
                            const double* __restrict__ ly, const double* __restrict__ lz, double coef, double shift, double thresh,
                            FT* __restrict__ t) {
    const long long ncell = (long long)ex * ey * ez;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % ex), j = jofs + (int)((c / ex) % ey), k = kofs + (int)(c / ((long long)ex * ey));
        if (i < mx && j < my && k < mz) {
            const double den = coef * (lx[i] + ly[j] + (lz ? lz[k] : 0.0)) + shift;
            t[c] = fabs(den) > thresh ? (FT)((double)t[c] / den) : (FT)0;
        }
    }
}

// wall-normal boundary unknowns (index >= m along the own axis) have the row -1 * u
template <class FT>
__global__ void k_fdm_walls(int nx, int ny, int nzl, int kofs, int mx, int my, int mz, const FT* __restrict__ in, FT* __restrict__ out) {
    const long long ncell = (long long)nx * ny * nzl;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % nx), j = (int)((c / nx) % ny), k = kofs + (int)(c / ((long long)nx * ny));
        if (i >= mx || j >= my || k >= mz) out[c] = -in[c];
    }
}

// slab layout [kl][j][i]  <->  all-to-all buffer packed by destination y-chunk: [r][kl][jj][i]
struct TfbChunks { int n; int j0[TFB_MAX_RANKS + 1]; long long dsp[TFB_MAX_RANKS]; };
template <bool PACK, class FT>
__global__ void k_a2a_pack(int nx, int ny, int nzl, TfbChunks ch, FT* __restrict__ slab, FT* __restrict__ buf) {
    const long long ncell = (long long)nx * ny * nzl;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % nx), j = (int)((c / nx) % ny), kl = (int)(c / ((long long)nx * ny));
        int r = 0;
        while (r + 1 < ch.n && j >= ch.j0[r + 1]) r++;
        const int cy = ch.j0[r + 1] - ch.j0[r];
        const long long b = ch.dsp[r] + ((long long)kl * cy + (j - ch.j0[r])) * nx + i;
        if (PACK) buf[b] = slab[c]; else slab[c] = buf[b];
    }
}

// pinned Neumann Poisson: make the rhs compatible / shift the solution so that cell `pc` is the pin
template <class FT>
__global__ void k_sum(long long n, const FT* __restrict__ x, double* __restrict__ out) {
    double a = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); w++) s += red[w];
        atomicAdd(out, s);
    }
}
template <class FT>
__global__ void k_pin_rhs(FT* rp, long long pc, double* scal) {   // scal[0] = sum(rp) on entry
    const double r0 = (double)rp[pc];
    scal[1] = r0;
    rp[pc] = (FT)(-(scal[0] - r0));
}
template <class FT>
__global__ void k_pin_shift(long long n, FT* q, long long pc, const double* __restrict__ scal, double* qpin, double sign) {
    const double q0 = *qpin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        q[i] = (i == pc) ? (FT)(sign * scal[1]) : (FT)((double)q[i] - q0);
}
template <class FT>
__global__ void k_copy1(const FT* src, double* dst) { *dst = (double)*src; }

// One dense transform along an axis (fp64 / fp32 SIMT path; the large 3-D solves use tfb_fdm_tc.cuh instead).
template <class FT>
static int axis_gemm(tfb_ctx* c, bool trans, const FT* A, FT* C, const FT* Q, int ldq, int M, int K, int N,
                     long long sm, long long sk, long long sb, int batches) {
    dim3 grid((M + 63) / 64, (N + 63) / 64, batches);
    if (trans) k_axis_gemm<true, FT><<<grid, 256, 0, c->stream>>>(A, C, Q, ldq, M, K, N, sm, sk, sb);
    else k_axis_gemm<false, FT><<<grid, 256, 0, c->stream>>>(A, C, Q, ldq, M, K, N, sm, sk, sb);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// z-slab runs: sizes of the all-to-all between the slab layout and the pencil layout (all z, a
// chunk of y), and the work arrays.  Called once after tfb_comm_init.
static int dist_setup(tfb_ctx* c) {
    tfb_solver_state* s = c->solver;
    if (s->dist_ready || c->nranks == 1) return 0;
    const int G = c->nranks, nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz;
    TFB_CHECK(ny >= G, "fewer y-lines than ranks");
    const int base = ny / G, rem = ny % G;
    for (int r = 0; r <= G; r++) s->j0s[r] = r * base + std::min(r, rem);
    const int cyme = s->j0s[c->rank + 1] - s->j0s[c->rank];
    long long ds = 0;
    for (int r = 0; r < G; r++) {
        const int cy = s->j0s[r + 1] - s->j0s[r];
        s->a2a_cnt_slab[r] = (long long)c->nzl * cy * nx;      // my planes, rank r's y-chunk
        s->a2a_dsp_slab[r] = ds;
        ds += s->a2a_cnt_slab[r];
        const int nzr = c->slab_k0[r + 1] - c->slab_k0[r];
        s->a2a_cnt_pen[r] = (long long)nzr * cyme * nx;        // rank r's planes, my y-chunk
        s->a2a_dsp_pen[r] = (long long)c->slab_k0[r] * cyme * nx;
    }
    const size_t ncell = (size_t)c->n_local / c->desc.dof, npen = (size_t)nz * cyme * nx;
    TFB_CUDA(cudaMalloc(&s->xg, sizeof(double) * (size_t)c->plane_rows * (c->nzl + 2)));
    TFB_CUDA(cudaMemset(s->xg, 0, sizeof(double) * (size_t)c->plane_rows * (c->nzl + 2)));
    TFB_CUDA(cudaMalloc(&s->pen[0], sizeof(double) * npen));
    TFB_CUDA(cudaMalloc(&s->pen[1], sizeof(double) * npen));
    TFB_CUDA(cudaMalloc(&s->sbuf, sizeof(double) * ncell));
    TFB_CUDA(cudaMalloc(&s->rbuf, sizeof(double) * ncell));
    s->dist_ready = true;
    return 0;
}

static bool tc_ready(const tfb_ctx* c, int v);
static int fdm_solve_tc(tfb_ctx* c, int nv, const int* vars, float* const* in, float* const* mid, float* const* out);
// out = Op_v^-1 in  (SoA arrays of the local slab; `in` is preserved, tmp is scratch)
template <class FT> static inline FT* const* fdm_q(const FdmVar& f);
template <> inline double* const* fdm_q<double>(const FdmVar& f) { return f.Q; }
template <> inline float* const* fdm_q<float>(const FdmVar& f) { return f.Qf; }

template <class FT>
static int fdm_solve(tfb_ctx* c, int v, FT* in, FT* tmp, FT* out) {
    tfb_solver_state* s = c->solver;
    const FdmVar& f = s->var[v];
    FT* const* Q = fdm_q<FT>(f);
    TFB_CHECK(f.present, "FDM operator missing for a variable (tfb_fdm_set)");
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, nzl = c->nzl, k0 = c->desc.k0;
    const long long ncell = (long long)nx * ny * nzl;
    const bool three = c->desc.dim == 3 && nz > 1;
    const int mx = f.m[0], my = f.m[1], mz = three ? f.m[2] : nz;
    const double thresh = 1e-12 * fabs(f.coef) * f.maxden;
    if constexpr (sizeof(FT) == 4) {
        if (s->precond_tc && tc_ready(c, v)) {
            // tensor-core path: x/y transforms per plane (3xTF32), Thomas sweeps along z
            float* iv[1] = {in};
            float* mv[1] = {tmp};
            float* ov[1] = {out};
            if (fdm_solve_tc(c, 1, &v, iv, mv, ov)) return -1;
            if (mx < nx || my < ny || mz < nz) {
                k_fdm_walls<FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(nx, ny, nzl, k0, mx, my, mz, in, out);
                TFB_LAUNCHED();
            }
            TFB_CUDA(cudaGetLastError());
            return 0;
        }
    }
    // forward: x, y on the local planes
    if (axis_gemm(c, false, in, tmp, Q[0], mx, ny * nzl, mx, mx, nx, 1, 0, 1)) return -1;
    if (axis_gemm(c, false, tmp, out, Q[1], my, nx, my, my, 1, nx, (long long)nx * ny, nzl)) return -1;
    FT* cur = out;
    FT* oth = tmp;
    if (c->nranks == 1) {
        if (three) {
            if (axis_gemm(c, false, cur, oth, Q[2], mz, nx * ny, mz, mz, 1, (long long)nx * ny, 0, 1)) return -1;
            std::swap(cur, oth);
        }
        k_fdm_scale<FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(nx, ny, nz, 0, 0, mx, my, mz, f.lam[0], f.lam[1],
                                                             three ? f.lam[2] : nullptr, f.coef, f.shift, thresh, cur);
        TFB_LAUNCHED();
        if (three) {
            if (axis_gemm(c, true, cur, oth, Q[2], mz, nx * ny, mz, mz, 1, (long long)nx * ny, 0, 1)) return -1;
            std::swap(cur, oth);
        }
    } else {
        // z couples the slabs: transpose to the pencil layout (all z, my y-chunk), transform, scale,
        // transform back, transpose back (two NCCL all-to-alls per solve)
        TFB_CHECK(three, "z-slabs need a 3-D grid");
        if (dist_setup(c)) return -1;
        TfbChunks ch;
        ch.n = c->nranks;
        for (int r = 0; r <= c->nranks; r++) ch.j0[r] = s->j0s[r];
        for (int r = 0; r < c->nranks; r++) ch.dsp[r] = s->a2a_dsp_slab[r];
        const int cyme = s->j0s[c->rank + 1] - s->j0s[c->rank];
        const long long npen = (long long)nz * cyme * nx, lines = (long long)cyme * nx;
        FT *pen0 = (FT*)s->pen[0], *pen1 = (FT*)s->pen[1], *sbuf = (FT*)s->sbuf, *rbuf = (FT*)s->rbuf;
        k_a2a_pack<true, FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(nx, ny, nzl, ch, cur, sbuf);
        TFB_LAUNCHED();
        if (tfb_alltoallv_bytes(c, sbuf, s->a2a_cnt_slab, s->a2a_dsp_slab, pen0, s->a2a_cnt_pen, s->a2a_dsp_pen, (int)sizeof(FT))) return -1;
        if (axis_gemm(c, false, pen0, pen1, Q[2], mz, (int)lines, mz, mz, 1, lines, 0, 1)) return -1;
        k_fdm_scale<FT><<<vec_blocks(npen), 256, 0, c->stream>>>(nx, cyme, nz, s->j0s[c->rank], 0, mx, my, mz, f.lam[0], f.lam[1],
                                                            f.lam[2], f.coef, f.shift, thresh, pen1);
        TFB_LAUNCHED();
        if (axis_gemm(c, true, pen1, pen0, Q[2], mz, (int)lines, mz, mz, 1, lines, 0, 1)) return -1;
        if (tfb_alltoallv_bytes(c, pen0, s->a2a_cnt_pen, s->a2a_dsp_pen, rbuf, s->a2a_cnt_slab, s->a2a_dsp_slab, (int)sizeof(FT))) return -1;
        k_a2a_pack<false, FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(nx, ny, nzl, ch, cur, rbuf);
        TFB_LAUNCHED();
    }
    if (axis_gemm(c, true, cur, oth, Q[1], my, nx, my, my, 1, nx, (long long)nx * ny, nzl)) return -1;
    std::swap(cur, oth);
    // last transform must land in `out`
    FT* dst = (cur == out) ? tmp : out;
    if (axis_gemm(c, true, cur, dst, Q[0], mx, ny * nzl, mx, mx, nx, 1, 0, 1)) return -1;
    if (dst != out) TFB_CUDA(cudaMemcpyAsync(out, dst, sizeof(FT) * ncell, cudaMemcpyDeviceToDevice, c->stream));
    if (mx < nx || my < ny || mz < nz) {
        k_fdm_walls<FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(nx, ny, nzl, k0, mx, my, mz, in, out);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------
// Tensor-core path of the FDM solves (tfb_fdm_tc.cuh): x/y transforms of whole planes as 3xTF32 tcgen05
// contractions, the z direction as precomputed Thomas sweeps.  fp32 storage; used for preconditioning only.
// ------------------------------------------------------------------------------------
using TcGeo = tfbtc::Geo<TFB_TC_KC>;
static int g_tc_grid_per_sm = 0;

static bool tc_ready(const tfb_ctx* c, int v) {
    const tfb_solver_state* s = c->solver;
    const FdmVar& f = s->var[v];
    return c->desc.dim == 3 && c->desc.nz > 1 && f.present && f.tcA[0][0] && f.tcA[1][0] && f.pencil[2] &&
           f.pencil_n[2] == c->desc.nz;
}

static int tc_buffers(tfb_ctx* c) {
    tfb_solver_state* s = c->solver;
    const long long need = c->n_local / c->desc.dof;
    if (need > s->tc32_cap) {
        for (auto& p : s->tc32) { cudaFree(p); p = nullptr; }
        for (auto& p : s->tc32) TFB_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)need * TFB_MAXVAR));
        s->tc32_cap = need;
    }
    return 0;
}

// out[q] = Qy'(bvar[q]) * in[q] * Qx'(bvar[q])^T on every plane of the local slab (forward: transposed eigenvector
// matrices, backward: the matrices)
static int tc_planes(tfb_ctx* c, int nv, const int* bvar, float* const* in, float* const* out, bool backward) {
    tfb_solver_state* s = c->solver;
    const int nx = c->desc.nx, ny = c->desc.ny;
    tfbtc::PlaneArgs a{};
    TFB_CHECK(nv <= tfbtc::MAXQ, "too many arrays");
    for (int q = 0; q < nv; q++) {
        const FdmVar& f = s->var[bvar[q]];
        a.in[q] = in[q]; a.out[q] = out[q];
        a.A1[q] = f.tcA[0][backward ? 1 : 0];
        a.A2[q] = f.tcA[1][backward ? 1 : 0];
    }
    a.narr = nv; a.nplanes = c->nzl; a.rows = ny; a.cols = nx; a.plane_stride = (long long)nx * ny;
    a.k1pad = (nx + TFB_TC_KC - 1) / TFB_TC_KC * TFB_TC_KC;
    a.k2pad = (ny + TFB_TC_KC - 1) / TFB_TC_KC * TFB_TC_KC;
    a.n1 = (ny + 15) / 16 * 16; a.n2 = (nx + 15) / 16 * 16;
    const long long items = (long long)nv * c->nzl;
    const int grid = (int)std::min<long long>(items, 148ll * std::max(g_tc_grid_per_sm, 1));
    tfbtc::tfb_fdm_plane_kernel<TFB_TC_KC, TFB_TC_STAGES><<<grid, tfbtc::THREADS, tfbtc::plane_kernel_smem<TFB_TC_KC, TFB_TC_STAGES>(), c->stream>>>(a);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// Thomas factors of variable v for the current eigenvalues / coefficient / shift.  z-slabs factor their own block and
// set up the interface (spike) data of tfb_fdm_tc.cuh: one all-gather of four coefficients per mode, once per parameter set.
static int tc_thomas_setup(tfb_ctx* c, int v) {
    tfb_solver_state* s = c->solver;
    FdmVar& f = s->var[v];
    if (!f.th_dirty) return 0;
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, G = c->nranks;
    const long long modes = (long long)nx * ny;
    const int k0 = c->desc.k0;
    const int mzl = std::max(0, std::min(c->desc.k1, f.m[2]) - k0);      // unknowns of this variable in the local block
    const long long need = modes * std::max(1, c->nzl);
    if (need > f.th_cap) {
        cudaFree(f.th_inv); cudaFree(f.th_cp);
        f.th_inv = f.th_cp = nullptr;
        TFB_CUDA(cudaMalloc(&f.th_inv, sizeof(float) * (size_t)need));
        TFB_CUDA(cudaMalloc(&f.th_cp, sizeof(float) * (size_t)need));
        if (G > 1) {
            TFB_CHECK(G <= TFB_SPIKE_MAXG, "too many z-slabs for the interface system");
            cudaFree(f.sp_v); cudaFree(f.sp_w); cudaFree(f.sp_coef); cudaFree(f.sp_coef_all); cudaFree(f.sp_weights);
            TFB_CUDA(cudaMalloc(&f.sp_v, sizeof(float) * (size_t)need));
            TFB_CUDA(cudaMalloc(&f.sp_w, sizeof(float) * (size_t)need));
            TFB_CUDA(cudaMalloc(&f.sp_coef, sizeof(float) * 4 * (size_t)modes));
            TFB_CUDA(cudaMalloc(&f.sp_coef_all, sizeof(float) * 4 * (size_t)modes * G));
            TFB_CUDA(cudaMalloc(&f.sp_weights, sizeof(float) * 4 * (size_t)modes * G));
        }
        f.th_cap = need;
    }
    f.mz_local = mzl;
    tfbtc::ThomasVar t;
    t.lx = f.lam[0]; t.ly = f.lam[1]; t.zk = f.pencil[2]; t.inv = f.th_inv; t.cp = f.th_cp;
    t.coef = f.coef; t.shift = f.shift; t.k0 = k0; t.mz = mzl; t.mx = f.m[0]; t.my = f.m[1];
    // an all-Neumann operator (pressure Poisson, pinned scalars) is singular in its constant mode: on one GPU the last
    // pivot of that mode vanishes and is dropped; on slabs the top rank pins it explicitly so that every block is regular
    t.pin_last = (G > 1 && c->rank == G - 1 && (v == c->desc.dim || f.pin_cell >= 0)) ? 1 : 0;
    t.mu_eps = 1e-10 * f.maxden;
    const double thresh = 1e-12 * fabs(f.coef) * f.maxden;
    const unsigned nb = (unsigned)((modes + 127) / 128);
    tfbtc::tfb_thomas_setup_kernel<<<nb, 128, 0, c->stream>>>(t, nx, ny, 0, nz, thresh);
    TFB_LAUNCHED();
    if (G > 1) {
        tfbtc::SpikeVar sp;
        sp.inv = f.th_inv; sp.cp = f.th_cp; sp.zk = f.pencil[2]; sp.v = f.sp_v; sp.w = f.sp_w; sp.coef4 = f.sp_coef;
        sp.coef = f.coef; sp.k0 = k0; sp.mz = mzl; sp.nz_active = f.m[2]; sp.modes = modes;
        const std::vector<double>& zk = f.pencil_h[2];
        const double lo_g = (k0 > 0 && mzl > 0) ? f.coef * zk[k0] : 0.0;
        const double up_g = (mzl > 0 && k0 + mzl < f.m[2]) ? f.coef * zk[2 * (size_t)nz + k0 + mzl - 1] : 0.0;
        tfbtc::tfb_spike_setup_kernel<<<nb, 128, 0, c->stream>>>(sp);
        tfbtc::tfb_spike_scale_kernel<<<nb, 128, 0, c->stream>>>(sp, lo_g, up_g);
        TFB_LAUNCHED(); TFB_LAUNCHED();
        if (tfb_allgather_f32(c, f.sp_coef, f.sp_coef_all, 4 * (size_t)modes)) return -1;
        tfbtc::tfb_spike_weights_kernel<<<(unsigned)((modes + 63) / 64), 64, 0, c->stream>>>(G, c->rank, modes, f.sp_coef_all, f.sp_weights);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    f.th_dirty = false;
    return 0;
}

// in-place tridiagonal solves along z of nv arrays laid out [k][modes] (z-slabs: the local blocks; `iface` then
// receives the first / last unknown of every local solution)
static int tc_thomas(tfb_ctx* c, int nv, const int* vars, float* const* x, long long modes, float* iface) {
    tfb_solver_state* s = c->solver;
    tfbtc::ThomasArgs a{};
    for (int q = 0; q < nv; q++) {
        const FdmVar& f = s->var[vars[q]];
        a.x[q] = x[q]; a.inv[q] = f.th_inv; a.cp[q] = f.th_cp; a.zk[q] = f.pencil[2]; a.coef[q] = f.coef; a.mz[q] = f.mz_local;
    }
    a.narr = nv; a.nz = c->desc.nz; a.k0 = c->desc.k0; a.modes = modes; a.iface = iface;
    static int tb = -1, bs = 128;
    if (tb < 0) {
        const char* e = getenv("TFB_THOMAS_TB");
        tb = e ? atoi(e) : 8;
        if (const char* b2 = getenv("TFB_THOMAS_BS")) bs = atoi(b2);
    }
    dim3 grid((unsigned)((modes + bs - 1) / bs), nv);
    if (tb == 16) tfbtc::tfb_thomas_kernel<16><<<grid, bs, 0, c->stream>>>(a);
    else if (tb == 32) tfbtc::tfb_thomas_kernel<32><<<grid, bs, 0, c->stream>>>(a);
    else if (tb == 4) tfbtc::tfb_thomas_kernel<4><<<grid, bs, 0, c->stream>>>(a);
    else tfbtc::tfb_thomas_kernel<8><<<grid, bs, 0, c->stream>>>(a);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// out[q] = Op_{vars[q]}^-1 in[q] for nv SoA fp32 arrays of the local slab; mid[q] is scratch (in == out is fine).
static int fdm_solve_tc(tfb_ctx* c, int nv, const int* vars, float* const* a, float* const* b, float* const* out) {
    tfb_solver_state* s = c->solver;
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, nzl = c->nzl;
    for (int q = 0; q < nv; q++) {
        TFB_CHECK(tc_ready(c, vars[q]), "tensor-core FDM data missing for a variable");
        if (tc_thomas_setup(c, vars[q])) return -1;
    }
    if (tc_planes(c, nv, vars, a, b, false)) return -1;
    const long long modes = (long long)nx * ny;
    if (c->nranks == 1) {
        if (tc_thomas(c, nv, vars, b, modes, nullptr)) return -1;
    } else {
        // the z lines cross the slabs: local sweeps, one all-gather of the interface values, correction with the spikes
        const int G = c->nranks;
        if (!s->sp_send) {
            TFB_CUDA(cudaMalloc(&s->sp_send, sizeof(float) * TFB_MAXVAR * 2 * (size_t)modes));
            TFB_CUDA(cudaMalloc(&s->sp_recv, sizeof(float) * TFB_MAXVAR * 2 * (size_t)modes * G));
        }
        if (tc_thomas(c, nv, vars, b, modes, s->sp_send)) return -1;
        if (tfb_allgather_f32(c, s->sp_send, s->sp_recv, (size_t)nv * 2 * modes)) return -1;
        tfbtc::SpikeFixArgs fa{};
        for (int q = 0; q < nv; q++) {
            const FdmVar& f = s->var[vars[q]];
            fa.x[q] = b[q]; fa.v[q] = f.sp_v; fa.w[q] = f.sp_w; fa.weights[q] = f.sp_weights; fa.mz[q] = f.mz_local;
        }
        fa.gathered = s->sp_recv; fa.narr = nv; fa.G = G; fa.modes = modes;
        dim3 grid((unsigned)((modes + 127) / 128), nv);
        tfbtc::tfb_spike_fix_kernel<<<grid, 128, 0, c->stream>>>(fa);
        TFB_LAUNCHED();
        TFB_CUDA(cudaGetLastError());
        (void)nz; (void)nzl;
    }
    return tc_planes(c, nv, vars, b, out, true);
}


// two-slot copy of the gradient block G (values only) for the fused head of the scaled-mass preconditioner
static int gell_refresh(tfb_ctx* c, tfb_mat* m) {
    tfb_solver_state* s = c->solver;
    if (s->gell_owner == m && s->gell_version == m->version) return 0;
    const int dof = c->desc.dof, dim = c->desc.dim;
    const long long ncell = c->n_local / dof, plane = (long long)c->desc.nx * c->desc.ny;
    if (!s->gell) {
        TFB_CUDA(cudaMalloc(&s->gell, sizeof(float) * (size_t)dim * 2 * ncell));
        TFB_CUDA(cudaMalloc(&s->dp32, sizeof(float) * (size_t)(ncell + plane)));
        TFB_CUDA(cudaMemset(s->dp32, 0, sizeof(float) * (size_t)(ncell + plane)));
        TFB_CUDA(cudaMalloc(&s->gell_misfit, sizeof(int)));
        const int nax[3] = {c->desc.nx, c->desc.ny, c->desc.nz};
        for (int ax = 0; ax < 3; ax++) {      // 1 / (cell width): row 0 of the metric table of the axis
            std::vector<double> h(nax[ax]);
            TFB_CUDA(cudaMemcpy(h.data(), c->d_met[ax], sizeof(double) * nax[ax], cudaMemcpyDeviceToHost));
            for (auto& x : h) x = 1.0 / x;
            TFB_CUDA(cudaMalloc(&s->d_ih[ax], sizeof(double) * (nax[ax] + 1)));
            TFB_CUDA(cudaMemset(s->d_ih[ax], 0, sizeof(double) * (nax[ax] + 1)));
            TFB_CUDA(cudaMemcpy(s->d_ih[ax], h.data(), sizeof(double) * nax[ax], cudaMemcpyHostToDevice));
        }
    }
    TFB_CUDA(cudaMemsetAsync(s->gell_misfit, 0, sizeof(int), c->stream));
    tfbtc::GEll g{s->gell, s->gell_misfit};
    tfbtc::tfb_gell_build_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, c->stream>>>(
        ncell, dof, dim, c->desc.nx, c->desc.ny, c->row0 / dof, s->subG.row_ptr, s->subG.col, s->subG.vals, g);
    TFB_LAUNCHED();
    int misfit = 0;
    TFB_CUDA(cudaMemcpyAsync(&misfit, s->gell_misfit, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->nranks > 1) {      // every rank must take the same path: the fused head contains a halo exchange
        double mf = misfit;
        TFB_CUDA(cudaMemcpyAsync(s->d_scal + 100, &mf, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (tfb_allreduce_sum(c, s->d_scal + 100, 1)) return -1;
        TFB_CUDA(cudaMemcpyAsync(&mf, s->d_scal + 100, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        misfit = mf != 0.0;
    }
    s->gell_ok = misfit == 0;
    s->gell_owner = m; s->gell_version = m->version;
    return 0;
}

// scaled-mass block preconditioner of a velocity-pressure problem in five launches (head, x/y forward, Thomas,
// x/y backward, tail):  z_p = dp = gamma r_p / |cell|,  z_u = FDM^-1 (r_u - G dp)
static int precond_fused_tc(tfb_ctx* c, int prow, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    const int dof = c->desc.dof, dim = c->desc.dim;
    const long long ncell = c->n_local / dof;
    if (tc_buffers(c)) return -1;
    int vars[3];
    float *a[3], *b[3];
    tfbtc::PreArgs pa{};
    tfbtc::IntArgs ia{};
    for (int v = 0; v < dim; v++) {
        const FdmVar& f = s->var[v];
        vars[v] = v;
        a[v] = s->tc32[0] + (size_t)v * s->tc32_cap;
        b[v] = s->tc32[1] + (size_t)v * s->tc32_cap;
        pa.comp[v] = a[v];
        ia.comp[v] = a[v]; ia.var[v] = v; ia.mx[v] = f.m[0]; ia.my[v] = f.m[1]; ia.mz[v] = f.m[2];
    }
    const long long pin_cell = prow >= 0 ? prow / dof : -1, cell0 = c->row0 / dof;
    pa.dp = s->dp32; pa.gval = s->gell; pa.hx = c->d_met[0]; pa.hy = c->d_met[1]; pa.hz = c->d_met[2];
    pa.gamma = s->gamma; pa.ncell = ncell;
    pa.pin_local = (pin_cell >= cell0 && pin_cell < cell0 + ncell) ? pin_cell - cell0 : -1;
    pa.dof = dof; pa.dim = dim; pa.nx = c->desc.nx; pa.ny = c->desc.ny; pa.k0 = c->desc.k0;
    ia.nv = dim; ia.dof = dof; ia.nx = c->desc.nx; ia.ny = c->desc.ny; ia.k0 = c->desc.k0; ia.ncell = ncell;
    const unsigned nb = vec_blocks(ncell);
    // dof = 4 and an input vector that can take its halo plane in place: row-mapped head and tail, no separate dp pass
    const bool rowmapped = dof == 4 && dim == 3 && (c->nranks == 1 || ghost_capable(c, r));
    if (rowmapped) {
        if (c->nranks > 1 && tfb_halo_up_f64(c, r, const_cast<double*>(r) + c->n_local, (size_t)c->plane_rows)) return -1;
        tfbtc::Pre4Args p4{};
        for (int v = 0; v < 3; v++) p4.comp[v] = a[v];
        p4.dp = s->dp32; p4.gval = s->gell; p4.ihx = s->d_ih[0]; p4.ihy = s->d_ih[1]; p4.ihz = s->d_ih[2];
        p4.gamma = s->gamma; p4.ncell = ncell; p4.pin_local = pa.pin_local;
        p4.nx = c->desc.nx; p4.ny = c->desc.ny; p4.k0 = c->desc.k0;
        const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(((long long)c->desc.nx * c->desc.ny + 255) / 256, 64));
        tfbtc::tfb_tc_pre4_kernel<<<dim3(gx, c->nzl), 256, 0, c->stream>>>(p4, r);
        TFB_LAUNCHED();
    } else {
        tfbtc::tfb_tc_dp_kernel<<<nb, 256, 0, c->stream>>>(pa, r);
        TFB_LAUNCHED();
        // the gradient of the top plane reaches into the slab above: its first plane of dp
        if (tfb_halo_up_f32(c, s->dp32, s->dp32 + ncell, (size_t)c->desc.nx * c->desc.ny)) return -1;
        tfbtc::tfb_tc_pre_kernel<<<nb, 256, 0, c->stream>>>(pa, r);
        TFB_LAUNCHED();
    }
    if (fdm_solve_tc(c, dim, vars, a, b, a)) return -1;
    if (rowmapped) {
        const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>(((long long)c->desc.nx * c->desc.ny + 255) / 256, 64));
        tfbtc::tfb_tc_post4_kernel<<<dim3(gx, c->nzl), 256, 0, c->stream>>>(ia, s->dp32, r, z);
    }
    else tfbtc::tfb_tc_post_kernel<<<nb, 256, 0, c->stream>>>(ia, s->dp32, dim, r, z);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// z(rows of the velocity components) = FDM^-1 (r - sub) through the tensor-core path
static int velocity_fdm_tc(tfb_ctx* c, const double* r, double* z, int skip, const double* sub) {
    tfb_solver_state* s = c->solver;
    const int dof = c->desc.dof, dim = c->desc.dim;
    const long long ncell = c->n_local / dof;
    if (tc_buffers(c)) return -1;
    int vars[TFB_MAXVAR], nv = 0;
    float *a[TFB_MAXVAR], *b[TFB_MAXVAR];
    for (int v = 0; v < dim; v++) {
        if (v == skip) continue;
        a[nv] = s->tc32[0] + (size_t)nv * s->tc32_cap;
        b[nv] = s->tc32[1] + (size_t)nv * s->tc32_cap;
        vars[nv++] = v;
    }
    tfbtc::DeintArgs da{};
    tfbtc::IntArgs ia{};
    for (int q = 0; q < nv; q++) {
        const FdmVar& f = s->var[vars[q]];
        da.comp[q] = a[q]; da.var[q] = vars[q];
        ia.comp[q] = a[q]; ia.var[q] = vars[q]; ia.mx[q] = f.m[0]; ia.my[q] = f.m[1]; ia.mz[q] = f.m[2];
    }
    da.nv = nv; da.dof = dof; da.ncell = ncell;
    ia.nv = nv; ia.dof = dof; ia.nx = c->desc.nx; ia.ny = c->desc.ny; ia.k0 = c->desc.k0; ia.ncell = ncell;
    tfbtc::tfb_deint_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(da, r, sub);
    TFB_LAUNCHED();
    if (fdm_solve_tc(c, nv, vars, a, b, a)) return -1;
    tfbtc::tfb_int_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(ia, r, sub, z);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// pinned Poisson solve on SoA arrays: q = Lp_pinned^-1 rp ; rp is modified at the pin.
// pin_cell is a GLOBAL cell index; with z-slabs the sum and the pinned value are all-reduced.
template <class FT>
static int poisson_solve(tfb_ctx* c, int pvar, long long pin_cell, FT* rp, FT* tmp, FT* q, double pin_sign = 1.0) {
    tfb_solver_state* s = c->solver;
    const long long ncell = c->n_local / c->desc.dof;
    const long long cell0 = c->row0 / c->desc.dof;
    const bool owner = pin_cell >= cell0 && pin_cell < cell0 + ncell;
    const long long pl = pin_cell - cell0;
    if (pin_cell >= 0) {
        TFB_CUDA(cudaMemsetAsync(s->d_scal, 0, sizeof(double) * 3, c->stream));
        k_sum<FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(ncell, rp, s->d_scal);
        TFB_LAUNCHED();
        if (tfb_allreduce_sum(c, s->d_scal, 1)) return -1;
        if (owner) { k_pin_rhs<FT><<<1, 1, 0, c->stream>>>(rp, pl, s->d_scal); TFB_LAUNCHED(); }
    }
    if (fdm_solve<FT>(c, pvar, rp, tmp, q)) return -1;
    if (pin_cell >= 0) {
        if (owner) { k_copy1<FT><<<1, 1, 0, c->stream>>>(q + pl, s->d_scal + 2); TFB_LAUNCHED(); }
        if (tfb_allreduce_sum(c, s->d_scal + 2, 1)) return -1;   // non-owners contribute 0
        k_pin_shift<FT><<<vec_blocks(ncell), 256, 0, c->stream>>>(ncell, q, owner ? pl : -1, s->d_scal, s->d_scal + 2, pin_sign);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    return 0;
}

template <class BT> static int multi_dot(tfb_ctx* c, const BT* V, int nv, const double* w, double* d_out, long long ld = 0, bool reduce = true);
template <class BT> static int multi_axpy(tfb_ctx* c, const BT* V, int nv, const double* d_h, double sign, double* w, double* d_nrm2 = nullptr,
                                          const double* base = nullptr, double bscale = 1.0, double wscale = 1.0, long long ld = 0,
                                          bool reduce = true);
template <class BT> __global__ void k_store_scaled(long long n, const double* __restrict__ scal, int idx, const double* __restrict__ x, BT* __restrict__ y);

// y = x on the rows of the selected variables, 0 elsewhere
// ------------------------------------------------------------------------------------
// Coupled (vertical velocity, scalar) solve -- Rayleigh-Benard (tfb_joint.h)
// ------------------------------------------------------------------------------------
// Horizontal means of the vertical couplings, read off the Jacobian: for every (w, T) entry of a
// w-row and every (T, w) entry of a T-row in the SAME column of cells, value / (hx hy) / (nx ny)
// is added to the table row of its z-offset.  One thread per matrix row.
__global__ void k_joint_couplings(long long nrows, long long row0, int dof, int wv, int sv, int nx, int ny, int nz,
                                  const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ vals,
                                  const double* __restrict__ hx, const double* __restrict__ hy, double* __restrict__ jz) {
    const long long rl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (rl >= nrows) return;
    const long long row = row0 + rl;
    const int rv = (int)(row % dof);
    if (rv != wv && rv != sv) return;
    const long long cell = row / dof, plane = (long long)nx * ny;
    const int i = (int)(cell % nx), j = (int)((cell / nx) % ny), k = (int)(cell / plane);
    if (rv == wv && k >= nz - 1) return;                 // wall row
    const double scale = 1.0 / (hx[i] * hy[j] * (double)plane);
    const int other = rv == wv ? sv : wv;
    double acc0 = 0.0, acc1 = 0.0;                       // same plane / neighbouring plane
    for (int e = row_ptr[rl]; e < row_ptr[rl + 1]; e++) {
        const long long cc = col[e];
        if ((int)(cc % dof) != other) continue;
        const long long ccell = cc / dof;
        if (ccell % plane != cell % plane) continue;
        const int dk = (int)(ccell / plane) - k;
        if (rv == sv && (int)(ccell / plane) >= nz - 1) continue;   // wall face: not an unknown
        if (dk == 0) acc0 += vals[e];
        else if (dk == (rv == wv ? 1 : -1)) acc1 += vals[e];
    }
    if (rv == wv) { atomicAdd(jz + TFB_JZ_B0 * nz + k, acc0 * scale); atomicAdd(jz + TFB_JZ_BP * nz + k, acc1 * scale); }
    else          { atomicAdd(jz + TFB_JZ_C0 * nz + k, acc0 * scale); atomicAdd(jz + TFB_JZ_CM * nz + k, acc1 * scale); }
}

// one thread per horizontal mode (a, b): banded solve along z
template <class VT>
__global__ void __launch_bounds__(128)
k_joint_lines(int ex, int ey, int jofs, int nz, const double* __restrict__ zc, const double* __restrict__ lx,
              const double* __restrict__ ly, double cv, double cT, VT* __restrict__ w, VT* __restrict__ T,
              double* __restrict__ al, double* __restrict__ be) {
    extern __shared__ double s_zc[];
    for (int q = threadIdx.x; q < TFB_JZ_ROWS * nz; q += blockDim.x) s_zc[q] = zc[q];
    __syncthreads();
    const long long modes = (long long)ex * ey;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= modes) return;
    const double mu = lx[m % ex] + ly[jofs + m / ex];
    tfb_joint_line(nz, s_zc, mu, cv, cT, modes, w + m, T + m, al + m, be + m);
}

// factor every horizontal mode once per matrix / substitute per application (tensor-core path, fp32 storage)
__global__ void __launch_bounds__(128)
k_joint_factor(int ex, int ey, int nz, const double* __restrict__ zc, const double* __restrict__ lx, const double* __restrict__ ly,
               double cv, double cT, float* __restrict__ fac) {
    extern __shared__ double s_zc[];
    for (int q = threadIdx.x; q < TFB_JZ_ROWS * nz; q += blockDim.x) s_zc[q] = zc[q];
    __syncthreads();
    const long long modes = (long long)ex * ey, span = (long long)(2 * nz - 1) * modes;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= modes) return;
    const double mu = lx[m % ex] + ly[m / ex];
    tfb_joint_factor_line<float>(nz, s_zc, mu, cv, cT, modes, fac + m, fac + span + m, fac + 2 * span + m, fac + 3 * span + m);
}
// Substitution of tfb_joint_substitute_line (csrc/tfb_joint.h, the tested statement of the recurrences) with the loads of
// KB planes (2 KB rows of the interleaved system) issued ahead of the dependent chain, software-pipelined: one thread per
// mode would otherwise pay a memory latency per row.
template <int KB>
__global__ void __launch_bounds__(128)
k_joint_substitute(int ex, int ey, int nz, const double* __restrict__ zc, double cv, double cT, float* __restrict__ w,
                   float* __restrict__ T, const float* __restrict__ fac) {
    extern __shared__ double s_zc[];
    for (int q = threadIdx.x; q < TFB_JZ_ROWS * nz; q += blockDim.x) s_zc[q] = zc[q];
    __syncthreads();
    const long long st = (long long)ex * ey, span = (long long)(2 * nz - 1) * st;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= st) return;
    const double* KW_LO = s_zc + TFB_JZ_KW_LO * nz;
    const double* KT_LO = s_zc + TFB_JZ_KT_LO * nz;
    float* wp = w + m;
    float* Tp = T + m;
    const float* s1p = fac + m;
    const float* piv = fac + span + m;
    const float* al = fac + 2 * span + m;
    const float* be = fac + 3 * span + m;
    // ---- forward: g_i = (r_i - s2_i g_{i-2} - s1'_i g_{i-1}) / pivot_i, rows (T_k, w_k) of KB planes per block ----
    float rT[KB], rw[KB], a0[KB], p0[KB], a1[KB], p1[KB], nT[KB], nw[KB], na0[KB], np0[KB], na1[KB], np1[KB];
    auto load_f = [&](int k0, float (&xT)[KB], float (&xw)[KB], float (&s0)[KB], float (&q0)[KB], float (&s1)[KB], float (&q1)[KB]) {
#pragma unroll
        for (int t = 0; t < KB; t++) {
            const int k = k0 + t;
            const bool okT = k < nz, okw = k < nz - 1;
            xT[t] = okT ? Tp[(long long)k * st] : 0.f;
            s0[t] = okT ? s1p[(2LL * k) * st] : 0.f;
            q0[t] = okT ? piv[(2LL * k) * st] : 0.f;
            xw[t] = okw ? wp[(long long)k * st] : 0.f;
            s1[t] = okw ? s1p[(2LL * k + 1) * st] : 0.f;
            q1[t] = okw ? piv[(2LL * k + 1) * st] : 0.f;
        }
    };
    float g1 = 0.f, g2 = 0.f;
    load_f(0, rT, rw, a0, p0, a1, p1);
    for (int k0 = 0; k0 < nz; k0 += KB) {
        if (k0 + KB < nz) load_f(k0 + KB, nT, nw, na0, np0, na1, np1);
#pragma unroll
        for (int t = 0; t < KB; t++) {
            const int k = k0 + t;
            if (k < nz) {
                const float s2 = k > 0 ? (float)(cT * KT_LO[k]) : 0.f;
                const float g = (rT[t] - s2 * g2 - a0[t] * g1) * p0[t];
                Tp[(long long)k * st] = g;
                g2 = g1; g1 = g;
            }
            if (k < nz - 1) {
                const float s2 = k > 0 ? (float)(cv * KW_LO[k]) : 0.f;
                const float g = (rw[t] - s2 * g2 - a1[t] * g1) * p1[t];
                wp[(long long)k * st] = g;
                g2 = g1; g1 = g;
            }
        }
#pragma unroll
        for (int t = 0; t < KB; t++) { rT[t] = nT[t]; rw[t] = nw[t]; a0[t] = na0[t]; p0[t] = np0[t]; a1[t] = na1[t]; p1[t] = np1[t]; }
    }
    // ---- backward: y_i = g_i - al_i y_{i+1} - be_i y_{i+2} ----
    auto load_b = [&](int k1, float (&xT)[KB], float (&xw)[KB], float (&s0)[KB], float (&q0)[KB], float (&s1)[KB], float (&q1)[KB]) {
#pragma unroll
        for (int t = 0; t < KB; t++) {
            const int k = k1 - 1 - t;
            const bool okT = k >= 0, okw = k >= 0 && k < nz - 1;
            xT[t] = okT ? Tp[(long long)k * st] : 0.f;
            s0[t] = okT ? al[(2LL * k) * st] : 0.f;
            q0[t] = okT ? be[(2LL * k) * st] : 0.f;
            xw[t] = okw ? wp[(long long)k * st] : 0.f;
            s1[t] = okw ? al[(2LL * k + 1) * st] : 0.f;
            q1[t] = okw ? be[(2LL * k + 1) * st] : 0.f;
        }
    };
    float y1 = 0.f, y2 = 0.f;
    load_b(nz, rT, rw, a0, p0, a1, p1);
    for (int k1 = nz; k1 > 0; k1 -= KB) {
        if (k1 - KB > 0) load_b(k1 - KB, nT, nw, na0, np0, na1, np1);
#pragma unroll
        for (int t = 0; t < KB; t++) {
            const int k = k1 - 1 - t;
            if (k >= 0 && k < nz - 1) {
                const float y = rw[t] - a1[t] * y1 - p1[t] * y2;
                wp[(long long)k * st] = y;
                y2 = y1; y1 = y;
            }
            if (k >= 0) {
                const float y = rT[t] - a0[t] * y1 - p0[t] * y2;
                Tp[(long long)k * st] = y;
                y2 = y1; y1 = y;
            }
        }
#pragma unroll
        for (int t = 0; t < KB; t++) { rT[t] = nT[t]; rw[t] = nw[t]; a0[t] = na0[t]; p0[t] = np0[t]; a1[t] = na1[t]; p1[t] = np1[t]; }
    }
}

__global__ void k_negate_var(long long ncell, int dof, int v, const double* __restrict__ r, double* __restrict__ z) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x)
        z[i * dof + v] = -r[i * dof + v];
}

static int joint_refresh(tfb_ctx* c, tfb_mat* m) {
    tfb_solver_state* s = c->solver;
    if (s->jz_owner == m && s->jz_version == m->version) return 0;
    const int nz = c->desc.nz;
    TFB_CUDA(cudaMemsetAsync(s->d_jz + TFB_JZ_B0 * nz, 0, sizeof(double) * 4 * nz, c->stream));
    k_joint_couplings<<<(unsigned)((c->n_local + 255) / 256), 256, 0, c->stream>>>(
        c->n_local, c->row0, c->desc.dof, s->joint_w, s->joint_s, c->desc.nx, c->desc.ny, nz, c->d_row_ptr, c->d_col,
        m->d_vals, c->d_met[0], c->d_met[1], s->d_jz);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    if (tfb_allreduce_sum(c, s->d_jz + TFB_JZ_B0 * nz, 4 * nz)) return -1;
    s->jz_owner = m; s->jz_version = m->version;
    return 0;
}

// (z_w, z_T) = Fwt^-1 (r_w, r_T) on interleaved vectors: the x/y transforms of the vertical velocity's
// fast-diagonalisation basis, one banded solve per horizontal mode, transforms back.
static int joint_solve_tc(tfb_ctx* c, const double* r, double* z);
static int joint_solve(tfb_ctx* c, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    if (s->precond_tc && c->nranks == 1) return joint_solve_tc(c, r, z);
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, nzl = c->nzl, dof = c->desc.dof;
    const int wv = s->joint_w, sv = s->joint_s;
    // horizontal basis: the vertical velocity's (exact for the viscous block; the scalar block then sees w's
    // side-wall folds, a boundary-layer-sized defect the Krylov iteration absorbs).  With Pr > 1 the viscous
    // block dominates, and the scalar's own basis was measurably worse at low Rayleigh numbers.
    const FdmVar& f = s->var[wv];
    TFB_CHECK(f.present && s->var[sv].present && f.m[0] == nx && f.m[1] == ny, "FDM basis of the vertical velocity missing");
    const long long plane = (long long)nx * ny, ncell = plane * nzl;
    const bool dist = c->nranks > 1;
    if (dist && dist_setup(c)) return -1;
    const int cyme = dist ? s->j0s[c->rank + 1] - s->j0s[c->rank] : ny;
    const long long lines = (long long)cyme * nx, npen = lines * nz;
    if (!s->jbuf[0]) {
        TFB_CUDA(cudaMalloc(&s->jbuf[0], sizeof(double) * 2 * ncell));
        TFB_CUDA(cudaMalloc(&s->jbuf[1], sizeof(double) * 2 * ncell));
        TFB_CUDA(cudaMalloc(&s->jab, sizeof(double) * 4 * std::max(ncell, npen)));
    }
    double *a = s->jbuf[0], *b = s->jbuf[1];
    const unsigned vb = vec_blocks(ncell);
    k_deinterleave<double><<<vb, 256, 0, c->stream>>>(ncell, dof, wv, r, a);
    k_deinterleave<double><<<vb, 256, 0, c->stream>>>(ncell, dof, sv, r, a + ncell);
    TFB_LAUNCHED(); TFB_LAUNCHED();
    if (axis_gemm<double>(c, false, a, b, f.Q[0], nx, ny * nzl * 2, nx, nx, nx, 1, 0, 1)) return -1;
    if (axis_gemm<double>(c, false, b, a, f.Q[1], ny, nx, ny, ny, 1, nx, plane, nzl * 2)) return -1;
    const size_t jsmem = sizeof(double) * TFB_JZ_ROWS * nz;
    if (!dist) {
        k_joint_lines<double><<<(unsigned)((plane + 127) / 128), 128, jsmem, c->stream>>>(
            nx, ny, 0, nz, s->d_jz, f.lam[0], f.lam[1], f.coef, s->var[sv].coef, a, a + ncell, s->jab, s->jab + 2 * ncell);
        TFB_LAUNCHED();
    } else {
        // the lines run through every slab: transpose w and T to the pencil layout (all z, my y-chunk) with
        // one all-to-all each, solve, transpose back -- the same exchange the FDM z-transform uses
        TfbChunks ch;
        ch.n = c->nranks;
        for (int q = 0; q <= c->nranks; q++) ch.j0[q] = s->j0s[q];
        for (int q = 0; q < c->nranks; q++) ch.dsp[q] = s->a2a_dsp_slab[q];
        for (int h = 0; h < 2; h++) {
            k_a2a_pack<true, double><<<vb, 256, 0, c->stream>>>(nx, ny, nzl, ch, a + h * ncell, s->sbuf);
            TFB_LAUNCHED();
            if (tfb_alltoallv_bytes(c, s->sbuf, s->a2a_cnt_slab, s->a2a_dsp_slab, s->pen[h], s->a2a_cnt_pen, s->a2a_dsp_pen, (int)sizeof(double))) return -1;
        }
        k_joint_lines<double><<<(unsigned)((lines + 127) / 128), 128, jsmem, c->stream>>>(
            nx, cyme, s->j0s[c->rank], nz, s->d_jz, f.lam[0], f.lam[1], f.coef, s->var[sv].coef, s->pen[0], s->pen[1],
            s->jab, s->jab + 2 * npen);
        TFB_LAUNCHED();
        for (int h = 0; h < 2; h++) {
            if (tfb_alltoallv_bytes(c, s->pen[h], s->a2a_cnt_pen, s->a2a_dsp_pen, s->rbuf, s->a2a_cnt_slab, s->a2a_dsp_slab, (int)sizeof(double))) return -1;
            k_a2a_pack<false, double><<<vb, 256, 0, c->stream>>>(nx, ny, nzl, ch, a + h * ncell, s->rbuf);
            TFB_LAUNCHED();
        }
    }
    if (axis_gemm<double>(c, true, a, b, f.Q[1], ny, nx, ny, ny, 1, nx, plane, nzl * 2)) return -1;
    if (axis_gemm<double>(c, true, b, a, f.Q[0], nx, ny * nzl * 2, nx, nx, nx, 1, 0, 1)) return -1;
    k_interleave<double><<<vb, 256, 0, c->stream>>>(ncell, dof, wv, a, z, 1.0);
    k_interleave<double><<<vb, 256, 0, c->stream>>>(ncell, dof, sv, a + ncell, z, 1.0);
    TFB_LAUNCHED(); TFB_LAUNCHED();
    if (c->desc.k0 + nzl == nz) {   // the top-wall rows of w carry a -1 diagonal (the slab that owns the last plane)
        const long long top = (ncell - plane) * dof;
        k_negate_var<<<vec_blocks(plane), 256, 0, c->stream>>>(plane, dof, wv, r + top, z + top);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// Coupled (w, T) solve on the tensor-core path: x/y transforms of both arrays in the vertical velocity's basis
// (3xTF32 plane kernel), the pentadiagonal line solve per horizontal mode on fp32 arrays (fp64 arithmetic), back.
static int joint_solve_tc(tfb_ctx* c, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, nzl = c->nzl, dof = c->desc.dof;
    const int wv = s->joint_w, sv = s->joint_s;
    const FdmVar& f = s->var[wv];
    TFB_CHECK(tc_ready(c, wv) && s->var[sv].present && f.m[0] == nx && f.m[1] == ny, "tensor-core basis of the vertical velocity missing");
    const long long plane = (long long)nx * ny, ncell = plane * nzl;
    if (tc_buffers(c)) return -1;
    float* a[2] = {s->tc32[0], s->tc32[0] + s->tc32_cap};
    float* b[2] = {s->tc32[1], s->tc32[1] + s->tc32_cap};
    const int basis[2] = {wv, wv};
    tfbtc::DeintArgs da{};
    da.comp[0] = a[0]; da.comp[1] = a[1]; da.var[0] = wv; da.var[1] = sv; da.nv = 2; da.dof = dof; da.ncell = ncell;
    tfbtc::tfb_deint_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(da, r, nullptr);
    TFB_LAUNCHED();
    if (tc_planes(c, 2, basis, a, b, false)) return -1;
    const size_t jsmem = sizeof(double) * TFB_JZ_ROWS * nz;
    if (!s->jfac) TFB_CUDA(cudaMalloc(&s->jfac, sizeof(float) * 4 * (size_t)(2 * nz - 1) * plane));
    if (s->jfac_owner != s->jz_owner || s->jfac_version != s->jz_version) {
        // the couplings were re-read for this matrix (joint_refresh): factor all modes once, every application substitutes
        k_joint_factor<<<(unsigned)((plane + 127) / 128), 128, jsmem, c->stream>>>(nx, ny, nz, s->d_jz, f.lam[0], f.lam[1], f.coef,
                                                                                   s->var[sv].coef, s->jfac);
        TFB_LAUNCHED();
        s->jfac_owner = s->jz_owner; s->jfac_version = s->jz_version;
    }
    k_joint_substitute<4><<<(unsigned)((plane + 127) / 128), 128, jsmem, c->stream>>>(nx, ny, nz, s->d_jz, f.coef, s->var[sv].coef,
                                                                                   b[0], b[1], s->jfac);
    TFB_LAUNCHED();
    if (tc_planes(c, 2, basis, b, a, true)) return -1;
    tfbtc::IntArgs ia{};
    ia.comp[0] = a[0]; ia.comp[1] = a[1]; ia.var[0] = wv; ia.var[1] = sv;
    ia.mx[0] = ia.mx[1] = nx; ia.my[0] = ia.my[1] = ny; ia.mz[0] = nz - 1; ia.mz[1] = nz;   // top-wall rows of w: -r
    ia.nv = 2; ia.dof = dof; ia.nx = nx; ia.ny = ny; ia.k0 = c->desc.k0; ia.ncell = ncell;
    tfbtc::tfb_int_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(ia, r, nullptr, z);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_mask_copy(long long n, int dof, unsigned mask, const double* __restrict__ x, double* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = ((mask >> (int)(i % dof)) & 1u) ? x[i] : 0.0;
}

// z(velocity rows) += FDM_v^-1 r_v for every velocity component (the diffusion part of the block)
template <class FT>
static int velocity_fdm(tfb_ctx* c, const double* r, double* z, int skip = -1) {
    tfb_solver_state* s = c->solver;
    if (s->precond_tc) return velocity_fdm_tc(c, r, z, skip, nullptr);
    const int dof = c->desc.dof, dim = c->desc.dim;
    const long long ncell = c->n_local / dof;
    FT *c0 = (FT*)s->comp[0], *c1 = (FT*)s->comp[1], *c2 = (FT*)s->comp[2];
    const unsigned vb = vec_blocks(ncell);
    for (int v = 0; v < dim; v++) {
        if (v == skip) continue;
        k_deinterleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, v, r, c0);
        if (fdm_solve<FT>(c, v, c0, c1, c2)) return -1;
        k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, v, c2, z, 1.0);
        TFB_LAUNCHED(); TFB_LAUNCHED();
    }
    return 0;
}

// Approximate inverse of the block the velocity sub-solve works on.  Default: the velocity block,
// one FDM (diffusion) solve per component.  Coupled mode (Rayleigh-Benard): the (velocity, scalar)
// block -- horizontal components by FDM, (w, T) by the coupled line solve.
// Tensor-core version of the coupled block: u, v (diffusion solves) and (w, T) (coupled line solve) share one pass of
// the plane kernel in each direction -- four arrays per launch instead of two launches of two.
static int block_fdm_tc_joint(tfb_ctx* c, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    const int nx = c->desc.nx, ny = c->desc.ny, nz = c->desc.nz, dof = c->desc.dof;
    const int wv = s->joint_w, sv = s->joint_s;
    const FdmVar& fw = s->var[wv];
    TFB_CHECK(tc_ready(c, 0) && tc_ready(c, 1) && tc_ready(c, wv) && s->var[sv].present && fw.m[0] == nx && fw.m[1] == ny,
              "tensor-core data of the coupled block missing");
    const long long plane = (long long)nx * ny, ncell = plane * c->nzl;
    if (tc_buffers(c)) return -1;
    if (tc_thomas_setup(c, 0) || tc_thomas_setup(c, 1)) return -1;
    const int vars[4] = {0, 1, wv, sv}, basis[4] = {0, 1, wv, wv};
    float *a[4], *b[4];
    tfbtc::DeintArgs da{};
    tfbtc::IntArgs ia{};
    for (int q = 0; q < 4; q++) {
        a[q] = s->tc32[0] + (size_t)q * s->tc32_cap;
        b[q] = s->tc32[1] + (size_t)q * s->tc32_cap;
        da.comp[q] = a[q]; da.var[q] = vars[q];
        ia.comp[q] = a[q]; ia.var[q] = vars[q];
        const FdmVar& f = s->var[vars[q]];
        ia.mx[q] = q < 2 ? f.m[0] : nx; ia.my[q] = q < 2 ? f.m[1] : ny;
        ia.mz[q] = q < 2 ? f.m[2] : (q == 2 ? nz - 1 : nz);      // top-wall rows of w: -r
    }
    da.nv = 4; da.dof = dof; da.ncell = ncell;
    ia.nv = 4; ia.dof = dof; ia.nx = nx; ia.ny = ny; ia.k0 = c->desc.k0; ia.ncell = ncell;
    tfbtc::tfb_deint_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(da, r, nullptr);
    TFB_LAUNCHED();
    if (tc_planes(c, 4, basis, a, b, false)) return -1;
    if (tc_thomas(c, 2, vars, b, plane, nullptr)) return -1;
    const size_t jsmem = sizeof(double) * TFB_JZ_ROWS * nz;
    if (!s->jfac) TFB_CUDA(cudaMalloc(&s->jfac, sizeof(float) * 4 * (size_t)(2 * nz - 1) * plane));
    if (s->jfac_owner != s->jz_owner || s->jfac_version != s->jz_version) {
        k_joint_factor<<<(unsigned)((plane + 127) / 128), 128, jsmem, c->stream>>>(nx, ny, nz, s->d_jz, fw.lam[0], fw.lam[1], fw.coef,
                                                                                   s->var[sv].coef, s->jfac);
        TFB_LAUNCHED();
        s->jfac_owner = s->jz_owner; s->jfac_version = s->jz_version;
    }
    k_joint_substitute<4><<<(unsigned)((plane + 127) / 128), 128, jsmem, c->stream>>>(nx, ny, nz, s->d_jz, fw.coef, s->var[sv].coef,
                                                                                   b[2], b[3], s->jfac);
    TFB_LAUNCHED();
    if (tc_planes(c, 4, basis, b, a, true)) return -1;
    tfbtc::tfb_int_kernel<<<vec_blocks(ncell), 256, 0, c->stream>>>(ia, r, nullptr, z);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

template <class FT>
static int block_fdm(tfb_ctx* c, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    if (!s->joint_on) return velocity_fdm<FT>(c, r, z);
    if (s->precond_tc && c->nranks == 1 && c->desc.dim == 3) return block_fdm_tc_joint(c, r, z);
    if (velocity_fdm<FT>(c, r, z, s->joint_w)) return -1;
    return joint_solve(c, r, z);
}

// Velocity sub-solve  z_u = F^-1 r_u  of the block preconditioner.  F = the velocity-velocity block
// of J (diffusion + linearised convection [+ Coriolis]).  inner_its == 0: one FDM solve (exact for
// the diffusion part only).  inner_its > 0: that many steps (at most) of right-preconditioned GMRES
// on F with the FDM solve as its preconditioner -- the outer FGMRES is flexible, so the inner
// iteration may stop on a loose tolerance.  z's velocity rows must be zero on entry.
template <class FT>
static int velocity_solve(tfb_ctx* c, tfb_mat* m, int prow, const double* ru, double* z) {
    tfb_solver_state* s = c->solver;
    const int dof = c->desc.dof, dim = c->desc.dim;
    const long long n = c->n_local;
    const unsigned velmask = ((1u << dim) - 1u) | (s->joint_on ? 1u << s->joint_s : 0u);   // the block's variables
    if (s->inner_its <= 0) return block_fdm<FT>(c, ru, z);
    const int k = s->inner_its;
    if (k > s->inner_cap) {
        cudaFree(s->d_Vi); cudaFree(s->d_Zi);
        s->d_Vi = s->d_Zi = nullptr;
        TFB_CUDA(cudaMalloc(&s->d_Vi, sizeof(double) * (size_t)n * (k + 1)));
        TFB_CUDA(cudaMalloc(&s->d_Zi, sizeof(double) * (size_t)n * k));
        s->inner_cap = k;
    }
    double* V = s->d_Vi;
    double* Z = s->d_Zi;
    double* w = s->vec[0];              // free at this point of apply_precond
    double* d_hh = s->d_scal + 8;       // needs k + 3 <= 8 doubles beyond: use a dedicated slice of d_scal
    const unsigned nb = vec_blocks(n);
    k_mask_copy<<<nb, 256, 0, c->stream>>>(n, dof, velmask, ru, w);
    TFB_LAUNCHED();
    double beta = 0.0;
    if (multi_dot<double>(c, w, 1, w, d_hh)) return -1;
    TFB_CUDA(cudaMemcpyAsync(&beta, d_hh, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    beta = sqrt(beta);
    if (beta == 0.0) return 0;
    TFB_CUDA(cudaMemcpyAsync(s->d_scal + 5, &beta, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_store_scaled<double><<<nb, 256, 0, c->stream>>>(n, s->d_scal, 5, w, V);
    TFB_LAUNCHED();
    std::vector<double> H((size_t)(k + 1) * k, 0.0), g(k + 1, 0.0), cs(k), sn(k), hcol(k + 2), y(k);
    auto Hx = [&](int i, int j) -> double& { return H[(size_t)j * (k + 1) + i]; };
    g[0] = beta;
    int j = 0;
    for (; j < k; j++) {
        double* zj = Z + (size_t)j * n;
        TFB_CUDA(cudaMemsetAsync(zj, 0, sizeof(double) * n, c->stream));
        if (block_fdm<FT>(c, V + (size_t)j * n, zj)) return -1;
        if (spmv(c, m, zj, w, prow, velmask, velmask)) return -1;
        if (multi_dot<double>(c, V, j + 1, w, d_hh)) return -1;
        if (multi_axpy<double>(c, V, j + 1, d_hh, -1.0, w, d_hh + k + 1)) return -1;
        TFB_CUDA(cudaMemcpyAsync(hcol.data(), d_hh, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaMemcpyAsync(hcol.data() + k + 1, d_hh + k + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        const double hn = sqrt(hcol[k + 1]);
        for (int i = 0; i <= j; i++) Hx(i, j) = hcol[i];
        Hx(j + 1, j) = hn;
        for (int i = 0; i < j; i++) {
            const double t = cs[i] * Hx(i, j) + sn[i] * Hx(i + 1, j);
            Hx(i + 1, j) = -sn[i] * Hx(i, j) + cs[i] * Hx(i + 1, j);
            Hx(i, j) = t;
        }
        const double d = hypot(Hx(j, j), Hx(j + 1, j));
        cs[j] = Hx(j, j) / d; sn[j] = Hx(j + 1, j) / d;
        Hx(j, j) = d; Hx(j + 1, j) = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
        s->inner_total++;
        if (fabs(g[j + 1]) <= s->inner_tol * beta || hn == 0.0 || j + 1 == k) { j++; break; }
        TFB_CUDA(cudaMemcpyAsync(s->d_scal + 5, &hn, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_store_scaled<double><<<nb, 256, 0, c->stream>>>(n, s->d_scal, 5, w, V + (size_t)(j + 1) * n);
        TFB_LAUNCHED();
    }
    const int kk = j;
    for (int i = kk - 1; i >= 0; i--) {
        double acc = g[i];
        for (int l = i + 1; l < kk; l++) acc -= Hx(i, l) * y[l];
        y[i] = acc / Hx(i, i);
    }
    TFB_CUDA(cudaMemcpyAsync(d_hh, y.data(), sizeof(double) * kk, cudaMemcpyHostToDevice, c->stream));
    if (multi_axpy<double>(c, Z, kk, d_hh, 1.0, z, nullptr)) return -1;
    TFB_CUDA(cudaStreamSynchronize(c->stream));   // y (host) must outlive the copy
    return 0;
}

// z = P^-1 r  (interleaved vectors of length n_local)
// Scaled-mass Schur complement: z_p = ta_p = gamma * r_p / (hx hy hz); the pinned pressure row is -1 * p0 = r
__global__ void k_schur_mass(long long ncell, int dof, int pv, int nx, int ny, int k0, const double* __restrict__ hx,
                             const double* __restrict__ hy, const double* __restrict__ hz, double gamma, long long pin_local,
                             const double* __restrict__ r, double* __restrict__ z, double* __restrict__ ta) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(cell % nx), j = (int)((cell / nx) % ny), k = k0 + (int)(cell / ((long long)nx * ny));
        const double rp = r[cell * dof + pv];
        const double v = cell == pin_local ? -rp : gamma * rp / ((hx[i] * hy[j]) * hz[k]);
        z[cell * dof + pv] = v;
        ta[cell * dof + pv] = v;
    }
}

template <class FT> static int velocity_fdm(tfb_ctx* c, const double* r, double* z, int skip);
__global__ void k_lincomb(long long n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                          double cc, const double* __restrict__ z, double* __restrict__ out);
// gamma = 2.5 * rho(A Ah^-1) * |c_visc|, rho from a few power iterations on the velocity block (A: the matrix's velocity
// block incl. convection, Ah: its diffusion part as solved by the FDM).  With diffusion-only velocity solves the
// preconditioned velocity block has its spectrum in [1, rho]; the pressure block is placed inside that range.  Numpy
// prototypes on the oracle's matrices (16^3, Re 100 / 400, uniform and stretched): 115 / 305 / 115 iterations against
// 121 / 436 / 111 with the least-squares commutator, at 40 % of its cost per application.
static int schur_gamma_refresh(tfb_ctx* c, tfb_mat* m, int prow) {
    tfb_solver_state* s = c->solver;
    if (s->gam_owner == m && s->gam_version == m->version) return 0;
    const int dim = c->desc.dim;
    const long long n = c->n_local;
    const unsigned velmask = (1u << dim) - 1u;
    double *v = s->vec[0], *y = s->vec[1], *w = s->vec[2];
    k_mask_copy<<<vec_blocks(n), 256, 0, c->stream>>>(n, c->desc.dof, velmask, s->d_mass, v);
    TFB_LAUNCHED();
    const int NIT = 8;
    double nrm[NIT + 1];
    for (int it = 0; it <= NIT; it++) {
        double n2 = 0.0;
        if (multi_dot<double>(c, v, 1, v, s->d_h)) return -1;
        TFB_CUDA(cudaMemcpyAsync(&n2, s->d_h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        nrm[it] = sqrt(n2);
        if (it == NIT || !(nrm[it] > 0.0)) break;
        TFB_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * n, c->stream));
        if (velocity_fdm<double>(c, v, y, -1)) return -1;
        if (spmv(c, m, y, w, prow, velmask, velmask)) return -1;
        k_lincomb<<<vec_blocks(n), 256, 0, c->stream>>>(n, 1.0 / nrm[it], w, 0.0, nullptr, 0.0, nullptr, v);   // growth = |v| next round
        TFB_LAUNCHED();
    }
    // |v_{k+1}| is the growth factor of step k (v was normalised before each product); the dominant eigenvalues can
    // be a complex pair, so the estimate is the geometric mean of the last four factors
    double rho = 1.0;
    if (nrm[NIT] > 0.0) rho = pow(nrm[NIT] * nrm[NIT - 1] * nrm[NIT - 2] * nrm[NIT - 3], 0.25);
    rho = std::max(1.0, rho);
    s->gamma_rho = rho;
    s->gamma = 2.5 * rho * fabs(s->var[0].coef);
    s->gam_owner = m; s->gam_version = m->version;
    return 0;
}

template <class FT>
static int apply_precond_t(tfb_ctx* c, tfb_mat* m, int prow, const double* r, double* z) {
    tfb_solver_state* s = c->solver;
    const int dof = c->desc.dof, dim = c->desc.dim, pv = dim;
    const long long n = c->n_local, ncell = n / dof;
    const unsigned velmask = (1u << dim) - 1u, pmask = 1u << pv, smask = ((1u << dof) - 1u) & ~(velmask | pmask);
    const long long pin_cell = prow >= 0 ? prow / dof : -1;   // global cell of the pinned pressure
    FT *c0 = (FT*)s->comp[0], *c1 = (FT*)s->comp[1], *c2 = (FT*)s->comp[2];
    double *ta = s->vec[0], *tb = s->vec[1], *tc = s->vec[2], *ru = s->vec[3];
    const unsigned vb = vec_blocks(ncell);
    if (s->schur_mass && s->precond_tc && s->gell_ok && s->inner_its <= 0 && !s->joint_on && !smask)
        return precond_fused_tc(c, prow, r, z);
    TFB_CUDA(cudaMemsetAsync(z, 0, sizeof(double) * n, c->stream));
    // ---- scalars: s = At^-1 r_s ; ru = r - B s (velocity rows) ----
    TFB_CUDA(cudaMemcpyAsync(ru, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    const bool joint = s->joint_on;
    if (smask && !joint) {
        for (int v = pv + 1; v < dof; v++) {
            k_deinterleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, v, r, c0);
            if (s->var[v].pin_cell >= 0) {
                if (poisson_solve<FT>(c, v, s->var[v].pin_cell, c0, c1, c2, s->var[v].pin_sign)) return -1;
            } else if (fdm_solve<FT>(c, v, c0, c1, c2)) return -1;
            k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, v, c2, z, 1.0);
            TFB_LAUNCHED(); TFB_LAUNCHED();
        }
        if (sub_spmv(c, s->subB, z, ta)) return -1;                      // B s
        k_axpy<<<vec_blocks(n), 256, 0, c->stream>>>(n, -1.0, ta, ru);
        TFB_LAUNCHED();
    }
    if (s->schur_mass) {
        // ---- pressure: dp = gamma r_p / V (scaled mass matrix), then u = Ah^-1 (ru - G dp) ----
        const long long cell0 = (long long)c->desc.nx * c->desc.ny * c->desc.k0;
        const long long pin_local = (pin_cell >= cell0 && pin_cell < cell0 + ncell) ? pin_cell - cell0 : -1;
        TFB_CUDA(cudaMemsetAsync(ta, 0, sizeof(double) * n, c->stream));
        k_schur_mass<<<vb, 256, 0, c->stream>>>(ncell, dof, pv, c->desc.nx, c->desc.ny, c->desc.k0, c->d_met[0], c->d_met[1],
                                               c->d_met[2], s->gamma, pin_local, r, z, ta);
        TFB_LAUNCHED();
        if (sub_spmv(c, s->subG, ta, tb)) return -1;                     // G dp
        if (s->precond_tc && s->inner_its <= 0 && !joint && !smask) {
            // tensor-core path: r - G dp is formed while the components are split off
            if (velocity_fdm_tc(c, r, z, -1, tb)) return -1;
            TFB_CUDA(cudaGetLastError());
            return 0;
        }
        k_axpy<<<vec_blocks(n), 256, 0, c->stream>>>(n, -1.0, tb, ru);
        TFB_LAUNCHED();
        if (velocity_solve<FT>(c, m, prow, ru, z)) return -1;
        TFB_CUDA(cudaGetLastError());
        return 0;
    }
    // ---- pressure: dp = -Lp^-1 D M^-1 A M^-1 G Lp^-1 r_p ----
    k_deinterleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, pv, r, c0);
    if (poisson_solve<FT>(c, pv, pin_cell, c0, c1, c2)) return -1;
    TFB_CUDA(cudaMemsetAsync(ta, 0, sizeof(double) * n, c->stream));
    k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, pv, c2, ta, 1.0);
    TFB_LAUNCHED(); TFB_LAUNCHED();
    if (sub_spmv(c, s->subG, ta, tb, s->d_mass)) return -1;              // M^-1 G t
    if (!joint) {
        if (spmv(c, m, tb, tc, prow, velmask, velmask, s->d_mass)) return -1; // M^-1 A (.)
    } else {
        // scalar eliminated: the commutator sees  A - B At^-1 C  (At^-1: the scalar's diffusion solve)
        const int sv = s->joint_s;
        if (spmv(c, m, tb, tc, prow, velmask, velmask)) return -1;
        if (sub_spmv(c, s->subC, tb, ta)) return -1;
        k_deinterleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, sv, ta, c0);
        if (fdm_solve<FT>(c, sv, c0, c1, c2)) return -1;
        TFB_CUDA(cudaMemsetAsync(ta, 0, sizeof(double) * n, c->stream));
        k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, sv, c2, ta, 1.0);
        TFB_LAUNCHED(); TFB_LAUNCHED();
        if (sub_spmv(c, s->subB, ta, tb)) return -1;
        k_axpy<<<vec_blocks(n), 256, 0, c->stream>>>(n, -1.0, tb, tc);
        k_rowdiv<<<vec_blocks(n), 256, 0, c->stream>>>(n, s->d_mass, tc);
        TFB_LAUNCHED(); TFB_LAUNCHED();
    }
    if (sub_spmv(c, s->subD, tc, ta)) return -1;                         // D (.)
    k_deinterleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, pv, ta, c0);
    if (poisson_solve<FT>(c, pv, pin_cell, c0, c1, c2)) return -1;
    k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, pv, c2, z, -1.0);  // dp into z
    TFB_LAUNCHED(); TFB_LAUNCHED();
    // ---- velocities: u = Ah^-1 (ru - G dp) ----
    TFB_CUDA(cudaMemsetAsync(ta, 0, sizeof(double) * n, c->stream));
    k_interleave<FT><<<vb, 256, 0, c->stream>>>(ncell, dof, pv, c2, ta, -1.0);
    TFB_LAUNCHED();
    if (sub_spmv(c, s->subG, ta, tb)) return -1;                         // G dp
    k_axpy<<<vec_blocks(n), 256, 0, c->stream>>>(n, -1.0, tb, ru);
    TFB_LAUNCHED();
    if (velocity_solve<FT>(c, m, prow, ru, z)) return -1;
    TFB_CUDA(cudaGetLastError());
    return 0;
}

static int apply_precond(tfb_ctx* c, tfb_mat* m, int prow, const double* r, double* z) {
    return c->solver->precond_single ? apply_precond_t<float>(c, m, prow, r, z) : apply_precond_t<double>(c, m, prow, r, z);
}

// ------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------
extern "C" int tfb_fdm_set(tfb_ctx* c, int var, int axis, int m, const double* Q, const double* lam, double coef) {
    TFB_CHECK(c && var >= 0 && var < TFB_MAXVAR && axis >= 0 && axis < 3 && m > 0 && Q && lam, "bad arguments");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    tfb_solver_state* s = solver_of(c);
    FdmVar& f = s->var[var];
    if (f.m[axis] != m) {
        cudaFree(f.Q[axis]); cudaFree(f.lam[axis]); cudaFree(f.Qf[axis]);
        TFB_CUDA(cudaMalloc(&f.Q[axis], sizeof(double) * m * m));
        TFB_CUDA(cudaMalloc(&f.lam[axis], sizeof(double) * m));
        TFB_CUDA(cudaMalloc(&f.Qf[axis], sizeof(float) * m * m));
        f.m[axis] = m;
    }
    TFB_CUDA(cudaMemcpy(f.Q[axis], Q, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    TFB_CUDA(cudaMemcpy(f.lam[axis], lam, sizeof(double) * m, cudaMemcpyHostToDevice));
    {
        std::vector<float> qf((size_t)m * m);
        for (size_t i = 0; i < qf.size(); i++) qf[i] = (float)Q[i];
        TFB_CUDA(cudaMemcpy(f.Qf[axis], qf.data(), sizeof(float) * qf.size(), cudaMemcpyHostToDevice));
    }
    f.coef = coef;
    f.present = true;
    f.th_dirty = true;
    double mx = 0.0;
    for (int i = 0; i < m; i++) mx = std::max(mx, fabs(lam[i]));
    f.maxden = std::max(f.maxden, 3.0 * mx);
    // tensor-core copies of the x / y bases: Q^T (forward) and Q (backward), tf32 hi/lo split, swizzled K-chunks
    const int nax = axis == 0 ? c->desc.nx : c->desc.ny;
    if (axis < 2 && c->desc.dim == 3 && c->desc.nz > 1 && nax <= tfbtc::MROWS && m <= nax) {
        const int kpad = (nax + TFB_TC_KC - 1) / TFB_TC_KC * TFB_TC_KC;
        std::vector<double> qt((size_t)m * m);
        for (int i = 0; i < m; i++)
            for (int a = 0; a < m; a++) qt[(size_t)a * m + i] = Q[(size_t)i * m + a];
        std::vector<float> fmt;
        for (int dir = 0; dir < 2; dir++) {
            tfbtc::tfb_tc_format_matrix<TFB_TC_KC>(dir == 0 ? qt.data() : Q, m, m, m, kpad, fmt);
            cudaFree(f.tcA[axis][dir]);
            f.tcA[axis][dir] = nullptr;
            TFB_CUDA(cudaMalloc(&f.tcA[axis][dir], sizeof(float) * fmt.size()));
            TFB_CUDA(cudaMemcpy(f.tcA[axis][dir], fmt.data(), sizeof(float) * fmt.size(), cudaMemcpyHostToDevice));
        }
        auto kern = tfbtc::tfb_fdm_plane_kernel<TFB_TC_KC, TFB_TC_STAGES>;
        const int smem = (int)tfbtc::plane_kernel_smem<TFB_TC_KC, TFB_TC_STAGES>();
        TFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        TFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        TFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, tfbtc::THREADS, smem));
        g_tc_grid_per_sm = std::max(1, std::min(occ, 2));   // TMEM: 256 of 512 columns per CTA
    }
    return 0;
}

// The 1-D stencils behind the eigen-decompositions of tfb_fdm_set: K = tridiag(lower, diag, upper) and the diagonal
// mass M of variable `var` along `axis` (hostprep._pencil_km), m entries each.  The tensor-core path solves the z
// direction with them (Thomas sweeps per horizontal mode) instead of transforming along z.
extern "C" int tfb_fdm_set_pencil(tfb_ctx* c, int var, int axis, int m, const double* lower, const double* diag,
                                  const double* upper, const double* mass) {
    TFB_CHECK(c && var >= 0 && var < TFB_MAXVAR && axis >= 0 && axis < 3 && m > 0 && lower && diag && upper && mass, "bad arguments");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    tfb_solver_state* s = solver_of(c);
    FdmVar& f = s->var[var];
    const int n = axis == 0 ? c->desc.nx : axis == 1 ? c->desc.ny : c->desc.nz;
    TFB_CHECK(m <= n, "pencil longer than the grid");
    if (f.pencil_n[axis] != n) {
        cudaFree(f.pencil[axis]);
        f.pencil[axis] = nullptr;
        TFB_CUDA(cudaMalloc(&f.pencil[axis], sizeof(double) * 4 * n));
        f.pencil_n[axis] = n;
    }
    std::vector<double> tab((size_t)4 * n, 0.0);
    for (int i = 0; i < m; i++) { tab[i] = lower[i]; tab[n + i] = diag[i]; tab[2 * n + i] = upper[i]; tab[3 * n + i] = mass[i]; }
    TFB_CUDA(cudaMemcpy(f.pencil[axis], tab.data(), sizeof(double) * 4 * n, cudaMemcpyHostToDevice));
    f.pencil_h[axis] = tab;
    f.th_dirty = true;
    return 0;
}

extern "C" int tfb_fdm_pin(tfb_ctx* c, int var, int64_t cell, double sign) {
    TFB_CHECK(c && var >= 0 && var < TFB_MAXVAR, "bad arguments");
    tfb_solver_state* s = solver_of(c);
    s->var[var].pin_cell = cell;
    s->var[var].pin_sign = sign;
    return 0;
}

extern "C" int tfb_joint_set(tfb_ctx* c, int wvar, int svar, int nz, const double* zops) {
    TFB_CHECK(c && zops && nz == c->desc.nz && nz > 2, "bad arguments");
    TFB_CHECK(c->desc.dim == 3 && wvar == 2 && svar > c->desc.dim && svar < c->desc.dof, "coupled solve: w and a scalar of a 3-D problem");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    tfb_solver_state* s = solver_of(c);
    if (!s->d_jz) {
        TFB_CUDA(cudaMalloc(&s->d_jz, sizeof(double) * TFB_JZ_ROWS * nz));
        TFB_CUDA(cudaMemset(s->d_jz, 0, sizeof(double) * TFB_JZ_ROWS * nz));
    }
    TFB_CUDA(cudaMemcpy(s->d_jz, zops, sizeof(double) * 8 * nz, cudaMemcpyHostToDevice));
    s->joint_w = wvar; s->joint_s = svar;
    s->joint_ready = true;
    s->jz_owner = nullptr;
    return 0;
}

static int ensure_buffers(tfb_ctx* c, int krylov, bool single = false) {
    tfb_solver_state* s = solver_of(c);
    const long long n = c->n_local, ncell = n / c->desc.dof;
    if (!s->d_mass) {
        TFB_CUDA(cudaMalloc(&s->d_mass, sizeof(double) * n));
        std::vector<double> diag(n);
        if (tfb_mass_diag(c, diag.data())) return -1;
        for (auto& d : diag) if (d == 0.0) d = 1.0;
        TFB_CUDA(cudaMemcpy(s->d_mass, diag.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
        for (auto& p : s->comp) TFB_CUDA(cudaMalloc(&p, sizeof(double) * ncell));
        for (int i = 0; i < 8; i++) {
            const size_t len = (size_t)n + 2 * (size_t)c->plane_rows;
            TFB_CUDA(cudaMalloc(&s->vec_base[i], sizeof(double) * len));
            TFB_CUDA(cudaMemset(s->vec_base[i], 0, sizeof(double) * len));
            s->vec[i] = s->vec_base[i] + c->plane_rows;
        }
        TFB_CUDA(cudaMalloc(&s->d_scal, sizeof(double) * 128));
    }
    if (krylov > s->cap || (krylov > 0 && single != s->basis_single)) {
        cudaFree(s->d_V); cudaFree(s->d_Z); cudaFree(s->d_h);
        s->d_V = s->d_Z = s->d_h = nullptr;
        s->cap = 0;
        size_t freeb = 0, total = 0;
        TFB_CUDA(cudaMemGetInfo(&freeb, &total));
        const size_t vbytes = single ? sizeof(float) : sizeof(double);
        const size_t nn = (size_t)n + 2 * (size_t)c->plane_rows;      // room for the halo planes of every vector (idr_run)
        const size_t need = nn * ((size_t)krylov + 1) * vbytes + sizeof(double) * (size_t)n * krylov;
        TFB_CHECK(need < freeb * 0.9, "Krylov basis does not fit in device memory; lower 'Restart'");
        TFB_CUDA(cudaMalloc(&s->d_V, vbytes * nn * (krylov + 1)));
        s->dv_stride = 0;
        s->basis_single = single;
        TFB_CUDA(cudaMalloc(&s->d_Z, sizeof(double) * (size_t)n * krylov));
        TFB_CUDA(cudaMalloc(&s->d_h, sizeof(double) * (krylov + 8)));
        s->cap = krylov;
    }
    return 0;
}

// h[0..nv) = V^T w   (one pass over the basis)
template <class BT>
static int multi_dot(tfb_ctx* c, const BT* V, int nv, const double* w, double* d_out, long long ld, bool reduce) {
    const long long n = c->n_local;
    if (ld <= 0) ld = n;
    TFB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * nv, c->stream));
    const unsigned nb = (unsigned)std::min<long long>((n + 1023) / 1024, 148 * 8);
    for (int v0 = 0; v0 < nv; v0 += 512) {     // 8 warps x 512 partial sums = 32 KB of shared memory
        const int cnt = std::min(512, nv - v0);
        k_all_dots<BT><<<nb, 256, sizeof(double) * 8 * cnt, c->stream>>>(n, V + (size_t)v0 * ld, ld, cnt, w, d_out + v0);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    return reduce ? tfb_allreduce_sum(c, d_out, nv) : 0;
}
// w += sign * V h   (one pass over the basis); d_nrm2 (optional, zeroed here) receives the LOCAL |w|^2
template <class BT>
static int multi_axpy(tfb_ctx* c, const BT* V, int nv, const double* d_h, double sign, double* w, double* d_nrm2,
                      const double* base, double bscale, double wscale, long long ld, bool reduce) {
    const long long n = c->n_local;
    if (ld <= 0) ld = n;
    TFB_CHECK((!base && wscale == 1.0) || nv <= 2048, "the general form of multi_axpy is single-chunk");
    if (nv == 0 && (base || wscale != 1.0)) {     // no basis vectors: only the scaling / base term
        k_all_axpy<BT><<<vec_blocks(n), 256, sizeof(double), c->stream>>>(n, V, ld, 0, d_h, sign, w, d_nrm2, base, bscale, wscale);
        TFB_LAUNCHED();
    }
    if (d_nrm2) TFB_CUDA(cudaMemsetAsync(d_nrm2, 0, sizeof(double), c->stream));
    for (int v0 = 0; v0 < nv; v0 += 2048) {
        const int cnt = std::min(2048, nv - v0);
        const bool last = v0 + cnt >= nv;
        k_all_axpy<BT><<<vec_blocks(n), 256, sizeof(double) * cnt, c->stream>>>(n, V + (size_t)v0 * ld, ld, cnt, d_h + v0, sign, w,
                                                                          last ? d_nrm2 : nullptr, base, bscale, wscale);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    if (d_nrm2 && reduce) return tfb_allreduce_sum(c, d_nrm2, 1);
    return 0;
}

extern "C" int tfb_spmv(tfb_mat* m, const double* x, double* y) {
    TFB_CHECK(m && x && y, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CHECK(c->nranks == 1, "host-vector spmv is single-GPU only");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (ensure_buffers(c, 0)) return -1;
    tfb_solver_state* s = c->solver;
    TFB_CUDA(cudaMemcpyAsync(s->vec[0], x, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    if (spmv(c, m, s->vec[0], s->vec[1], -1)) return -1;
    TFB_CUDA(cudaMemcpyAsync(y, s->vec[1], sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int tfb_precond_apply(tfb_mat* m, const double* r, double* z, int pressure_row) {
    TFB_CHECK(m && r && z, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (ensure_buffers(c, 0)) return -1;
    if (dist_setup(c)) return -1;
    tfb_solver_state* s = c->solver;
    TFB_CUDA(cudaMemcpyAsync(s->vec[4], r, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    s->joint_on = s->joint_ready;
    if (sub_refresh(c, m, pressure_row)) return -1;
    if (apply_precond(c, m, pressure_row, s->vec[4], s->vec[5])) return -1;
    TFB_CUDA(cudaMemcpyAsync(z, s->vec[5], sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// diagnostics: the coupled (w, T) solve alone, host vectors; rows of other variables come back zero.
// table_out (optional): the TFB_JZ_ROWS x nz coefficient table after the refresh from `m`.
extern "C" int tfb_joint_apply(tfb_mat* m, const double* r, double* z, double* table_out) {
    TFB_CHECK(m && r && z, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (ensure_buffers(c, 0)) return -1;
    tfb_solver_state* s = c->solver;
    TFB_CHECK(s->joint_ready, "tfb_joint_set was not called");
    TFB_CUDA(cudaMemcpyAsync(s->vec[4], r, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    TFB_CUDA(cudaMemsetAsync(s->vec[5], 0, sizeof(double) * c->n_local, c->stream));
    if (joint_refresh(c, m)) return -1;
    if (joint_solve(c, s->vec[4], s->vec[5])) return -1;
    TFB_CUDA(cudaMemcpyAsync(z, s->vec[5], sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    if (table_out)
        TFB_CUDA(cudaMemcpyAsync(table_out, s->d_jz, sizeof(double) * TFB_JZ_ROWS * c->desc.nz, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// FGMRES driver.  BT = storage type of the Krylov basis V: double, or float ("compressed basis":
// all arithmetic stays fp64, the basis is only STORED in fp32, which halves the traffic of the
// orthogonalisation; the residual estimate is then only trusted up to ~1e-6 per cycle and every
// cycle restarts from the true fp64 residual).  Z (preconditioned directions) is always fp64.
template <class BT>
static int fgmres_run(tfb_mat* m, const double* b, double* x, const tfb_solve_opts* o, tfb_solve_info* info) {
    tfb_ctx* c = m->ctx;
    const long long n = c->n_local;
    const int mk = std::max(1, std::min(o->restart, o->maxit));
    constexpr bool SINGLE = sizeof(BT) == 4;
    if (ensure_buffers(c, mk, SINGLE)) return -1;
    tfb_solver_state* s = c->solver;
    s->dv_stride = 0;
    const int prow = o->pressure_row;
    cudaEvent_t e0, e1;
    TFB_CUDA(cudaEventCreate(&e0)); TFB_CUDA(cudaEventCreate(&e1));
    TFB_CUDA(cudaEventRecord(e0, c->stream));

    if (sub_refresh(c, m, prow)) return -1;
    double* d_b = s->vec[4];
    double* d_x = s->vec[5];
    double* w = s->vec[6];        // vector being orthogonalised (fp64)
    double* v64 = s->vec[7];      // fp64 copy of the current basis vector when the basis is fp32
    TFB_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    TFB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
    BT* V = reinterpret_cast<BT*>(s->d_V);
    double* Z = s->d_Z;
    double* d_h = s->d_h;
    std::vector<double> H((size_t)(mk + 1) * mk, 0.0), g(mk + 1), cs(mk), sn(mk), hcol(mk + 2), y(mk);
    auto Hx = [&](int i, int j) -> double& { return H[(size_t)j * (mk + 1) + i]; };

    int total_its = 0, converged = 0, reorth = 0, cycles = 0, stalled = 0;
    const int stall_limit = o->stall_cycles > 0 ? o->stall_cycles : 3;
    const bool prof = o->verbose >= 1;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (prof) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); evs.push_back(e); } };
    double bnorm = 0.0, relres = 1.0, prev_true = 1e300;
    {
        if (multi_dot<double>(c, d_b, 1, d_b, d_h)) return -1;
        TFB_CUDA(cudaMemcpyAsync(&bnorm, d_h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        bnorm = sqrt(bnorm);
    }
    if (bnorm == 0.0) {
        memset(x, 0, sizeof(double) * n);
        if (info) { info->iters = 0; info->converged = 1; info->relres = 0.0; info->setup_ms = info->solve_ms = 0; }
        return 0;
    }
    // with an fp32 basis the estimate of a cycle is good for ~6 digits: end the cycle there
    const double cycle_gain = SINGLE ? 1e-6 : 0.0;
    while (total_its < o->maxit && !converged) {
        // true residual r = b - J x  -> w
        if (total_its == 0) {
            TFB_CUDA(cudaMemcpyAsync(w, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            if (spmv(c, m, d_x, s->vec[0], prow)) return -1;
            k_sub<<<vec_blocks(n), 256, 0, c->stream>>>(n, d_b, s->vec[0], w);
            TFB_LAUNCHED();
        }
        if (multi_dot<double>(c, w, 1, w, d_h)) return -1;
        double beta = 0.0;
        TFB_CUDA(cudaMemcpyAsync(&beta, d_h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        beta = sqrt(beta);
        relres = beta / bnorm;
        if (relres <= o->tol) { converged = 1; break; }
        // stagnation: several restart cycles in a row without a 2x reduction of the true residual
        if (cycles > 0 && relres > 0.5 * prev_true) { if (++stalled >= stall_limit) break; } else stalled = 0;
        prev_true = std::min(prev_true, relres);
        cycles++;
        TFB_CUDA(cudaMemcpyAsync(s->d_scal + 4, &beta, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        k_store_scaled<BT><<<vec_blocks(n), 256, 0, c->stream>>>(n, s->d_scal, 4, w, V);
        TFB_LAUNCHED();
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        const double cycle_target = std::max(o->tol * bnorm, cycle_gain * beta);
        int j = 0;
        for (; j < mk && total_its < o->maxit; j++, total_its++) {
            const BT* vjb = V + (size_t)j * n;
            const double* vj;
            if (SINGLE) {
                k_widen<<<vec_blocks(n), 256, 0, c->stream>>>(n, reinterpret_cast<const float*>(vjb), v64);
                TFB_LAUNCHED();
                vj = v64;
            } else {
                vj = reinterpret_cast<const double*>(vjb);
            }
            double* zj = Z + (size_t)j * n;
            mark();
            if (apply_precond(c, m, prow, vj, zj)) return -1;
            mark();
            if (spmv(c, m, zj, w, prow)) return -1;
            mark();
            // classical Gram-Schmidt with single-pass fused kernels; the second sweep only runs when
            // the first one cancelled most of w (DGKS criterion)
            if (multi_dot<double>(c, w, 1, w, d_h + mk + 2)) return -1;            // |w|^2 before
            if (multi_dot<BT>(c, V, j + 1, w, d_h)) return -1;
            if (multi_axpy<BT>(c, V, j + 1, d_h, -1.0, w, d_h + mk + 1)) return -1;   // fused |w|^2 after
            double nrm[2];
            TFB_CUDA(cudaMemcpyAsync(hcol.data(), d_h, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, c->stream));
            TFB_CUDA(cudaMemcpyAsync(nrm, d_h + mk + 1, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
            TFB_CUDA(cudaStreamSynchronize(c->stream));
            std::vector<double> h2(j + 1, 0.0);
            double hn2 = nrm[0];
            if (nrm[0] < 0.25 * nrm[1]) {
                if (multi_dot<BT>(c, V, j + 1, w, d_h)) return -1;
                if (multi_axpy<BT>(c, V, j + 1, d_h, -1.0, w, d_h + mk + 1)) return -1;
                TFB_CUDA(cudaMemcpyAsync(h2.data(), d_h, sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, c->stream));
                TFB_CUDA(cudaMemcpyAsync(&hn2, d_h + mk + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                TFB_CUDA(cudaStreamSynchronize(c->stream));
                reorth++;
            }
            const double hn = sqrt(hn2);
            mark();
            for (int i = 0; i <= j; i++) Hx(i, j) = hcol[i] + h2[i];
            Hx(j + 1, j) = hn;
            if (hn > 0.0) {
                TFB_CUDA(cudaMemcpyAsync(s->d_scal + 4, &hn, sizeof(double), cudaMemcpyHostToDevice, c->stream));
                k_store_scaled<BT><<<vec_blocks(n), 256, 0, c->stream>>>(n, s->d_scal, 4, w, V + (size_t)(j + 1) * n);
                TFB_LAUNCHED();
            }
            // Givens rotations
            for (int i = 0; i < j; i++) {
                const double t = cs[i] * Hx(i, j) + sn[i] * Hx(i + 1, j);
                Hx(i + 1, j) = -sn[i] * Hx(i, j) + cs[i] * Hx(i + 1, j);
                Hx(i, j) = t;
            }
            const double d = hypot(Hx(j, j), Hx(j + 1, j));
            cs[j] = Hx(j, j) / d; sn[j] = Hx(j + 1, j) / d;
            Hx(j, j) = d; Hx(j + 1, j) = 0.0;
            g[j + 1] = -sn[j] * g[j];
            g[j] = cs[j] * g[j];
            relres = fabs(g[j + 1]) / bnorm;
            if (o->verbose > 1) fprintf(stderr, "  fgmres %4d  relres %.3e\n", total_its + 1, relres);
            if (fabs(g[j + 1]) <= cycle_target || hn == 0.0) { j++; total_its++; break; }
        }
        // x += Z y with H y = g; convergence is decided on the TRUE residual at the top of the loop
        const int k = j;
        for (int i = k - 1; i >= 0; i--) {
            double acc = g[i];
            for (int l = i + 1; l < k; l++) acc -= Hx(i, l) * y[l];
            y[i] = acc / Hx(i, i);
        }
        TFB_CUDA(cudaMemcpyAsync(d_h, y.data(), sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
        if (multi_axpy<double>(c, Z, k, d_h, 1.0, d_x)) return -1;
        TFB_CUDA(cudaStreamSynchronize(c->stream));
    }
    // true residual of the returned iterate
    if (spmv(c, m, d_x, s->vec[0], prow)) return -1;
    k_sub<<<vec_blocks(n), 256, 0, c->stream>>>(n, d_b, s->vec[0], s->vec[1]);
    TFB_LAUNCHED();
    if (multi_dot<double>(c, s->vec[1], 1, s->vec[1], d_h)) return -1;
    double rr = 0.0;
    TFB_CUDA(cudaMemcpyAsync(&rr, d_h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaMemcpyAsync(x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaEventRecord(e1, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    relres = sqrt(rr) / bnorm;
    if (prof && evs.size() >= 4) {
        double t[3] = {0, 0, 0};
        for (size_t e = 0; e + 3 < evs.size(); e += 4)
            for (int ph = 0; ph < 3; ph++) { float x_ms = 0; cudaEventElapsedTime(&x_ms, evs[e + ph], evs[e + ph + 1]); t[ph] += x_ms; }
        fprintf(stderr, "tfb_solve: %d its in %d cycle(s), %.1f ms total: precond %.1f ms, operator %.1f ms, orthogonalisation %.1f ms, %d re-orth sweeps, %s basis\n",
                total_its, cycles, ms, t[0], t[1], t[2], reorth, SINGLE ? "fp32" : "fp64");
    }
    for (auto e : evs) cudaEventDestroy(e);
    if (info) {
        info->iters = total_its;
        info->converged = relres <= o->tol * 1.0001;
        info->relres = relres;
        info->setup_ms = (float)reorth;   // number of second Gram-Schmidt sweeps (diagnostic)
        info->solve_ms = ms;
    }
    return relres <= o->tol * 1.0001 ? 0 : 1;
}

// out = a*x + b*y + c*z (null pointers are skipped)
__global__ void k_lincomb(long long n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                          double cc, const double* __restrict__ z, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = a * x[i];
        if (y) v += b * y[i];
        if (z) v += cc * z[i];
        out[i] = v;
    }
}

// Right-preconditioned BiCGStab (the preconditioner is a fixed linear operator): 7 work vectors instead
// of a Krylov basis -- the memory-light alternative when even 180 GB cannot hold an un-restarted GMRES
// basis.  Two preconditioner applications and two operator products per iteration; convergence is
// confirmed on the true residual, from which the recurrence is restarted if necessary.
static int bicgstab_run(tfb_mat* m, const double* b, double* x, const tfb_solve_opts* o, tfb_solve_info* info) {
    tfb_ctx* c = m->ctx;
    const long long n = c->n_local;
    if (ensure_buffers(c, 5, false)) return -1;
    tfb_solver_state* s = c->solver;
    s->dv_stride = 0;
    const int prow = o->pressure_row;
    cudaEvent_t e0, e1;
    TFB_CUDA(cudaEventCreate(&e0)); TFB_CUDA(cudaEventCreate(&e1));
    TFB_CUDA(cudaEventRecord(e0, c->stream));
    if (sub_refresh(c, m, prow)) return -1;
    double* d_b = s->vec[4];
    double* d_x = s->vec[5];
    double* W = s->d_V;   // 6 vectors
    double *r = W, *rh = W + n, *p = W + 2 * n, *v = W + 3 * n, *sv = W + 4 * n, *t = W + 5 * n;
    double *y = s->d_Z, *z = s->d_Z + n;
    double* d_h = s->d_h;
    const unsigned nb = vec_blocks(n);
    TFB_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    TFB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
    auto dot = [&](const double* a, const double* bb, double& out) -> int {
        if (multi_dot<double>(c, a, 1, bb, d_h)) return -1;
        TFB_CUDA(cudaMemcpyAsync(&out, d_h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    double bn2 = 0.0;
    if (dot(d_b, d_b, bn2)) return -1;
    const double bnorm = sqrt(bn2);
    int its = 0, converged = 0, stalled = 0;
    const int stall_limit = o->stall_cycles > 0 ? o->stall_cycles : 3;
    double relres = 1.0;
    if (bnorm == 0.0) {
        memset(x, 0, sizeof(double) * n);
        if (info) { info->iters = 0; info->converged = 1; info->relres = 0.0; info->setup_ms = info->solve_ms = 0; }
        return 0;
    }
    double prev_true = 1e300;
    while (its < o->maxit && !converged) {
        // (re)start from the true residual
        if (its == 0) TFB_CUDA(cudaMemcpyAsync(r, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        else {
            if (spmv(c, m, d_x, s->vec[0], prow)) return -1;
            k_sub<<<nb, 256, 0, c->stream>>>(n, d_b, s->vec[0], r);
            TFB_LAUNCHED();
        }
        double rr = 0.0;
        if (dot(r, r, rr)) return -1;
        relres = sqrt(rr) / bnorm;
        if (relres <= o->tol) { converged = 1; break; }
        if (relres > 0.5 * prev_true && its > 0) { if (++stalled >= stall_limit) break; } else stalled = 0;
        prev_true = std::min(prev_true, relres);
        TFB_CUDA(cudaMemcpyAsync(rh, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        TFB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * n, c->stream));
        TFB_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * n, c->stream));
        double rho = 1.0, alpha = 1.0, omega = 1.0;
        for (; its < o->maxit; its++) {
            double rho_new = 0.0;
            if (dot(rh, r, rho_new)) return -1;
            if (rho_new == 0.0 || omega == 0.0) break;                      // breakdown: restart
            const double beta = (rho_new / rho) * (alpha / omega);
            k_lincomb<<<nb, 256, 0, c->stream>>>(n, 1.0, r, beta, p, -beta * omega, v, p);   // p = r + beta (p - omega v)
            TFB_LAUNCHED();
            if (apply_precond(c, m, prow, p, y)) return -1;
            if (spmv(c, m, y, v, prow)) return -1;
            double rhv = 0.0;
            if (dot(rh, v, rhv)) return -1;
            if (rhv == 0.0) break;
            alpha = rho_new / rhv;
            k_lincomb<<<nb, 256, 0, c->stream>>>(n, 1.0, r, -alpha, v, 0.0, nullptr, sv);          // s = r - alpha v
            TFB_LAUNCHED();
            double ss = 0.0;
            if (dot(sv, sv, ss)) return -1;
            if (sqrt(ss) / bnorm <= o->tol) {
                k_axpy<<<nb, 256, 0, c->stream>>>(n, alpha, y, d_x);
                TFB_LAUNCHED();
                its++;
                break;
            }
            if (apply_precond(c, m, prow, sv, z)) return -1;
            if (spmv(c, m, z, t, prow)) return -1;
            double ts = 0.0, tt = 0.0;
            if (dot(t, sv, ts) || dot(t, t, tt)) return -1;
            if (tt == 0.0) break;
            omega = ts / tt;
            k_lincomb<<<nb, 256, 0, c->stream>>>(n, 1.0, d_x, alpha, y, omega, z, d_x);            // x += alpha y + omega z
            k_lincomb<<<nb, 256, 0, c->stream>>>(n, 1.0, sv, -omega, t, 0.0, nullptr, r);          // r = s - omega t
            TFB_LAUNCHED(); TFB_LAUNCHED();
            rho = rho_new;
            double rn = 0.0;
            if (dot(r, r, rn)) return -1;
            relres = sqrt(rn) / bnorm;
            if (o->verbose > 1) fprintf(stderr, "  bicgstab %4d  relres %.3e\n", its + 1, relres);
            if (relres <= o->tol) { its++; break; }
        }
    }
    if (spmv(c, m, d_x, s->vec[0], prow)) return -1;
    k_sub<<<nb, 256, 0, c->stream>>>(n, d_b, s->vec[0], s->vec[1]);
    TFB_LAUNCHED();
    double rr = 0.0;
    if (dot(s->vec[1], s->vec[1], rr)) return -1;
    TFB_CUDA(cudaMemcpyAsync(x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaEventRecord(e1, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    relres = sqrt(rr) / bnorm;
    if (o->verbose >= 1) fprintf(stderr, "tfb_solve: BiCGStab %d its, %.1f ms, true relres %.2e\n", its, ms, relres);
    if (info) {
        info->iters = its; info->converged = relres <= o->tol * 1.0001; info->relres = relres;
        info->setup_ms = 0.f; info->solve_ms = ms;
    }
    return relres <= o->tol * 1.0001 ? 0 : 1;
}


// Shadow vectors of IDR(s) are never stored: entry (vector j, GLOBAL row g) is +1 or -1 according to bit j of a 64-bit hash
// of g (Rademacher vectors), so a z-slab run uses the same shadow space as a single-GPU run and the s dot products P^T w
// cost one pass over w -- s sign flips and additions per entry -- instead of s + 1 vector reads (and s vectors of HBM).
__device__ __forceinline__ unsigned long long tfb_shadow_hash(unsigned long long grow) {
    unsigned long long h = grow * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return h;
}
template <int NS>   // NS = 8 or 16 accumulators
__global__ void __launch_bounds__(256) k_shadow_dots(long long n, long long row0, int nv, const double* __restrict__ w,
                                                     double* __restrict__ out) {
    double acc[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) acc[j] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double wv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) wv[u] = i0 + u * stride < n ? w[i0 + u * stride] : 0.0;      // four loads in flight
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long wbits = __double_as_longlong(wv[u]);
            const unsigned long long h = tfb_shadow_hash((unsigned long long)(row0 + i0 + u * stride));
#pragma unroll
            for (int j = 0; j < NS; j++)
                acc[j] += __longlong_as_double(wbits ^ (long long)(((h >> (2 * j + 7)) & 1ull) << 63));
        }
    }
    __shared__ double red[8][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NS; j++) {
        double a = acc[j];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[warp][j] = a;
    }
    __syncthreads();
    if (threadIdx.x < nv) {
        double t = 0.0;
        for (int wv = 0; wv < 8; wv++) t += red[wv][threadIdx.x];
        atomicAdd(&out[threadIdx.x], t);
    }
}
// out[0..S) = P^T w, local part (the caller all-reduces it together with whatever else is pending)
static int shadow_dots(tfb_ctx* c, int S, const double* w, double* d_out) {
    const long long n = c->n_local;
    TFB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * S, c->stream));
    const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148 * 8);
    if (S <= 8) k_shadow_dots<8><<<nb, 256, 0, c->stream>>>(n, c->row0, S, w, d_out);
    else k_shadow_dots<16><<<nb, 256, 0, c->stream>>>(n, c->row0, S, w, d_out);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// One sweep of the IDR recurrence over a family V (the G or the U vectors):
//   v_k -= sum_{j<k} al[j] V_j ;   tgt += bsig * v_k ;   optionally nrm2 += |tgt|^2 (local part)
// i.e. the bi-orthogonalisation of the new vector and the update of the residual (or the iterate) in one pass.
__global__ void __launch_bounds__(256) k_idr_sweep(long long n, const double* __restrict__ V, long long ld, int k,
                                                   const double* __restrict__ al, double* __restrict__ vk, double bsig,
                                                   double* __restrict__ tgt, double* __restrict__ nrm2) {
    __shared__ double als[16];
    __shared__ double red[8];
    if (threadIdx.x < k) als[threadIdx.x] = al[threadIdx.x];
    __syncthreads();
    double loc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double a = vk[i];
        for (int j = 0; j < k; j++) a -= als[j] * V[(long long)j * ld + i];
        if (k > 0) vk[i] = a;
        const double tn = tgt[i] + bsig * a;
        tgt[i] = tn;
        loc += tn * tn;
    }
    if (nrm2) {
        for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int wv = 0; wv < 8; wv++) t += red[wv];
            atomicAdd(nrm2, t);
        }
    }
}

// Right-preconditioned IDR(s) in the bi-orthogonal form (van Gijzen & Sonneveld, ACM TOMS 38, 2011).  Short
// recurrences: 3s+5 vectors and O(s) vector passes per operator product instead of a Krylov basis that GMRES reads
// twice per iteration (at 128^3 the orthogonalisation is 45% of an un-restarted FGMRES solve).  Needs a FIXED
// preconditioner, so it is not combined with the inner iterations ('Velocity Iterations', coupled (w,T) solve).
// The bi-orthogonalisation of a new g-vector against the shadow space uses ONE fused multi-dot: with M = P^T G lower
// triangular the recursive coefficients follow from the s dot products by forward substitution on the host.
// Convergence is confirmed on the true residual, from which the recurrence is restarted if necessary.
static int idr_run(tfb_mat* m, const double* b, double* x, const tfb_solve_opts* o, tfb_solve_info* info, int sdim) {
    tfb_ctx* c = m->ctx;
    const long long n = c->n_local;
    const int S = std::max(1, std::min(sdim, 16));
    if (ensure_buffers(c, 2 * S + 3, false)) return -1;
    tfb_solver_state* s = c->solver;
    const int prow = o->pressure_row;
    cudaEvent_t e0, e1;
    TFB_CUDA(cudaEventCreate(&e0)); TFB_CUDA(cudaEventCreate(&e1));
    TFB_CUDA(cudaEventRecord(e0, c->stream));
    if (sub_refresh(c, m, prow)) return -1;
    double* d_b = s->vec[4];
    double* d_x = s->vec[5];
    double* vh = s->vec[6];     // preconditioned vector
    double* tmp = s->vec[7];
    // every vector of the recurrence sits between two planes of slack (stride nn), so each can be the input of a
    // z-slab operator product without a ghosted copy
    const long long nn = n + 2 * c->plane_rows;
    s->dv_stride = nn; s->dv_count = 2 * S + 3;
    double* G = s->d_V + c->plane_rows;       // S vectors
    double* U = G + (size_t)S * nn;           // S vectors
    double* r = U + (size_t)S * nn;           // r and t adjacent: one fused pass gives (r.t, t.t)
    double* t = r + nn;
    double* v = t + nn;
    // device scalars: [0,S) dot products, [S] the pending |r|^2 (local part until the next reduction), then S coefficients
    const int SD = std::max(S, 2);            // the omega step needs two dot products
    double* d_dot = s->d_h;
    double* d_nrm = s->d_h + SD;
    double* d_coef = s->d_h + SD + 1;
    const unsigned nb = vec_blocks(n);
    TFB_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    TFB_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * n, c->stream));
    auto fetch = [&](double* host, const double* dev, int cnt) -> int {
        TFB_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    auto put = [&](double* dev, const double* host, int cnt) -> int {
        TFB_CUDA(cudaMemcpyAsync(dev, host, sizeof(double) * cnt, cudaMemcpyHostToDevice, c->stream));
        return 0;
    };
    double bn2 = 0.0;
    if (multi_dot<double>(c, d_b, 1, d_b, d_dot) || fetch(&bn2, d_dot, 1)) return -1;
    std::vector<double> hbuf(SD + 1);
    // all-reduce the S dot products and the pending residual norm in one go, fetch them with one synchronisation
    auto reduce_fetch = [&]() -> int {
        if (tfb_allreduce_sum(c, d_dot, SD + 1)) return -1;
        return fetch(hbuf.data(), d_dot, SD + 1);
    };
    const double bnorm = sqrt(bn2);
    if (bnorm == 0.0) {
        memset(x, 0, sizeof(double) * n);
        if (info) { info->iters = 0; info->converged = 1; info->relres = 0.0; info->setup_ms = info->solve_ms = 0; }
        return 0;
    }
    // per-phase device timers (Verbose): events at the phase boundaries of every product, summed after the solve
    enum { PH_FORM_V = 0, PH_PRECOND, PH_FORM_U, PH_OPERATOR, PH_DOTS, PH_BIORTH, PH_UPDATE, PH_COUNT };
    const bool prof = o->verbose >= 1;
    std::vector<cudaEvent_t> pev;
    std::vector<int> pph;
    auto mark = [&](int phase_that_ended) {
        if (!prof) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream);
        pev.push_back(e); pph.push_back(phase_that_ended);
    };
    int its = 0, converged = 0, cycles = 0, stalled = 0;
    // IDR is only ever the first attempt of 'auto' (FGMRES is the fallback), so it gives up after one stalled restart
    // unless the caller asks for more
    const int stall_limit = o->stall_cycles > 0 ? o->stall_cycles : 1;
    double relres = 1.0, prev_true = 1e300;
    std::vector<double> M((size_t)S * S), f(S), cf(S), d(S), al(S);
    auto Mx = [&](int i, int j) -> double& { return M[(size_t)i * S + j]; };
    // One application of the fixed preconditioner is ~45 short launches (30 of them cuBLAS calls whose host side costs
    // about as much as the 20 us kernel they start) between two host synchronisations of the recurrence: on one GPU,
    // with no inner iteration, it has no host dependence, so the second application with the same (in, out) pair is
    // captured into a CUDA graph and replayed for the rest of the solve.  The first one runs eagerly (lazy module loads).  Slot 0: v -> vh, slot 1: r -> vh.  TFB_NO_GRAPH=1 keeps the eager path.
    static int graphs_allowed = -1;
    if (graphs_allowed < 0) { const char* e = getenv("TFB_NO_GRAPH"); graphs_allowed = !(e && e[0] == '1'); }
    bool use_graph = graphs_allowed && c->nranks == 1 && s->inner_its <= 0 && !s->joint_on;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int64_t gnodes[2] = {0, 0};
    int gseen[2] = {0, 0};
    auto precond = [&](const double* in, double* out, int slot) -> int {
        if (!use_graph) return apply_precond(c, m, prow, in, out);
        if (gexec[slot]) {
            TFB_CUDA(cudaGraphLaunch(gexec[slot], c->stream));
            g_tfb_launches += gnodes[slot];
            return 0;
        }
        if (gseen[slot]++ == 0) return apply_precond(c, m, prow, in, out);
        const int64_t before = g_tfb_launches;
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            use_graph = false;
            return apply_precond(c, m, prow, in, out);
        }
        const int rc = apply_precond(c, m, prow, in, out);
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (rc == 0 && e == cudaSuccess && g) e = cudaGraphInstantiate(&gexec[slot], g, 0);
        if (g) cudaGraphDestroy(g);
        if (rc != 0 || e != cudaSuccess || !gexec[slot]) {     // not capturable here: eager for the rest of the solve
            cudaGetLastError();
            gexec[slot] = nullptr;
            use_graph = false;
            g_tfb_launches = before;
            return apply_precond(c, m, prow, in, out);
        }
        gnodes[slot] = g_tfb_launches - before;
        TFB_CUDA(cudaGraphLaunch(gexec[slot], c->stream));
        return 0;
    };
    struct GraphGuard {
        cudaGraphExec_t* g;
        ~GraphGuard() { for (int i = 0; i < 2; i++) if (g[i]) cudaGraphExecDestroy(g[i]); }
    } graph_guard{gexec};
    while (its < o->maxit && !converged) {
        // (re)start from the true residual
        if (its == 0) TFB_CUDA(cudaMemcpyAsync(r, d_b, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        else {
            if (spmv(c, m, d_x, tmp, prow)) return -1;
            k_sub<<<nb, 256, 0, c->stream>>>(n, d_b, tmp, r);
            TFB_LAUNCHED();
        }
        double rr = 0.0;
        if (multi_dot<double>(c, r, 1, r, d_dot) || fetch(&rr, d_dot, 1)) return -1;
        relres = sqrt(rr) / bnorm;
        if (relres <= o->tol) { converged = 1; break; }
        // stagnation: several restarts in a row without a 2x reduction of the true residual
        if (cycles > 0 && relres > 0.5 * prev_true) { if (++stalled >= stall_limit) break; } else stalled = 0;
        prev_true = std::min(prev_true, relres);
        cycles++;
        TFB_CUDA(cudaMemsetAsync(G - c->plane_rows, 0, sizeof(double) * (size_t)2 * S * nn, c->stream));   // G and U
        std::fill(M.begin(), M.end(), 0.0);
        for (int i = 0; i < S; i++) Mx(i, i) = 1.0;
        double om = 1.0;
        bool done = false, breakdown = false;
        // The norm of the updated residual is left on the device as a local partial sum (`pending`) and is reduced and
        // fetched together with the next set of dot products: one all-reduce and one host synchronisation per operator
        // product; convergence is noticed one product late, and that product's vectors are simply dropped.
        bool pending = false;
        auto check_pending = [&]() {
            if (!pending) return;
            pending = false;
            relres = sqrt(hbuf[SD]) / bnorm;
            if (o->verbose > 1) fprintf(stderr, "  idr(%d) %4d  relres %.3e\n", S, its, relres);
            if (!(relres == relres)) breakdown = true;
            else if (relres <= o->tol) done = true;
        };
        TFB_CUDA(cudaMemsetAsync(d_nrm, 0, sizeof(double), c->stream));
        while (its < o->maxit && !done && !breakdown) {
            if (shadow_dots(c, S, r, d_dot) || reduce_fetch()) return -1;                           // f = P^T r
            for (int i = 0; i < S; i++) f[i] = hbuf[i];
            check_pending();
            if (done || breakdown) break;
            for (int k = 0; k < S && its < o->maxit; k++) {
                // c = M[k:,k:]^-1 f[k:]  (lower triangular)
                for (int i = k; i < S; i++) {
                    double acc = f[i];
                    for (int j = k; j < i; j++) acc -= Mx(i, j) * cf[j];
                    if (Mx(i, i) == 0.0) { breakdown = true; break; }
                    cf[i] = acc / Mx(i, i);
                }
                if (breakdown) break;
                // v = r - sum_{i>=k} c_i G_i
                mark(-1);
                if (put(d_coef, cf.data() + k, S - k)) return -1;
                if (multi_axpy<double>(c, G + (size_t)k * nn, S - k, d_coef, -1.0, v, nullptr, r, 1.0, 0.0, nn)) return -1;
                mark(PH_FORM_V);
                if (precond(v, vh, 0)) return -1;
                mark(PH_PRECOND);
                // U_k = om * vh + sum_{i>=k} c_i U_i   (the old U_k is part of the sum)
                double* Uk = U + (size_t)k * nn;
                double* Gk = G + (size_t)k * nn;
                if (multi_axpy<double>(c, Uk + nn, S - k - 1, d_coef + 1, 1.0, Uk, nullptr, vh, om, cf[k], nn)) return -1;
                mark(PH_FORM_U);
                if (spmv(c, m, Uk, Gk, prow)) return -1;                                            // G_k = A U_k
                mark(PH_OPERATOR);
                its++;
                // bi-orthogonalise against p_0..p_{k-1}: all dots in one pass, recursion on the host
                if (shadow_dots(c, S, Gk, d_dot) || reduce_fetch()) return -1;
                mark(PH_DOTS);
                check_pending();                   // residual after the PREVIOUS product
                if (done || breakdown) { its--; break; }
                for (int i = 0; i < S; i++) d[i] = hbuf[i];
                for (int i = 0; i < k; i++) {
                    double acc = d[i];
                    for (int j = 0; j < i; j++) acc -= al[j] * Mx(i, j);
                    al[i] = acc / Mx(i, i);
                }
                for (int i = k; i < S; i++) {
                    double acc = d[i];
                    for (int j = 0; j < k; j++) acc -= al[j] * Mx(i, j);
                    Mx(i, k) = acc;
                }
                if (Mx(k, k) == 0.0) { breakdown = true; break; }
                const double beta = f[k] / Mx(k, k);
                // G_k -= sum al_j G_j, r -= beta G_k (local |r|^2 pending);  U_k -= sum al_j U_j, x += beta U_k
                if (k > 0 && put(d_coef, al.data(), k)) return -1;
                TFB_CUDA(cudaMemsetAsync(d_nrm, 0, sizeof(double), c->stream));
                k_idr_sweep<<<nb, 256, 0, c->stream>>>(n, G, nn, k, d_coef, Gk, -beta, r, d_nrm);
                k_idr_sweep<<<nb, 256, 0, c->stream>>>(n, U, nn, k, d_coef, Uk, beta, d_x, nullptr);
                TFB_LAUNCHED(); TFB_LAUNCHED();
                TFB_CUDA(cudaGetLastError());
                pending = true;
                mark(PH_UPDATE);
                for (int i = k + 1; i < S; i++) f[i] -= beta * Mx(i, k);
            }
            if (done || breakdown || its >= o->maxit) break;
            // dimension-reduction step: r <- (I - om A Minv) r
            if (precond(r, vh, 1)) return -1;
            if (spmv(c, m, vh, t, prow)) return -1;
            its++;
            if (!pending) {          // |r|^2 is normally still pending from the last sweep
                TFB_CUDA(cudaMemsetAsync(d_nrm, 0, sizeof(double), c->stream));
                if (multi_dot<double>(c, r, 1, r, d_nrm, 0, false)) return -1;
            }
            if (multi_dot<double>(c, r, 2, t, d_dot, nn, false)) return -1;                         // (r.t, t.t)
            if (reduce_fetch()) return -1;
            const double rt = hbuf[0], tt = hbuf[1], rn2 = hbuf[SD];
            pending = true;
            check_pending();
            if (done || breakdown) { its--; break; }
            if (tt == 0.0) { breakdown = true; break; }
            om = rt / tt;
            const double rho = fabs(rt) / (sqrt(tt) * sqrt(rn2));
            if (rho < 0.7 && rho > 0.0) om *= 0.7 / rho;                                             // "maintaining the convergence"
            if (om == 0.0) { breakdown = true; break; }
            k_axpy<<<nb, 256, 0, c->stream>>>(n, om, vh, d_x);
            TFB_LAUNCHED();
            if (put(d_coef, &om, 1)) return -1;
            TFB_CUDA(cudaMemsetAsync(d_nrm, 0, sizeof(double), c->stream));
            if (multi_axpy<double>(c, t, 1, d_coef, -1.0, r, d_nrm, nullptr, 1.0, 1.0, 0, false)) return -1;
            pending = true;
        }
    }
    if (spmv(c, m, d_x, tmp, prow)) return -1;
    k_sub<<<nb, 256, 0, c->stream>>>(n, d_b, tmp, vh);
    TFB_LAUNCHED();
    double rr = 0.0;
    if (multi_dot<double>(c, vh, 1, vh, d_dot)) return -1;
    TFB_CUDA(cudaMemcpyAsync(&rr, d_dot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaMemcpyAsync(x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaEventRecord(e1, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    relres = sqrt(rr) / bnorm;
    if (o->verbose >= 1)
        fprintf(stderr, "tfb_solve: IDR(%d) %d operator products in %d cycle(s), %.1f ms, true relres %.2e\n", S, its, cycles, ms, relres);
    if (prof && pev.size() > 1) {
        double t[PH_COUNT] = {};
        int cnt[PH_COUNT] = {};
        for (size_t e = 1; e < pev.size(); e++) {
            if (pph[e] < 0) continue;
            float x_ms = 0.f;
            cudaEventElapsedTime(&x_ms, pev[e - 1], pev[e]);
            t[pph[e]] += x_ms; cnt[pph[e]]++;
        }
        const char* names[PH_COUNT] = {"form v", "preconditioner", "form U", "operator (+halo)", "shadow dots (+all-reduce, fetch)",
                                       "(unused)", "bi-orthogonalisation + r/x update"};
        fprintf(stderr, "tfb_solve: rank %d of %d, mean device time per product by phase [us]:", c->rank, c->nranks);
        double sum = 0.0;
        for (int ph = 0; ph < PH_COUNT; ph++) {
            const double us = cnt[ph] ? 1e3 * t[ph] / cnt[PH_OPERATOR] : 0.0;
            fprintf(stderr, " %s %.0f;", names[ph], us);
            sum += us;
        }
        fprintf(stderr, " sum %.0f\n", sum);
    }
    for (auto e : pev) cudaEventDestroy(e);
    if (info) {
        info->iters = its; info->converged = relres <= o->tol * 1.0001; info->relres = relres;
        info->setup_ms = 0.f; info->solve_ms = ms;
    }
    return relres <= o->tol * 1.0001 ? 0 : 1;
}

// solver state that follows from the options (shared by tfb_solve and the diagnostics)
static int configure_precond(tfb_ctx* c, const tfb_solve_opts* o) {
    tfb_solver_state* s = solver_of(c);
    if (dist_setup(c)) return -1;
    s->precond_single = (o->precond_flags & TFB_PREC_FP32) != 0;
    s->inner_its = std::max(0, std::min(24, o->inner_its));   // d_scal slice holds 2k+3 <= 56 doubles
    s->inner_total = 0;
    s->joint_on = s->joint_ready && !(o->precond_flags & TFB_PREC_NO_JOINT);
    s->schur_mass = (o->precond_flags & TFB_PREC_SCALED_MASS) != 0;
    bool tc = (o->precond_flags & TFB_PREC_TENSOR) != 0;
    for (int v = 0; tc && v < c->desc.dim; v++) tc = tc_ready(c, v);
    s->precond_tc = tc;
    if (tc) s->precond_single = true;     // every FDM sub-solve works on fp32 arrays and takes the tensor-core path
    if (const char* e = getenv("TFB_INNER_TOL")) s->inner_tol = atof(e);
    return 0;
}

extern "C" int tfb_solve(tfb_mat* m, const double* b, double* x, const tfb_solve_opts* o, tfb_solve_info* info) {
    TFB_CHECK(m && b && x && o, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (configure_precond(c, o)) return -1;
    if (o->method == TFB_METHOD_BICGSTAB) return bicgstab_run(m, b, x, o, info);
    if (o->method == TFB_METHOD_IDR) return idr_run(m, b, x, o, info, o->idr_s > 0 ? o->idr_s : 8);
    if (o->basis_fp32 == 1) return fgmres_run<float>(m, b, x, o, info);
    return fgmres_run<double>(m, b, x, o, info);
}

extern "C" int tfb_precond_apply_opts(tfb_mat* m, const double* r, double* z, const tfb_solve_opts* o) {
    TFB_CHECK(m && r && z && o, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (ensure_buffers(c, 1)) return -1;     // the gamma estimate of the scaled-mass variant uses the dot-product slots
    if (configure_precond(c, o)) return -1;
    tfb_solver_state* s = c->solver;
    TFB_CUDA(cudaMemcpyAsync(s->vec[4], r, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    if (sub_refresh(c, m, o->pressure_row)) return -1;
    if (apply_precond(c, m, o->pressure_row, s->vec[4], s->vec[5])) return -1;
    TFB_CUDA(cudaMemcpyAsync(z, s->vec[5], sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// average device time of y = J x over `reps` launches (vectors and matrix resident; the operands
// of a 3-D grid exceed the L2, small grids are measured warm)
extern "C" int tfb_spmv_bench(tfb_mat* m, int reps, int masked, float* ms_out) {
    TFB_CHECK(m && reps > 0 && ms_out, "bad arguments");
    tfb_ctx* c = m->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (ensure_buffers(c, 0)) return -1;
    tfb_solver_state* s = c->solver;
    const int dim = c->desc.dim;
    const unsigned velmask = (1u << dim) - 1u;
    TFB_CUDA(cudaMemcpyAsync(s->vec[0], s->d_mass, sizeof(double) * c->n_local, cudaMemcpyDeviceToDevice, c->stream));   // any non-zero x
    for (int w = 0; w < 3; w++)
        if (spmv(c, m, s->vec[0], s->vec[1], dim, masked ? velmask : 0u, masked ? velmask : 0u, nullptr)) return -1;
    // One event pair per product, read after the loop; the result is the MEDIAN.  On z-slabs every product meets both
    // neighbours in its halo exchange, so a single late rank (host launch skew at the start of the loop, a descheduled
    // host thread) lengthens individual samples by milliseconds: the mean of such a loop measured the skew (round 1:
    // 2.5 ms at 8 ranks against 0.31 ms at 4), a host synchronisation after every product made it worse (3.3 - 5.5 ms).
    std::vector<cudaEvent_t> ev(2 * (size_t)reps);
    for (auto& e : ev) TFB_CUDA(cudaEventCreate(&e));
    for (int r = 0; r < reps; r++) {
        TFB_CUDA(cudaEventRecord(ev[2 * r], c->stream));
        if (spmv(c, m, s->vec[0], s->vec[1], dim, masked ? velmask : 0u, masked ? velmask : 0u, nullptr)) return -1;
        TFB_CUDA(cudaEventRecord(ev[2 * r + 1], c->stream));
    }
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<float> t(reps, 0.f);
    for (int r = 0; r < reps; r++) cudaEventElapsedTime(&t[r], ev[2 * r], ev[2 * r + 1]);
    for (auto& e : ev) cudaEventDestroy(e);
    std::sort(t.begin(), t.end());
    *ms_out = t[reps / 2];
    return 0;
}
