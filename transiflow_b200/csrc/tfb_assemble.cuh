// CUDA kernels of the assembly path: fixed-pattern discovery, fused Jacobian+RHS assembly and
// the mass diagonal.  fp64, HBM-bound (no tensor cores: nothing here is a dense contraction).
//
// Mapping: one CTA per tile of TI=32 cells (x) by TJ lines (y) in one z-plane; inside the CTA
// one WARP per (equation, line): lane = cell along x, so the generated row function of one
// equation runs divergence-free across the warp.  The state neighbourhood (tile + one-cell
// halo, with the reference's padded-state semantics applied on load) is staged in shared
// memory as structure-of-arrays; each line's CSR values are staged in shared memory and written
// back as one contiguous, 128-bit vectorised span (rows of neighbouring cells are adjacent in
// CSR, Discretization.py:513).
#pragma once
#include "tfb_internal.h"
#include "gen/all_configs.h"

#define TFB_TI 32

struct TfbAsmArgs {
    TfbGrid g;
    TfbParams prm;
    const double* state;      // (nzl+2) planes, ghost plane first
    const double* frc_static; // local rows or null
    const int* row_ptr;       // local rows
    double* vals;
    double* rhs;
    int k0, nzl;
};

template <class Cfg>
struct TfbTile {
    static constexpr int NZP = Cfg::FLAT ? 1 : 3;
    template <int TJ> __host__ __device__ static constexpr int state_doubles() { return Cfg::DOF * NZP * (TJ + 2) * (TFB_TI + 2); }
    // per line: 32 cells * (sum of slots over the rows of a cell) + 2 (alignment slack)
    __host__ __device__ static constexpr int cell_slots() { return Cfg::CELL_SLOTS; }
};

template <class Cfg, int TJ>
struct SmemState {
    const double* base;   // sm_state
    int il, jl;           // lane / line inside the tile
    __device__ __forceinline__ double operator()(int d, int ox, int oy, int oz) const {
        constexpr int NZP = TfbTile<Cfg>::NZP;
        const int zp = NZP == 1 ? 0 : oz + 1;
        return base[((d * NZP + zp) * (TJ + 2) + (jl + 1 + oy)) * (TFB_TI + 2) + (il + 1 + ox)];
    }
};

__device__ __forceinline__ void tfb_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <class Cfg>
__device__ __forceinline__ bool tfb_is_interior(const TfbCell& c) {
    bool b = c.near[0] | c.far[0] | c.far2[0] | c.near[1] | c.far[1] | c.far2[1];
    if (!Cfg::FLAT) b = b | c.near[2] | c.far[2] | c.far2[2];
    if (Cfg::ID == 7) b = b | (c.i <= 1 && c.j <= 1);   // AMOC pinned salinity column
    return !b;
}

template <class Cfg, bool DO_J, bool DO_F, int TJ>
__global__ void __launch_bounds__(32 * Cfg::DOF * TJ)
tfb_assemble_kernel(const TfbAsmArgs a) {
    constexpr int DOF = Cfg::DOF;
    constexpr int NZP = TfbTile<Cfg>::NZP;
    constexpr int NTHREADS = 32 * DOF * TJ;
    constexpr int LINE_CAP = TFB_TI * TfbTile<Cfg>::cell_slots() + 2;
    extern __shared__ __align__(16) double smem[];
    double* sm_state = smem;
    double* sm_out = smem + ((TfbTile<Cfg>::template state_doubles<TJ>() + 1) & ~1);
    __shared__ int sm_span[TJ][2];

    const TfbGrid& g = a.g;
    const int il = threadIdx.x, d1 = threadIdx.y, jl = threadIdx.z;
    const int tid = (jl * DOF + d1) * 32 + il;
    const int i0 = blockIdx.x * TFB_TI, j0 = blockIdx.y * TJ;
    const int kl = blockIdx.z;           // local plane
    const int k = a.k0 + kl;             // global plane
    const int kofs = 1 - a.k0;           // global k -> plane index of the slab storage

    // ---- stage the state tile (+halo) in shared memory, SoA, padded-state semantics ----
    {
        constexpr int W = TFB_TI + 2, H = TJ + 2;
        constexpr int NCELL = NZP * H * W;
        for (int e = tid; e < NCELL * DOF; e += NTHREADS) {
            const int d = e % DOF;           // AoS order in global memory -> coalesced reads
            const int cidx = e / DOF;
            const int xx = cidx % W, yy = (cidx / W) % H, zz = cidx / (W * H);
            const int oz = NZP == 1 ? 0 : zz - 1;
            const double v = tfb_padded_load(g, a.state, kofs, i0 + xx - 1, j0 + yy - 1, k + oz, d);
            sm_state[((d * NZP + zz) * H + yy) * W + xx] = v;
        }
    }
    const int i = i0 + il, j = j0 + jl;
    const bool valid = i < g.nx && j < g.ny;
    const long long cell_local = ((long long)kl * g.ny + j) * g.nx + i;
    const long long row = cell_local * DOF + d1;
    if (DO_J && d1 == 0 && il == 0 && j < g.ny) {
        const int ilast = min(i0 + TFB_TI, g.nx);
        const long long r0 = (((long long)kl * g.ny + j) * g.nx + i0) * DOF;
        const long long r1 = (((long long)kl * g.ny + j) * g.nx + ilast) * DOF;
        sm_span[jl][0] = a.row_ptr[r0];
        sm_span[jl][1] = a.row_ptr[r1];
    }
    __syncthreads();

    double J[Cfg::MAXSLOT];
    double f = 0.0;
    unsigned m = 0u;
    int rp = 0;
    if (valid) {
        TfbCell c;
        tfb_make_cell<Cfg::NFORCE>(g, i, j, k, c);
        SmemState<Cfg, TJ> P{sm_state, il, jl};
        Cfg::template row<DO_J, DO_F>(d1, a.prm, c, P, J, f);
        if (DO_J) {
            const int ns = Cfg::nslot(d1);
            m = tfb_is_interior<Cfg>(c) ? ((1u << ns) - 1u) : Cfg::mask(d1, c);
            rp = a.row_ptr[row];
        }
        if (DO_F) {
            if (a.frc_static) f = f + a.frc_static[row];
            a.rhs[row] = f;
        }
    }
    if (DO_J) {
        // ---- stage this line's CSR values, then write the contiguous span with 128-bit stores ----
        double* out = sm_out + jl * LINE_CAP;
        const int gbase = sm_span[jl][0], gend = sm_span[jl][1];
        const int galign = gbase & ~1;
        if (valid) {
            int pos = rp - galign;
#pragma unroll
            for (int s = 0; s < Cfg::MAXSLOT; s++)
                if ((m >> s) & 1u) out[pos++] = J[s];
        }
        tfb_bar_sync(1 + jl, 32 * DOF);
        if (j < g.ny) {
            const int cnt = gend - galign;
            const int head = gbase - galign;    // 0 or 1
            const int t = d1 * 32 + il;
            double* gout = a.vals + galign;
            for (int e = 2 * t; e < cnt; e += 2 * 32 * DOF) {
                if (e >= head && e + 1 < cnt) {
                    double2 v2 = *reinterpret_cast<const double2*>(out + e);
                    *reinterpret_cast<double2*>(gout + e) = v2;
                } else {
                    if (e >= head) gout[e] = out[e];
                    if (e + 1 < cnt) gout[e + 1] = out[e + 1];
                }
            }
        }
    }
}

// ---- pattern discovery: structural row lengths, then global column indices ----
template <class Cfg>
__global__ void tfb_count_kernel(TfbGrid g, int k0, long long nrows, int* __restrict__ counts) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int d1 = (int)(row % Cfg::DOF);
    const long long cell = row / Cfg::DOF;
    const int i = (int)(cell % g.nx), j = (int)((cell / g.nx) % g.ny), k = k0 + (int)(cell / ((long long)g.nx * g.ny));
    TfbCell c;
    tfb_make_cell<0>(g, i, j, k, c);
    counts[row] = __popc(Cfg::mask(d1, c));
}

template <class Cfg>
__global__ void tfb_fill_cols_kernel(TfbGrid g, int k0, long long nrows, const int* __restrict__ row_ptr,
                                     int* __restrict__ col) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int d1 = (int)(row % Cfg::DOF);
    const long long cell = row / Cfg::DOF;
    const int i = (int)(cell % g.nx), j = (int)((cell / g.nx) % g.ny), k = k0 + (int)(cell / ((long long)g.nx * g.ny));
    TfbCell c;
    tfb_make_cell<0>(g, i, j, k, c);
    const unsigned m = Cfg::mask(d1, c);
    int pos = row_ptr[row];
    const int ns = Cfg::nslot(d1);
    for (int s = 0; s < ns; s++)
        if ((m >> s) & 1u) {
            int d2, dx, dy, dz;
            Cfg::slot(d1, s, d2, dx, dy, dz);
            col[pos++] = (int)tfb_column(g, i, j, k, d2, dx, dy, dz);
        }
}

// Discretization.mass_matrix (Discretization.py:417-437, 1120-1176): control-volume sizes.
__global__ void tfb_mass_kernel(TfbGrid g, int k0, long long nrows, double* __restrict__ diag) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int d = (int)(row % g.dof);
    const long long cell = row / g.dof;
    const int i = (int)(cell % g.nx), j = (int)((cell / g.nx) % g.ny), k = k0 + (int)(cell / ((long long)g.nx * g.ny));
    const double hcx = g.met[0][i], hux = g.met[0][g.nx + i];
    const double hcy = g.met[1][j], huy = g.met[1][g.ny + j];
    const double hcz = g.met[2][k], huz = g.met[2][g.nz + k];
    double v = 0.0;
    if (d == 0) v = (hux * hcy) * hcz;
    else if (d == 1) v = (huy * hcx) * hcz;
    else if (d == 2 && g.dim == 3) v = (huz * hcy) * hcx;
    else if (d > g.dim) v = (hcx * hcy) * hcz;
    diag[row] = v;
}
