// Structured SpMV y = J x for true 3-D grids: the matrix never reads its column indices.
//
// The sparsity pattern is the stencil the assembly kernels wrote: for row (cell, equation) the
// s-th structural slot multiplies the unknown d2 of the cell at offset (dx,dy,dz) -- a compile-time
// table per equation (Cfg::slot) -- and the slots that exist at a cell are given by the generated
// mask function.  So the kernel streams only the CSR VALUES (8 B/nnz instead of 12 B/nnz): each
// line's contiguous span of values is fetched by one TMA bulk load into shared memory
// (prefetched one plane ahead, mbarrier-signalled), x is staged in the same z-marching ring of
// planes as the state in the assembly kernel, and one warp per (equation, line) does the
// multiply-adds.  The pinned pressure row/column of the reference's solve (SciPy.py:95-106) and
// the per-variable row/column masks of the block preconditioner are applied on the fly.
#pragma once
#include "tfb_assemble.cuh"

struct TfbSpmvArgs {
    TfbGrid g;
    const double* x;          // indexed by GLOBAL row: plane k of the grid at x + k * plane
    int kvalid0, kvalid1;     // planes [kvalid0, kvalid1) of x may be read (others are zero)
    const int* row_ptr;       // local rows
    const double* vals;
    double* y;                // local rows
    const double* rowscale;   // optional, local rows
    int k0, nzl;
    int prow_cell_i, prow_cell_j, prow_cell_k, pvar;   // pinned pressure unknown (pvar < 0: none)
    unsigned rowmask, colmask;                          // 0 = all
    int plane_nnz;            // structural non-zeros of a plane away from the z walls (0: unknown, always read row_ptr)
    int kofs0, klim;          // local planes [kofs0, klim) of this launch
};

__device__ __forceinline__ void tfb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tfb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tfb_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tfb_bulk_load(double* smem_dst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// One row: sum over the structural slots.  Four independent accumulators break the fp64
// dependency chain; the pin / mask tests are compiled out (MASKED) or hoisted (pin_near).
template <class Cfg, int D1, int TJ, bool MASKED>
__device__ __forceinline__ double tfb_row_dot(const double* __restrict__ v, unsigned m, bool full, bool pin_near,
                                              const RingState<Cfg, TJ>& P, const TfbSpmvArgs& a, int i, int j, int k) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (full && !pin_near) {
#pragma unroll
        for (int s = 0; s < Cfg::nslot(D1); s++) {
            int d2, dx, dy, dz;
            Cfg::slot(D1, s, d2, dx, dy, dz);
            if (!MASKED || ((a.colmask >> d2) & 1u)) acc[s & 3] = fma(v[s], P(d2, dx, dy, dz), acc[s & 3]);
        }
    } else {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < Cfg::nslot(D1); s++) {
            int d2, dx, dy, dz;
            Cfg::slot(D1, s, d2, dx, dy, dz);
            if (full || ((m >> s) & 1u)) {
                bool take = !MASKED || ((a.colmask >> d2) & 1u);
                if (pin_near && d2 == a.pvar && i + dx == a.prow_cell_i && j + dy == a.prow_cell_j && k + dz == a.prow_cell_k) take = false;
                if (take) acc[s & 3] = fma(v[pos], P(d2, dx, dy, dz), acc[s & 3]);
                pos++;
            }
        }
    }
    return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

template <class Cfg, int TJ, bool MASKED>
__device__ __forceinline__ double tfb_row_dot_any(int d1, const double* __restrict__ v, unsigned m, bool full, bool pin_near,
                                                  const RingState<Cfg, TJ>& P, const TfbSpmvArgs& a, int i, int j, int k) {
    switch (d1) {
    case 0: return tfb_row_dot<Cfg, 0, TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    case 1: return tfb_row_dot<Cfg, 1, TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    case 2: return tfb_row_dot<Cfg, 2, TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    case 3: return tfb_row_dot<Cfg, (Cfg::DOF > 3 ? 3 : 0), TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    case 4: return tfb_row_dot<Cfg, (Cfg::DOF > 4 ? 4 : 0), TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    default: return tfb_row_dot<Cfg, (Cfg::DOF > 5 ? 5 : 0), TJ, MASKED>(v, m, full, pin_near, P, a, i, j, k);
    }
}

// (The scheduler-balancing deal of the assembly kernel, tfb_deal_item, was measured here too: 244 -> 250 us at dof 4,
// 422 -> 419 us at dof 5 -- this kernel waits on memory, not on issue slots; it keeps the plain order.)
template <class Cfg, int TJ, int KCH, bool MASKED>
__global__ void __launch_bounds__(32 * Cfg::DOF * TJ, (Cfg::DOF <= 4 && TJ <= 2 ? 3 : 2))
tfb_spmv_march_kernel(const TfbSpmvArgs a) {
    using M = TfbMarch<Cfg, TJ>;
    constexpr int DOF = Cfg::DOF, W = M::W, H = M::H, DSTR = M::DSTR, SLOT = M::SLOT, LINE_CAP = M::LINE_CAP;
    constexpr int NT = 32 * DOF * TJ;
    constexpr int ROWLEN = W * DOF, NEL = H * ROWLEN, NPT = (NEL + NT - 1) / NT;
    extern __shared__ __align__(16) double smem[];
    double* ring = smem;
    double* sm_val = smem + M::NSLOT * SLOT;
    __shared__ int sm_span[2][TJ][2];
    __shared__ __align__(8) unsigned long long sm_bar[2][TJ];

    const TfbGrid& g = a.g;
    const int il = threadIdx.x, d1 = threadIdx.y, jl = threadIdx.z;
    const int tid = (jl * DOF + d1) * 32 + il;
    const int i0 = blockIdx.x * TFB_TI, j0 = blockIdx.y * TJ;
    const int kbeg = a.kofs0 + blockIdx.z * KCH, kend = min(kbeg + KCH, a.klim);
    const long long plane = (long long)g.nx * g.ny * DOF;

    int l_src[NPT], l_dst[NPT];
    unsigned l_flags = 0u;
#pragma unroll
    for (int t = 0; t < NPT; t++) {
        const int e = tid + t * NT;
        const int yy = e / ROWLEN, cc = e - yy * ROWLEN;
        const int xx = cc / DOF, d = cc - xx * DOF;
        const int jj = j0 + yy - 1, ii = i0 - 1 + xx;
        const bool exists = e < NEL;
        const bool inside = exists && jj >= 0 && jj < g.ny && ii >= 0 && ii < g.nx;
        l_src[t] = inside ? ((jj * g.nx + ii) * DOF + d) : 0;
        l_dst[t] = d * DSTR + yy * W + xx;
        if (inside) l_flags |= 1u << t;
        if (exists) l_flags |= 1u << (16 + t);
    }
    auto fetch = [&](int kglob, double (&v)[NPT]) {
        const bool ok = kglob >= a.kvalid0 && kglob < a.kvalid1;
        const double* pl = a.x + (long long)kglob * plane;
#pragma unroll
        for (int t = 0; t < NPT; t++) v[t] = (ok && ((l_flags >> t) & 1u)) ? pl[l_src[t]] : 0.0;
    };
    auto deposit = [&](int slot, const double (&v)[NPT]) {
        double* dst = ring + slot * SLOT;
#pragma unroll
        for (int t = 0; t < NPT; t++)
            if ((l_flags >> (16 + t)) & 1u) dst[l_dst[t]] = v[t];
    };

    const int i = i0 + il, j = j0 + jl;
    const bool valid = i < g.nx && j < g.ny;
    const bool leader = d1 == 0 && il == 0 && j < g.ny;
    const int ilast = min(i0 + TFB_TI, g.nx);
    long long row = (((long long)kbeg * g.ny + j) * g.nx + i) * DOF + d1;
    long long r0 = (((long long)kbeg * g.ny + j) * g.nx + i0) * DOF;
    const long long rlen = (long long)(ilast - i0) * DOF;
    if (tid < 2 * TJ) {
        tfb_mbar_init(&sm_bar[tid / TJ][tid % TJ], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double pre[NPT];
#pragma unroll
    for (int p = 0; p < 3; p++) {
        fetch(a.k0 + kbeg - 1 + p, pre);
        deposit(p, pre);
    }
    if (kbeg + 1 < kend) fetch(a.k0 + kbeg + 2, pre);
    int rp = valid ? a.row_ptr[row] : 0;
    __syncthreads();   // barriers initialised
    // values of the first plane
    int cur_gb = 0, cur_ge = 0;      // CSR span of the line in the plane loaded last (leader threads)
    auto load_span = [&](int buf, int gb, int ge) {
        cur_gb = gb; cur_ge = ge;
        sm_span[buf][jl][0] = gb;
        sm_span[buf][jl][1] = ge;
        const int ga = gb & ~1;
        const unsigned bytes = (unsigned)(((ge - ga) + 1) & ~1) * 8u;
        tfb_mbar_expect_tx(&sm_bar[buf][jl], bytes);
        if (bytes) tfb_bulk_load(sm_val + (buf * TJ + jl) * LINE_CAP, a.vals + ga, bytes, &sm_bar[buf][jl]);
    };
    if (leader) load_span(0, a.row_ptr[r0], a.row_ptr[r0 + rlen]);

    TfbCell c;
    tfb_cell_flags<0>(g, i, j, a.k0 + kbeg, c);
    const bool xy_interior_lane = !(c.near[0] | c.far[0] | c.far2[0] | c.near[1] | c.far[1] | c.far2[1]) &&
                                  !(Cfg::ID == 7 && i <= 1 && j <= 1);
    // static-slot fast path only for warps without any wall cell (warp-uniform choice, no double pass)
    const bool xy_interior_warp = __all_sync(0xffffffffu, xy_interior_lane || !valid);
    const int kfar2 = tfb_far2_index(g.nz);
    const int cell_off = (jl + 1) * W + (il + 1);
    const bool row_on = !MASKED || ((a.rowmask >> d1) & 1u);
    int s0 = 0;
    __syncthreads();

    int step = 0;
    for (int kl = kbeg; kl < kend; kl++, step++) {
        const int k = a.k0 + kl;
        const bool more = kl + 1 < kend;
        const int s1 = (s0 + 1) & 3, s2 = (s0 + 2) & 3, s3 = (s0 + 3) & 3;
        const int buf = step & 1;
        int rp_next = 0;
        if (more) {
            deposit(s3, pre);
            if (kl + 2 < kend) fetch(k + 3, pre);
            // The row lengths of a plane depend on k only through its z-wall flags: between two planes without flags every
            // CSR offset moves by the same amount, so row_ptr (a dependent global load in front of the bulk load and of
            // the first multiply of the next step) is only read next to the walls.
            const bool flagged = k == 0 || k == g.nz - 1 || k == kfar2 || k + 1 == g.nz - 1 || k + 1 == kfar2;
            const bool shift = a.plane_nnz > 0 && !flagged;
            if (valid) rp_next = shift ? rp + a.plane_nnz : a.row_ptr[row + plane];
            if (leader) {                                   // values of the next plane, one step ahead
                if (shift) load_span(buf ^ 1, cur_gb + a.plane_nnz, cur_ge + a.plane_nnz);
                else load_span(buf ^ 1, a.row_ptr[r0 + plane], a.row_ptr[r0 + plane + rlen]);
            }
        }
        if (j < g.ny) tfb_mbar_wait(&sm_bar[buf][jl], (step >> 1) & 1);
        if (valid) {
            const double* vrow = sm_val + (buf * TJ + jl) * LINE_CAP + (rp - (sm_span[buf][jl][0] & ~1));
            c.k = k;
            c.near[2] = k == 0; c.far[2] = k == g.nz - 1; c.far2[2] = k == kfar2;
            c.cell0 = (i == 0 && j == 0 && k == 0);
            RingState<Cfg, TJ> P;
            P.pl[0] = ring + s0 * SLOT + cell_off;
            P.pl[1] = ring + s1 * SLOT + cell_off;
            P.pl[2] = ring + s2 * SLOT + cell_off;
            const bool z_interior = !(c.near[2] | c.far[2] | c.far2[2]);
            const bool full = xy_interior_warp && z_interior;
            const unsigned m = full ? 0u : ((xy_interior_lane && z_interior) ? ((1u << Cfg::nslot(d1)) - 1u) : Cfg::mask(d1, c));
            double acc = 0.0;
            if (row_on) {
                const bool pin_near = a.pvar >= 0 && abs(i - a.prow_cell_i) <= 1 && abs(j - a.prow_cell_j) <= 1 &&
                                      abs(k - a.prow_cell_k) <= 1;
                acc = tfb_row_dot_any<Cfg, TJ, MASKED>(d1, vrow, m, full, pin_near, P, a, i, j, k);
                if (d1 == a.pvar && i == a.prow_cell_i && j == a.prow_cell_j && k == a.prow_cell_k)
                    acc = MASKED ? 0.0 : -P(d1, 0, 0, 0);     // pinned row: -1 on the diagonal (full operator only)
                if (a.rowscale) acc /= a.rowscale[row];
            }
            a.y[row] = acc;
        }
        __syncthreads();      // ring slot s0 and the value buffer `buf` are free again; deposit visible
        rp = rp_next;
        row += plane;
        r0 += plane;
        s0 = s1;
    }
}
