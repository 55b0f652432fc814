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
    int kc0;                  // first z-chunk handled by this launch (pipelined host path)
    int kstep;                // planes per z-chunk of this launch (<= KCH; the host pipeline uses half chunks)
    int kofs0, klim;          // the chunks start at local plane kofs0 and stop before klim (z-slabs: interior planes first,
                              // the two planes next to the halo after the exchange has landed)
};

template <class Cfg>
struct TfbTile {
    static constexpr int NZP = Cfg::FLAT ? 1 : 3;
    // doubles per variable of the SoA tile, padded to 4 (mod 16) so that the AoS->SoA transpose
    // stores of a half-warp (4 cells x DOF) hit distinct shared-memory banks
    template <int TJ> __host__ __device__ static constexpr int dstride() {
        return NZP * (TJ + 2) * (TFB_TI + 2) + ((4 - NZP * (TJ + 2) * (TFB_TI + 2) % 16) + 16) % 16;
    }
    template <int TJ> __host__ __device__ static constexpr int state_doubles() { return Cfg::DOF * dstride<TJ>(); }
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
        return base[d * TfbTile<Cfg>::template dstride<TJ>() + (zp * (TJ + 2) + (jl + 1 + oy)) * (TFB_TI + 2) + (il + 1 + ox)];
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

// Interior rows have the full slot set: slot s lands at a compile-time offset of the row start,
// so every value is stored the moment it is computed (short register live ranges).
struct TfbSmemSink {
    double* row;   // smem position of this row's first slot
    __device__ __forceinline__ void put(int s, double x) { row[s] = x; }
};

template <class Cfg, bool DO_J, bool DO_F, int TJ, int MINB>
__global__ void __launch_bounds__(32 * Cfg::DOF * TJ, MINB)
tfb_assemble_kernel(const TfbAsmArgs a) {
    constexpr int DOF = Cfg::DOF;
    constexpr int NZP = TfbTile<Cfg>::NZP;
    constexpr int NTHREADS = 32 * DOF * TJ;
    constexpr int LINE_CAP = TFB_TI * TfbTile<Cfg>::cell_slots() + 2;
    extern __shared__ __align__(16) double smem[];
    double* sm_state = smem;
    double* sm_out = smem + ((TfbTile<Cfg>::template state_doubles<TJ>() + 1) & ~1);
    __shared__ int sm_span[TJ][2];
    __shared__ double sm_mx[TFB_NMET][TFB_TI], sm_my[TFB_NMET + 2][TJ], sm_mz[TFB_NMET];

    const TfbGrid& g = a.g;
    const int il = threadIdx.x, d1 = threadIdx.y, jl = threadIdx.z;
    const int tid = (jl * DOF + d1) * 32 + il;
    const int i0 = blockIdx.x * TFB_TI, j0 = blockIdx.y * TJ;
    const int kl = a.kofs0 + blockIdx.z * a.kstep;   // local plane (whole slab: kofs0 = 0, kstep = 1; the two planes next
                                                     // to the halo of a z-slab: kofs0 = 0, kstep = nzl - 1)
    const int k = a.k0 + kl;             // global plane
    const int kofs = 1 - a.k0;           // global k -> plane index of the slab storage

    // ---- stage the state tile (+halo) in shared memory, SoA, padded-state semantics ----
    // One warp per (z,y) row of the tile: W cells x DOF doubles are contiguous in global memory
    // (AoS), so the reads are fully coalesced; the transpose to SoA happens in the smem store.
    // Padded-state rules (utils.py:62-133): zero outside the domain, wall-normal velocity on the
    // far walls forced to zero, all three z-planes fold onto the single plane when nz == 1.
    {
        constexpr int W = TFB_TI + 2, H = TJ + 2, ROWS = NZP * H, ROWLEN = W * DOF;
        constexpr int NWARPS = NTHREADS / 32;
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < ROWS; r += NWARPS) {
            const int zz = r / H, yy = r - zz * H;
            const int jj = j0 + yy - 1;
            const int kk = (NZP == 1) ? (g.zfold ? 0 : k) : k + zz - 1;
            const bool rowvalid = jj >= 0 && jj < g.ny && kk >= 0 && kk < g.nz;
            const double* src = a.state + (((long long)(kk + kofs) * g.ny + jj) * g.nx + (i0 - 1)) * DOF;
            const bool zero_v = jj == g.ny - 1, zero_w = !g.zfold && kk == g.nz - 1;
            double* dst = sm_state + (zz * H + yy) * W;
#pragma unroll
            for (int e = lane; e < ROWLEN; e += 32) {
                const int xx = e / DOF, d = e - xx * DOF;
                const int ii = i0 - 1 + xx;
                double v = 0.0;
                if (rowvalid && ii >= 0 && ii < g.nx) {
                    v = src[e];
                    if ((d == 0 && ii == g.nx - 1) || (d == 1 && zero_v) || (d == 2 && zero_w)) v = 0.0;
                }
                dst[d * TfbTile<Cfg>::template dstride<TJ>() + xx] = v;
            }
        }
    }
    // grid metrics of the tile: 1-D arrays, staged once per CTA
    for (int e = tid; e < TFB_NMET * TFB_TI; e += NTHREADS) {
        const int mm = e / TFB_TI, xx = e % TFB_TI;
        sm_mx[mm][xx] = (i0 + xx < g.nx) ? g.met[0][mm * g.nx + i0 + xx] : 0.0;
    }
    for (int e = tid; e < (TFB_NMET + 2) * TJ; e += NTHREADS) {
        const int mm = e / TJ, yy = e % TJ;
        double v = 0.0;
        if (j0 + yy < g.ny) v = mm < TFB_NMET ? g.met[1][mm * g.ny + j0 + yy] : g.cor[(mm - TFB_NMET) * g.ny + j0 + yy];
        sm_my[mm][yy] = v;
    }
    if (tid < TFB_NMET) sm_mz[tid] = g.met[2][tid * g.nz + k];
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

    // ---- per-row work: values go to this line's staging area in shared memory ----
    double* out = sm_out + jl * LINE_CAP;
    const int gbase = DO_J ? sm_span[jl][0] : 0, gend = DO_J ? sm_span[jl][1] : 0;
    const int galign = gbase & ~1;
    if (valid) {
        TfbCell c;
        c.hcx = sm_mx[0][il]; c.hux = sm_mx[1][il]; c.rhcx = sm_mx[2][il]; c.rhpx = sm_mx[3][il];
        c.rhmx = sm_mx[4][il]; c.rhux = sm_mx[5][il]; c.wmx = sm_mx[6][il]; c.wpx = sm_mx[7][il];
        c.hcy = sm_my[0][jl]; c.huy = sm_my[1][jl]; c.rhcy = sm_my[2][jl]; c.rhpy = sm_my[3][jl];
        c.rhmy = sm_my[4][jl]; c.rhuy = sm_my[5][jl]; c.wmy = sm_my[6][jl]; c.wpy = sm_my[7][jl];
        c.cor1 = sm_my[8][jl]; c.cor2 = sm_my[9][jl];
        c.hcz = sm_mz[0]; c.huz = sm_mz[1]; c.rhcz = sm_mz[2]; c.rhpz = sm_mz[3];
        c.rhmz = sm_mz[4]; c.rhuz = sm_mz[5]; c.wmz = sm_mz[6]; c.wpz = sm_mz[7];
        tfb_cell_flags<Cfg::NFORCE>(g, i, j, k, c);
        SmemState<Cfg, TJ> P{sm_state, il, jl};
        double f = 0.0;
        const int rp = DO_J ? a.row_ptr[row] : 0;
        if (tfb_is_interior<Cfg>(c)) {
            TfbSmemSink sink{out + (rp - galign)};
            Cfg::template row<DO_J, DO_F, 0>(d1, a.prm, c, P, sink, f);
        } else {
            double J[Cfg::MAXSLOT];
            TfbArraySink sink{J};
            Cfg::template row<DO_J, DO_F, 2>(d1, a.prm, c, P, sink, f);
            if (DO_J) {
                const unsigned m = Cfg::mask(d1, c);
                int pos = rp - galign;
#pragma unroll
                for (int s = 0; s < Cfg::MAXSLOT; s++)
                    if ((m >> s) & 1u) out[pos++] = J[s];
            }
        }
        if (DO_F) {
            if (a.frc_static) f = f + a.frc_static[row];
            a.rhs[row] = f;
        }
    }
    if (DO_J) {
        // ---- write this line's contiguous CSR span: one TMA bulk store (smem -> global) ----
        // writers fence generic-proxy smem stores towards the async proxy, then the line's warps meet
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tfb_bar_sync(1 + jl, 32 * DOF);
        if (j < g.ny && d1 == 0) {
            const int head = gbase - galign;              // 0 or 1: first element not ours if 1
            const int cnt = gend - galign;
            const int body0 = head ? 2 : 0;               // first 16-byte aligned element we own
            const int body1 = cnt & ~1;                   // end of the aligned body
            if (il == 0 && body1 > body0) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(out + body0);
                double* dstp = a.vals + galign + body0;
                const unsigned bytes = (unsigned)(body1 - body0) * 8u;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(dstp), "r"(src), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (il == 1 && head && cnt > 1) a.vals[galign + 1] = out[1];
            if (il == 2 && (cnt & 1) && cnt - 1 >= head) a.vals[galign + cnt - 1] = out[cnt - 1];
            if (il == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

// =====================================================================================
// z-marching variant for true 3-D grids: a CTA owns a (TI x TJ) column of cells and sweeps KCH
// planes.  The state planes live in a 4-slot shared-memory ring (k-1,k,k+1 in use, one slot
// being refilled), so every plane is fetched from HBM/L2 once per column instead of three
// times.  Plane k+3 is requested with cp.async (global -> shared, zero-fill for everything the
// padded-state rules blank) BEFORE the rows of plane k are computed and has to land two steps
// later, so no register holds data in flight; the CSR spans leave through double-buffered TMA
// bulk stores.  One __syncthreads per plane.
// =====================================================================================
template <class Cfg, int TJ>
struct TfbMarch {
    static constexpr int W = TFB_TI + 2, H = TJ + 2, HW = H * W;
    static constexpr int DSTR = HW + ((4 - HW % 16) + 16) % 16;   // = 4 (mod 16): conflict-free transpose stores
    static constexpr int SLOT = Cfg::DOF * DSTR;                  // doubles per ring slot (one plane)
    static constexpr int NSLOT = 5;   // planes k-1, k, k+1 in use, k+2 landed, k+3 in flight
    // Staging of a line's CSR span: lane il (cell il of the line) stores slot s of its row at
    // (cell start) + (row offset) + s, i.e. the lanes of one store instruction are CELL_SLOTS doubles
    // apart.  57 (lid-driven cavity) is odd: 16 distinct 8-byte banks, conflict-free.  72 (Rayleigh-Benard,
    // heated cavity) = 8 mod 16: two banks, 16-way conflicts (ncu: 74 M conflict cycles, short-scoreboard
    // the top stall).  PAIRPAD shifts every PAIR of cells by 2 more doubles (16 bytes, so each pair is still
    // a legal bulk-copy source): 8 distinct banks, and the line leaves as 16 bulk stores instead of one.
    static constexpr bool PAIRPAD = (Cfg::CELL_SLOTS % 8) == 0;
    static constexpr int LINE_CAP = TFB_TI * Cfg::CELL_SLOTS + 2 + (PAIRPAD ? TFB_TI : 0);
    __host__ __device__ static constexpr int smem_doubles(bool do_j) { return NSLOT * SLOT + (do_j ? 2 * TJ * LINE_CAP : 0); }
};

template <class Cfg, int TJ>
struct RingState {
    const double* pl[3];   // planes k-1, k, k+1, already offset to this thread's cell
    __device__ __forceinline__ double operator()(int d, int ox, int oy, int oz) const {
        return pl[oz + 1][d * TfbMarch<Cfg, TJ>::DSTR + oy * TfbMarch<Cfg, TJ>::W + ox];
    }
};

// Which (equation d1, line jl) the warp `w` of a marching CTA works on.  Warp w issues on scheduler w % 4, and with the
// plain (jl * DOF + d1) order every short pressure row of a dof-4 CTA lands on scheduler 3.  item = jl * DOF + d1, one
// nibble per warp.  Two lines, dof 4: schedulers get {u0,u1} {v0,v1} {w0,p0} {w1,p1} in even CTAs and the mirrored deal
// in odd ones; dof 5 (one CTA per SM): {v0,T0,p0} {v1,T1,p1} {w0,u0} {w1,u1}.  One line, dof 5: {p,T} {u} {v} {w};
// dof 4: one equation per scheduler, rotated with the CTA index.  Other shapes keep the plain order.
template <int DOF, int TJ>
__device__ __forceinline__ void tfb_deal_item(int w, int& d1, int& jl) {
    if constexpr (((TJ == 2 || TJ == 1) && (DOF == 4 || DOF == 5)) || (TJ == 3 && DOF == 5)) {
        const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned long long table;
        if (TJ == 3) table = 0xD83CE947B652A10ull;      // {u0,u1,T0,p0} {v0,v1,T1,p1} {u2,v2,T2,p2} {w0,w1,w2}
        else if (TJ == 2) table = DOF == 4 ? (((lin ^ (lin / 148u)) & 1u) ? 0x54731062ull : 0x73546210ull) : 0x8350947261ull;
        else table = DOF == 4 ? (0x3210321032103210ull >> (4 * (lin & 3u))) : 0x42103ull;
        const int item = (int)((table >> (4 * w)) & 15ull);
        jl = item / DOF;
        d1 = item - jl * DOF;
    }
}

// OPT bit 0: warps are dealt to (equation, line) items through a table that spreads the short pressure rows over the
//            four warp schedulers (warp w issues on scheduler w % 4) instead of parking them all on one;
// (Measured and dropped: a third instantiation with only the x-face boundary code (BCM = 1) for the warps on the x walls,
// wall lanes compacted from a scratch row -- 0.238 -> 0.310 ms at 128^3: 40 % more code and spills under the 128-register cap.)
template <class Cfg, bool DO_J, bool DO_F, int TJ, int KCH, int MINB, int EXP = 0, int OPT = 0>
__global__ void __launch_bounds__(32 * Cfg::DOF * TJ, MINB)
tfb_assemble_march_kernel(const TfbAsmArgs a) {
    using M = TfbMarch<Cfg, TJ>;
    constexpr int DOF = Cfg::DOF, W = M::W, H = M::H, DSTR = M::DSTR, SLOT = M::SLOT, LINE_CAP = M::LINE_CAP;
    constexpr int NT = 32 * DOF * TJ;
    constexpr int ROWLEN = W * DOF, NEL = H * ROWLEN, NPT = (NEL + NT - 1) / NT;
    constexpr bool PAIRPAD = M::PAIRPAD;
    extern __shared__ __align__(16) double smem[];
    double* ring = smem;
    double* sm_out = smem + M::NSLOT * SLOT;   // SLOT is even (DSTR multiple of 4)
    __shared__ int sm_span[2][TJ][2];
    __shared__ double sm_mx[TFB_NMET][TFB_TI], sm_my[TFB_NMET + 2][TJ], sm_mz[TFB_NMET][KCH];
    __shared__ int sm_pstart[KCH + 1];   // first CSR offset of every plane of the chunk

    const TfbGrid& g = a.g;
    const int il = threadIdx.x;
    int d1 = threadIdx.y, jl = threadIdx.z;
    const int tid = (jl * DOF + d1) * 32 + il;
    if constexpr (OPT & 1) tfb_deal_item<DOF, TJ>(tid >> 5, d1, jl);
    const int i0 = blockIdx.x * TFB_TI, j0 = blockIdx.y * TJ;
    const int kbeg = a.kofs0 + (blockIdx.z + a.kc0) * a.kstep, kend = min(kbeg + a.kstep, a.klim);   // local planes
    const int kofs = 1 - a.k0;
    const long long plane = (long long)g.nx * g.ny * DOF;

    // ---- k-invariant descriptors of the tile elements this thread moves for every plane ----
    int l_src[NPT], l_dst[NPT];
    unsigned l_flags = 0u;   // bit t: fetch from global; bit 8+t: w-component; bit 16+t: element exists
#pragma unroll
    for (int t = 0; t < NPT; t++) {
        const int e = tid + t * NT;
        const int yy = e / ROWLEN, cc = e - yy * ROWLEN;
        const int xx = cc / DOF, d = cc - xx * DOF;
        const int jj = j0 + yy - 1, ii = i0 - 1 + xx;
        const bool exists = e < NEL;
        const bool inside = exists && jj >= 0 && jj < g.ny && ii >= 0 && ii < g.nx;
        // wall-normal velocities on the far x / y walls read as zero (utils.py:119,125)
        const bool zeroed = (d == 0 && ii == g.nx - 1) || (d == 1 && jj == g.ny - 1);
        l_src[t] = inside ? ((jj * g.nx + ii) * DOF + d) : 0;
        l_dst[t] = d * DSTR + yy * W + xx;
        if (inside && !zeroed) l_flags |= 1u << t;
        if (d == 2) l_flags |= 1u << (8 + t);
        if (exists) l_flags |= 1u << (16 + t);
    }
    // asynchronous copy of one state plane into a ring slot; src-size 0 zero-fills (outside the domain,
    // wall-normal velocities on the far walls)
    auto request = [&](int kglob, int slot) {
        const double* pl = a.state + (long long)(kglob + kofs) * plane;   // ghost planes are zero at the domain ends
        const bool zero_w = !g.zfold && kglob == g.nz - 1;                // utils.py:131
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(ring + slot * SLOT);
#pragma unroll
        for (int t = 0; t < NPT; t++)
            if ((l_flags >> (16 + t)) & 1u) {
                const bool take = ((l_flags >> t) & 1u) && !(zero_w && ((l_flags >> (8 + t)) & 1u));
                const unsigned nbytes = take ? 8u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;"
                             ::"r"(dst0 + 8u * (unsigned)l_dst[t]), "l"(pl + l_src[t]), "r"(nbytes) : "memory");
            }
    };

    // ---- prologue: metrics of the column, first three planes, first CSR spans ----
    for (int e = tid; e < TFB_NMET * TFB_TI; e += NT) {
        const int mm = e / TFB_TI, xx = e % TFB_TI;
        sm_mx[mm][xx] = (i0 + xx < g.nx) ? g.met[0][mm * g.nx + i0 + xx] : 0.0;
    }
    for (int e = tid; e < (TFB_NMET + 2) * TJ; e += NT) {
        const int mm = e / TJ, yy = e % TJ;
        double v = 0.0;
        if (j0 + yy < g.ny) v = mm < TFB_NMET ? g.met[1][mm * g.ny + j0 + yy] : g.cor[(mm - TFB_NMET) * g.ny + j0 + yy];
        sm_my[mm][yy] = v;
    }
    for (int e = tid; e < TFB_NMET * KCH; e += NT) {
        const int mm = e / KCH, zz = e % KCH;
        sm_mz[mm][zz] = (kbeg + zz < a.nzl) ? g.met[2][mm * g.nz + a.k0 + kbeg + zz] : 0.0;
    }
    if (DO_J)
        for (int e = tid; e <= KCH; e += NT) sm_pstart[e] = a.row_ptr[(long long)min(kbeg + e, a.nzl) * plane];
    {
#pragma unroll
        for (int p = 0; p < 3; p++) request(a.k0 + kbeg - 1 + p, p);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (kbeg + 1 < kend) request(a.k0 + kbeg + 2, 3);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    const int i = i0 + il, j = j0 + jl;
    const bool valid = i < g.nx && j < g.ny;
    const bool leader = DO_J && d1 == 0 && il == 0 && j < g.ny;
    const bool storer = PAIRPAD ? (DO_J && d1 == 0 && j < g.ny) : leader;   // threads that issue bulk stores
    const int padl = PAIRPAD ? (il & ~1) : 0;                               // staging shift of this lane's cell
    const int ilast = min(i0 + TFB_TI, g.nx);
    long long row = (((long long)kbeg * g.ny + j) * g.nx + i) * DOF + d1;              // local row of this thread
    long long r0 = (((long long)kbeg * g.ny + j) * g.nx + i0) * DOF;                  // span of this line
    const long long rlen = (long long)(ilast - i0) * DOF;
    int rp = (DO_J && valid) ? a.row_ptr[row] : 0;
    if (leader) {
        sm_span[0][jl][0] = a.row_ptr[r0];
        sm_span[0][jl][1] = a.row_ptr[r0 + rlen];
    }
    // x/y part of the cell context is plane-invariant
    TfbCell c;
    tfb_cell_flags<Cfg::NFORCE>(g, i, j, a.k0 + kbeg, c);
    const bool xy_interior_lane = !(c.near[0] | c.far[0] | c.far2[0] | c.near[1] | c.far[1] | c.far2[1]) &&
                                  !(Cfg::ID == 7 && i <= 1 && j <= 1);
    // The choice between the BC-free and the boundary instantiation is made per WARP: a warp that
    // holds even one wall cell (lanes 0 / 30 / 31 of the x-edge warps) runs the boundary code for all
    // its lanes instead of running both instantiations under divergence.
    const bool xy_interior = (EXP & 1) ? true : __all_sync(0xffffffffu, xy_interior_lane || !valid);
    const int kfar2 = tfb_far2_index(g.nz);
    const int cell_off = (jl + 1) * W + (il + 1);
    int s0 = 0;   // ring slot of plane k-1
    __syncthreads();

    int step = 0;
    for (int kl = kbeg; kl < kend; kl++, step++) {
        const int k = a.k0 + kl;
        const bool more = kl + 1 < kend;
        const int s1 = s0 + 1 >= 5 ? s0 - 4 : s0 + 1, s2 = s0 + 2 >= 5 ? s0 - 3 : s0 + 2, s4 = s0 + 4 >= 5 ? s0 - 1 : s0 + 4;
        // ---- request plane k+3 into the slot plane k-2 has left (everyone passed the last barrier) ----
        int rp_next = 0, span_next0 = 0, span_next1 = 0;
        if (kl + 2 < kend) request(k + 3, s4);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (more) {
            // The row lengths of a plane depend on k only through its wall flags, so two consecutive planes
            // with equal flags have identical layouts: the CSR offsets of plane k+1 are those of plane k
            // shifted by the plane's size.  row_ptr is only read where the flags change (the z walls).
            const bool same_layout = (k == 0) == (k + 1 == 0) && (k == g.nz - 1) == (k + 1 == g.nz - 1) &&
                                     (k == kfar2) == (k + 1 == kfar2);
            if (same_layout) {
                const int shift = sm_pstart[kl - kbeg + 1] - sm_pstart[kl - kbeg];
                rp_next = rp + shift;
                span_next0 = sm_span[step & 1][jl][0] + shift;
                span_next1 = sm_span[step & 1][jl][1] + shift;
            } else {
                if (DO_J && valid) rp_next = a.row_ptr[row + plane];
                if (leader) {
                    span_next0 = a.row_ptr[r0 + plane];
                    span_next1 = a.row_ptr[r0 + plane + rlen];
                }
            }
        }
        // ---- rows of plane k ----
        double* out = sm_out + ((step & 1) * TJ + jl) * LINE_CAP;
        const int gbase = DO_J ? sm_span[step & 1][jl][0] : 0, gend = DO_J ? sm_span[step & 1][jl][1] : 0;
        const int galign = gbase & ~1;
        if (valid) {
            const int zz = kl - kbeg;
            c.hcx = sm_mx[0][il]; c.hux = sm_mx[1][il]; c.rhcx = sm_mx[2][il]; c.rhpx = sm_mx[3][il];
            c.rhmx = sm_mx[4][il]; c.rhux = sm_mx[5][il]; c.wmx = sm_mx[6][il]; c.wpx = sm_mx[7][il];
            c.hcy = sm_my[0][jl]; c.huy = sm_my[1][jl]; c.rhcy = sm_my[2][jl]; c.rhpy = sm_my[3][jl];
            c.rhmy = sm_my[4][jl]; c.rhuy = sm_my[5][jl]; c.wmy = sm_my[6][jl]; c.wpy = sm_my[7][jl];
            c.cor1 = sm_my[8][jl]; c.cor2 = sm_my[9][jl];
            c.hcz = sm_mz[0][zz]; c.huz = sm_mz[1][zz]; c.rhcz = sm_mz[2][zz]; c.rhpz = sm_mz[3][zz];
            c.rhmz = sm_mz[4][zz]; c.rhuz = sm_mz[5][zz]; c.wmz = sm_mz[6][zz]; c.wpz = sm_mz[7][zz];
            c.k = k;
            c.near[2] = k == 0; c.far[2] = k == g.nz - 1; c.far2[2] = k == kfar2;
            c.cell0 = (i == 0 && j == 0 && k == 0);
            RingState<Cfg, TJ> P;
            P.pl[0] = ring + s0 * SLOT + cell_off;
            P.pl[1] = ring + s1 * SLOT + cell_off;
            P.pl[2] = ring + s2 * SLOT + cell_off;
            double f = 0.0;
            if (EXP & 4) {
                f = P.pl[1][d1 * DSTR];
            } else if (xy_interior && ((EXP & 1) || !(c.near[2] | c.far[2] | c.far2[2]))) {
                TfbSmemSink sink{out + (rp - galign) + padl};
                Cfg::template row<DO_J, DO_F, 0>(d1, a.prm, c, P, sink, f);
            } else {
                double J[Cfg::MAXSLOT];
                TfbArraySink sink{J};
                Cfg::template row<DO_J, DO_F, 2>(d1, a.prm, c, P, sink, f);
                if (DO_J) {
                    const unsigned m = Cfg::mask(d1, c);
                    int pos = rp - galign + padl;
#pragma unroll
                    for (int s = 0; s < Cfg::MAXSLOT; s++)
                        if ((m >> s) & 1u) out[pos++] = J[s];
                }
            }
            if (DO_F) {
                if (a.frc_static) f = f + a.frc_static[row];
                a.rhs[row] = f;
            }
        }
        if (leader && more) {
            sm_span[(step + 1) & 1][jl][0] = span_next0;
            sm_span[(step + 1) & 1][jl][1] = span_next1;
        }
        // staged CSR values -> async proxy; deposited plane + spans -> everyone.  The bulk store of
        // the previous plane must have drained its staging buffer before anyone refills it.
        if (DO_J) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // plane k+2 (requested one step ago) has landed
        __syncthreads();
        if (PAIRPAD) {
            if (storer) {
                // one bulk store per pair of cells (even lanes of the line's first warp); rp of these lanes is
                // the CSR offset at which their cell starts
                const int ncl = ilast - i0;
                const int rp2 = __shfl_down_sync(0xffffffffu, rp, 2);
                if (!(il & 1) && il < ncl) {
                    const int rel0 = rp - galign, rel1 = (il + 2 < ncl ? rp2 : gend) - galign;
                    const int head = rel0 & 1, b0 = rel0 + head, b1 = rel1 & ~1;
                    if (b1 > b0 && !(EXP & 2)) {
                        const unsigned src = (unsigned)__cvta_generic_to_shared(out + b0 + il);
                        double* dstp = a.vals + galign + b0;
                        const unsigned bytes = (unsigned)(b1 - b0) * 8u;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(dstp), "r"(src), "r"(bytes) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (head) a.vals[galign + rel0] = out[rel0 + il];
                    if (rel1 & 1) a.vals[galign + rel1 - 1] = out[rel1 - 1 + il];
                }
            }
        } else if (leader) {
            // one TMA bulk store per line (smem -> global), double-buffered staging
            const int head = gbase - galign, cnt = gend - galign;
            const int body0 = head ? 2 : 0, body1 = cnt & ~1;
            if (body1 > body0 && !(EXP & 2)) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(out + body0);
                double* dstp = a.vals + galign + body0;
                const unsigned bytes = (unsigned)(body1 - body0) * 8u;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(dstp), "r"(src), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (head && cnt > 1) a.vals[galign + 1] = out[1];
            if ((cnt & 1) && cnt - 1 >= head) a.vals[galign + cnt - 1] = out[cnt - 1];
        }
        rp = rp_next;
        row += plane;
        r0 += plane;
        s0 = s1;
    }
    if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
