// Direct solve for the 2-D configurations (lid-driven / heated cavity, double gyre, AMOC): the counterpart of the
// SciPy backend's SuperLU path (interface/SciPy.py:131-162,204-258) where the Krylov solver loses to it.
//
// Rows are ordered i-fastest, so with one grid line (all cells of a j, all unknowns: m = dof * nx) as a block the pinned
// Jacobian is BLOCK TRIDIAGONAL: line j couples to lines j-1, j, j+1 only.  Block elimination over the lines,
//     S_0 = B_0,   S_j = B_j - L_j S_{j-1}^-1 U_{j-1},
// keeps one dense m x m inverse per line (ny * m^2 doubles: 1.7 GB for AMOC 256 x 128, 20 MB for the 64 x 64 heated
// cavity); L_j, U_j stay sparse (they are read from the CSR matrix).  A solve is then two dense matrix-vector products
// per line.  S_j^-1 is formed by in-place Gauss-Jordan elimination with partial pivoting INSIDE the block, which is what
// makes the saddle-point structure (zero pressure diagonal) harmless; the elimination runs as one persistent kernel,
// all CTAs co-resident, two grid barriers per column.  Everything is fp64; the factors are cached on the tfb_mat like
// `jac.lu` on the reference's matrices, so the second solve of a corrector step only pays the substitution.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "tfb_internal.h"

struct tfb_direct_factor {
    int m = 0, nl = 0, prow = -2;
    uint64_t version = ~0ull;
    double* Sinv = nullptr;     // nl x m x m
    double* S = nullptr;        // m x m work (Gauss-Jordan runs here)
    double* W = nullptr;        // m x m work: L_j S_{j-1}^-1
    double* colbuf = nullptr;   // m
    double* rowbuf = nullptr;   // m
    int* piv = nullptr;         // m
    int* colsrc = nullptr;      // m
    int* swp = nullptr;         // 2 x 160: rows touched by the swaps of a panel (k_gj_blocked / k_gj_lookahead)
    unsigned* bar = nullptr;    // grid barrier counter + status word
    double* y = nullptr;        // n work vectors
    double* z = nullptr;
    double* rr = nullptr;
    double* vb = nullptr;       // n: right-hand side, solution and correction of a solve
    double* vx = nullptr;
    double* vdx = nullptr;
    float factor_ms = 0.f;
};

static void direct_release(tfb_direct_factor* f) {
    cudaFree(f->Sinv); cudaFree(f->S); cudaFree(f->W); cudaFree(f->colbuf); cudaFree(f->rowbuf); cudaFree(f->piv);
    cudaFree(f->colsrc); cudaFree(f->swp); cudaFree(f->bar); cudaFree(f->y); cudaFree(f->z); cudaFree(f->rr);
    cudaFree(f->vb); cudaFree(f->vx); cudaFree(f->vdx);
    delete f;
}

// A Newton loop factors one Jacobian per step: the work space (the ny line inverses above all) is parked on the context
// when its matrix goes away and taken over by the next one instead of being freed and allocated again.
void tfb_direct_free(tfb_mat* mat) {
    tfb_direct_factor* f = (tfb_direct_factor*)mat->direct;
    if (!f) return;
    mat->direct = nullptr;
    tfb_ctx* c = mat->ctx;
    if (c && !c->direct_pool) {
        cudaStreamSynchronize(c->stream);     // nothing of this matrix is in flight any more
        f->version = ~0ull;
        c->direct_pool = f;
        return;
    }
    direct_release(f);
}

void tfb_direct_pool_free(tfb_ctx* c) {
    if (c->direct_pool) direct_release((tfb_direct_factor*)c->direct_pool);
    c->direct_pool = nullptr;
}

// ---- blocks of line j out of the CSR matrix (pinned: row prow is -1 on the diagonal, column prow dropped) ----
// dense[(r - j m) * m + (c - (j + which) m)] = A(r, c) for the entries of the rows of line j whose column lies in line j + which
__global__ void k_dense_block(int m, int line, int which, const int* __restrict__ row_ptr, const int* __restrict__ col,
                              const double* __restrict__ vals, int prow, double* __restrict__ dense) {
    const int rl = blockIdx.x * blockDim.x + threadIdx.x;
    if (rl >= m) return;
    const long long r = (long long)line * m + rl;
    const long long c0 = (long long)(line + which) * m;
    if (r == prow) { if (which == 0) dense[(long long)rl * m + rl] = -1.0; return; }
    for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
        const long long c = col[e];
        if (c == prow || c < c0 || c >= c0 + m) continue;
        dense[(long long)rl * m + (c - c0)] = vals[e];
    }
}
// W[r][q] = sum_a L_j(r, a) Sinv_prev[a][q]   (rows of line j, columns a in line j - 1)
__global__ void k_sparse_times_dense(int m, int line, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                     const double* __restrict__ vals, int prow, const double* __restrict__ Sprev, double* __restrict__ W) {
    const int rl = blockIdx.y;
    const long long r = (long long)line * m + rl;
    const long long c0 = (long long)(line - 1) * m;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m; q += gridDim.x * blockDim.x) {
        double acc = 0.0;
        if (r != prow)
            for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
                const long long c = col[e];
                if (c == prow || c < c0 || c >= c0 + m) continue;
                acc += vals[e] * Sprev[(c - c0) * m + q];
            }
        W[(long long)rl * m + q] = acc;
    }
}
// S[r][c] -= sum_b W[r][b] U_{j-1}(b, c): one thread per (r, b), atomics on S (several b reach the same c)
__global__ void k_dense_times_sparse_sub(int m, int line, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                         const double* __restrict__ vals, int prow, const double* __restrict__ W, double* __restrict__ S) {
    const int bl = blockIdx.x * blockDim.x + threadIdx.x;      // row b of line j-1
    const int rl = blockIdx.y;
    if (bl >= m) return;
    const long long b = (long long)(line - 1) * m + bl;
    if (b == prow) return;
    const double w = W[(long long)rl * m + bl];
    if (w == 0.0) return;
    const long long c0 = (long long)line * m;
    for (int e = row_ptr[b]; e < row_ptr[b + 1]; e++) {
        const long long c = col[e];
        if (c == prow || c < c0 || c >= c0 + m) continue;
        atomicAdd(&S[(long long)rl * m + (c - c0)], -w * vals[e]);
    }
}

// ---- in-place Gauss-Jordan inversion with partial pivoting, one column at a time, persistent multi-CTA kernel ----
// (the fallback for blocks too large for the blocked kernel below, and TFB_DIRECT_UNBLOCKED=1)
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks, unsigned& epoch) {
    __syncthreads();
    if (nblocks == 1) return;
    if (threadIdx.x == 0) {
        epoch++;
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned target = epoch * nblocks;
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
        __threadfence();      // gpu-scope fence: the CTA's L1 must not serve lines written by other CTAs before the barrier
    }
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT) k_gauss_jordan(int m, double* __restrict__ A, double* __restrict__ colbuf, double* __restrict__ rowbuf,
                                                      int* __restrict__ piv, int* __restrict__ colsrc, unsigned* __restrict__ bar,
                                                      double* __restrict__ out, double tiny) {
    __shared__ double s_val[32];
    __shared__ int s_row[32];
    __shared__ int s_p;
    unsigned epoch = 0;
    const unsigned nb = gridDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    const long long mm = (long long)m * m;
    for (int k = 0; k < m; k++) {
        // pivot row of column k (every CTA scans the column itself: no exchange, deterministic tie-break)
        double best = -1.0;
        int brow = k;
        for (int i = k + threadIdx.x; i < m; i += blockDim.x) {
            const double v = fabs(A[(long long)i * m + k]);
            if (v > best) { best = v; brow = i; }
        }
        // warp-level reduction, then the eight warp results
        for (int o = 16; o > 0; o >>= 1) {
            const double v = __shfl_xor_sync(0xffffffffu, best, o);
            const int rw = __shfl_xor_sync(0xffffffffu, brow, o);
            if (v > best || (v == best && rw < brow)) { best = v; brow = rw; }
        }
        if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_row[threadIdx.x >> 5] = brow; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); w++)
                if (s_val[w] > s_val[0] || (s_val[w] == s_val[0] && s_row[w] < s_row[0])) { s_val[0] = s_val[w]; s_row[0] = s_row[w]; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_p = s_row[0];
            if (blockIdx.x == 0) { piv[k] = s_row[0]; if (!(s_val[0] > tiny)) bar[1] = 1u; }    // singular block
        }
        __syncthreads();
        const int p = s_p;
        // rows k and p swap (column k excepted: it is rewritten from colbuf below); old row p -> rowbuf, column k -> colbuf
        for (long long j = gtid; j < m; j += gsz) {
            const double ap = A[(long long)p * m + j];
            rowbuf[j] = ap;
            if (p != k && j != k) A[(long long)p * m + j] = A[(long long)k * m + j];
            double cv = A[j * m + k];                                   // j doubles as a row index here
            if (j == k) cv = A[(long long)p * m + k];
            else if (j == p) cv = A[(long long)k * m + k];
            colbuf[j] = cv;
        }
        grid_barrier(bar, nb, epoch);
        const double pivinv = 1.0 / colbuf[k];
        // one row per CTA at a time, threads along the row: coalesced, no index divisions
        for (int i = blockIdx.x; i < m; i += gridDim.x) {
            double* __restrict__ arow = A + (long long)i * m;
            if (i == k) {
                for (int j = threadIdx.x; j < m; j += blockDim.x) arow[j] = j == k ? pivinv : rowbuf[j] * pivinv;
            } else {
                const double f = colbuf[i] * pivinv;
                for (int j = threadIdx.x; j < m; j += blockDim.x) arow[j] = j == k ? -f : arow[j] - f * rowbuf[j];
            }
        }
        grid_barrier(bar, nb, epoch);
    }
    // inverse of the row-permuted matrix -> inverse: undo the swaps on the columns, composed into one gather
    if (gtid == 0) {
        for (int c = 0; c < m; c++) colsrc[c] = c;
        for (int k = m - 1; k >= 0; k--) {
            const int p = piv[k];
            if (p != k) { const int t = colsrc[k]; colsrc[k] = colsrc[p]; colsrc[p] = t; }
        }
    }
    grid_barrier(bar, nb, epoch);
    for (int i = blockIdx.x; i < m; i += gridDim.x)
        for (int j = threadIdx.x; j < m; j += blockDim.x) out[(long long)i * m + j] = A[(long long)i * m + colsrc[j]];
    (void)mm;
}

// ---- blocked in-place Gauss-Jordan inversion (default) ----
// The column-at-a-time kernel above pays two grid barriers per column: 2 m barriers per line block, 330 000 for the AMOC
// configuration (m = 1280, 128 lines), which is where its 2.1 s went.  Here NB columns at a time form a PANEL that CTA 0
// eliminates on its own in shared memory (all m rows of those columns; pivot search among the rows not used yet, barriers
// are __syncthreads), and the rest of the matrix receives one rank-NB update per panel from all CTAs:
//     A_J  <-  [rows outside the panel's pivot rows] A_J  +  P_new * A_K,J     (after the panel's row swaps; J: all other columns)
// with P_new the eliminated panel (A_KK^-1 in the pivot rows, -A_OK A_KK^-1 elsewhere) -- two grid barriers per PANEL.
// A block that fits in shared memory whole (m <= ~150: the 32 x 32 cavity) is a single panel and a single CTA.
// tools/proto/blocked_gauss_jordan.py is the numpy model of the algebra.
//   swp: [0] = number of touched rows nt, [1..64] = touched rows T (the pivot rows first), [65..128] = sigma: after the
//   swaps row T[a] holds what row T[sigma[a]] held before.
template <int NT>
__global__ void __launch_bounds__(NT) k_gj_blocked(int m, int NB, double* __restrict__ A, int* __restrict__ piv, int* __restrict__ swp,
                                                   int* __restrict__ colsrc, unsigned* __restrict__ bar, double* __restrict__ out, double tiny) {
    extern __shared__ double sm[];
    const int ld = NB | 1;                       // odd row pitch: column accesses are conflict-free
    double* P = sm;                              // m x ld panel (CTA 0); afterwards tmp / AK of the update
    const size_t psz = (size_t)m * ld + 2048;    // + tmp [64][32] of the update (which keeps the panel behind it)
    double* colbuf = sm + psz;                   // m
    double* rowbuf = colbuf + m;                 // NB
    double* rowold = rowbuf + NB;                // NB
    int* posof = reinterpret_cast<int*>(rowold + NB);   // m: index of a row in T, -1 if untouched
    int* spiv = posof + m;                       // NB
    __shared__ double s_val[NT / 32];
    __shared__ int s_row[NT / 32];
    __shared__ int s_T[64], s_sig[64], s_nt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    unsigned epoch = 0;
    const unsigned nblk = gridDim.x;
    for (int k0 = 0; k0 < m; k0 += NB) {
        const int nb = min(NB, m - k0);
        const bool whole = nb == m;              // no other columns: nothing to update
        if (blockIdx.x == 0) {
            // ---- panel into shared memory ----
            // thread <-> (column, row group): JW lanes along a row (the power of two that covers nb, at most 32), rows
            // strided by NT / JW; no index divisions, eight independent loads in flight per thread
            const int JW = nb <= 8 ? 8 : nb <= 16 ? 16 : 32;
            const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
            for (int j = jl; j < nb; j += JW)
                for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; v[u] = i < m ? A[(size_t)i * m + k0 + j] : 0.0; }
#pragma unroll
                    for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; if (i < m) P[i * ld + j] = v[u]; }
                }
            if (!whole) for (int i = tid; i < m; i += NT) posof[i] = (i >= k0 && i < k0 + nb) ? i - k0 : -1;
            __syncthreads();
            for (int s = 0; s < nb; s++) {
                const int k = k0 + s;
                // pivot of column s among the rows k..m-1 (smallest row on ties)
                double best = -1.0;
                int brow = k;
                for (int i = k + tid; i < m; i += NT) {
                    const double v = fabs(P[i * ld + s]);
                    if (v > best) { best = v; brow = i; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double v = __shfl_xor_sync(0xffffffffu, best, o);
                    const int rw = __shfl_xor_sync(0xffffffffu, brow, o);
                    if (v > best || (v == best && rw < brow)) { best = v; brow = rw; }
                }
                if (lane == 0) { s_val[warp] = best; s_row[warp] = brow; }
                __syncthreads();
                // every warp reduces the NW warp results again with shuffles (no second barrier, no serial loop)
                best = lane < NW ? s_val[lane] : -1.0;
                brow = lane < NW ? s_row[lane] : m;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double v = __shfl_xor_sync(0xffffffffu, best, o);
                    const int rw = __shfl_xor_sync(0xffffffffu, brow, o);
                    if (v > best || (v == best && rw < brow)) { best = v; brow = rw; }
                }
                const int p = brow;
                if (tid == 0) {
                    piv[k] = p;
                    spiv[s] = p;
                    if (!(best > tiny)) bar[1] = 1u;     // numerically singular block
                }
                // The row that becomes the pivot row (rowbuf), the row it displaces (rowold), and the multiplier of every
                // row: column s as it stands after the swap, times 1 / pivot.  The pivot row itself carries -1 / pivot and
                // starts from zero, so that one formula,  new = base - f * rowbuf  (and -f in column s), serves all rows.
                const double pivinv = 1.0 / P[p * ld + s];
                for (int j = tid; j < nb; j += NT) { rowbuf[j] = P[p * ld + j]; rowold[j] = P[k * ld + j]; }
                for (int i = tid; i < m; i += NT) {
                    double cv = P[i * ld + s];
                    if (i == p) cv = P[k * ld + s];
                    colbuf[i] = i == k ? -pivinv : cv * pivinv;
                }
                __syncthreads();
                const bool swapped = p != k;
                for (int j = jl; j < nb; j += JW) {
                    const double rb = rowbuf[j], ro = rowold[j];
                    const bool js = j == s;
                    double* col = P + j;
                    // every row by the same formula (the values it leaves in rows k and p are wrong) ...
#pragma unroll 4
                    for (int i = rg; i < m; i += nrg) {
                        const double f = colbuf[i];
                        const double v = col[i * ld] - f * rb;
                        col[i * ld] = js ? -f : v;
                    }
                    // ... then the owner of rows k and p in this column (the thread that just wrote them) sets them right:
                    // the pivot row starts from zero, the row that received the displaced row k from that row
                    if (rg == k % nrg) { const double f = colbuf[k]; col[k * ld] = js ? -f : 0.0 - f * rb; }
                    if (swapped && rg == p % nrg) { const double f = colbuf[p]; col[p * ld] = js ? -f : ro - f * rb; }
                }
                __syncthreads();
            }
            // ---- panel back to the matrix; which rows the swaps touched and where their contents went ----
            for (int j = jl; j < nb; j += JW)
                for (int i = rg; i < m; i += nrg) A[(size_t)i * m + k0 + j] = P[i * ld + j];
            if (!whole && tid == 0) {
                int nt = nb;
                for (int a = 0; a < nb; a++) { s_T[a] = k0 + a; s_sig[a] = a; }
                for (int s2 = 0; s2 < nb; s2++) {
                    const int p = spiv[s2];
                    int b = posof[p];
                    if (b < 0) { b = nt; posof[p] = nt; s_T[nt] = p; s_sig[nt] = nt; nt++; }
                    const int t = s_sig[s2]; s_sig[s2] = s_sig[b]; s_sig[b] = t;
                }
                swp[0] = nt;
                for (int a = 0; a < nt; a++) { swp[1 + a] = s_T[a]; swp[65 + a] = s_sig[a]; }
            }
        }
        if (whole) break;
        grid_barrier(bar, nblk, epoch);
        // ---- all CTAs: the other columns.  A CTA owns a contiguous range of W columns for the whole elimination (no hazards
        // between CTAs), keeps the eliminated panel in shared memory and streams its columns with eight rows in flight.
        if (tid == 0) s_nt = swp[0];
        if (tid < 64) { s_T[tid] = swp[1 + tid]; s_sig[tid] = swp[65 + tid]; }
        const int W = (m + (int)gridDim.x - 1) / (int)gridDim.x;          // columns per CTA
        const int WP = W <= 8 ? 8 : W <= 16 ? 16 : 32;                      // lanes per row (power of two)
        const int c0 = blockIdx.x * W, c1 = min(m, c0 + W);
        if (c0 < c1) {
            // P_new: the panel columns of every row (the K rows hold A_KK^-1), pitch ld as in CTA 0
            double* Ps = P + 2048;           // behind tmp [64][32]
            const int nb8 = (nb + 7) & ~7;   // columns nb..nb8 are zero: the update below runs in blocks of eight
            {
                const int JW = nb8 <= 8 ? 8 : nb8 <= 16 ? 16 : 32;
                const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
                if (jl < nb8)
                    for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; v[u] = (i < m && jl < nb) ? A[(size_t)i * m + k0 + jl] : 0.0; }
#pragma unroll
                        for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; if (i < m) Ps[i * ld + jl] = v[u]; }
                    }
            }
            __syncthreads();
            const int nt = s_nt;
            double* tmp = sm;              // [64][WP]: rows touched by the swaps, already permuted
            for (int cb = c0; cb < c1; cb += 32) {                         // (W > 32 only on small grids with few CTAs)
                const int jc = tid & (WP - 1), rg = tid / WP, nrg = NT / WP;
                const int j = cb + jc;
                const bool jok = jc < W && j < c1 && !(j >= k0 && j < k0 + nb);
                for (int a = rg; a < nt; a += nrg) tmp[a * WP + jc] = jok ? A[(size_t)s_T[s_sig[a]] * m + j] : 0.0;
                __syncthreads();
                // swapped rows outside the panel's pivot rows go back first (they are ordinary rows of the update below)
                for (int a = nb + rg; a < nt; a += nrg) if (jok) A[(size_t)s_T[a] * m + j] = tmp[a * WP + jc];
                __syncthreads();
                if (jok) {
                    double ak[32];
#pragma unroll
                    for (int s2 = 0; s2 < 32; s2++) ak[s2] = s2 < nb ? tmp[s2 * WP + jc] : 0.0;    // A_K,j after the swaps
                    for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                        double acc[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int i = i0 + u * nrg;
                            acc[u] = (i < m && !(i >= k0 && i < k0 + nb)) ? A[(size_t)i * m + j] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            const int i = i0 + u * nrg;
                            if (i < m) {
                                const double* pr = Ps + (size_t)i * ld;
                                double a2 = acc[u];
#pragma unroll
                                for (int blk = 0; blk < 4; blk++)
                                    if (nb > 8 * blk) {
#pragma unroll
                                        for (int t2 = 0; t2 < 8; t2++) a2 += pr[8 * blk + t2] * ak[8 * blk + t2];
                                    }
                                A[(size_t)i * m + j] = a2;
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        grid_barrier(bar, nblk, epoch);
    }
    // inverse of the row-permuted matrix -> inverse: undo the swaps on the columns, composed into one gather
    grid_barrier(bar, nblk, epoch);
    if (blockIdx.x == 0 && tid == 0) {
        for (int c = 0; c < m; c++) colsrc[c] = c;
        for (int k = m - 1; k >= 0; k--) {
            const int p = piv[k];
            if (p != k) { const int t = colsrc[k]; colsrc[k] = colsrc[p]; colsrc[p] = t; }
        }
    }
    grid_barrier(bar, nblk, epoch);
    for (int i = blockIdx.x; i < m; i += gridDim.x)
        for (int j = tid; j < m; j += NT) out[(size_t)i * m + j] = A[(size_t)i * m + colsrc[j]];
}


// ---- the same elimination with LOOK-AHEAD (opt-in: TFB_DIRECT_LOOKAHEAD=1, see direct_factor) ----
// In k_gj_blocked every CTA but one waits while the panel is eliminated, and that one waits while the others update.  Here
// CTA 0 owns no columns of its own: while the other CTAs apply panel n to their columns, CTA 0 applies it to the columns of
// panel n+1 only (straight into its shared-memory copy), eliminates that panel and publishes it -- one grid barrier per
// panel, and the update hides behind the elimination (which is the longer of the two).  The swap lists are double-buffered.
template <int NT>
__global__ void __launch_bounds__(NT) k_gj_lookahead(int m, int NB, double* __restrict__ A, int* __restrict__ piv, int* __restrict__ swp,
                                                     int* __restrict__ colsrc, unsigned* __restrict__ bar, double* __restrict__ out, double tiny) {
    extern __shared__ double sm[];
    const int ld = NB | 1;
    double* tmp = sm;                            // [64][32]: rows touched by the swaps of the current panel, permuted
    double* P = sm + 2048;                       // CTA 0: the panel being eliminated; others: the eliminated panel
    double* colbuf = sm + (size_t)m * ld + 2048; // m
    double* rowbuf = colbuf + m;                 // NB
    double* rowold = rowbuf + NB;                // NB
    int* posof = reinterpret_cast<int*>(rowold + NB);
    int* spiv = posof + m;
    __shared__ double s_val[NT / 32];
    __shared__ int s_row[NT / 32];
    __shared__ int s_T[64], s_sig[64], s_nt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    unsigned epoch = 0;
    const unsigned nblk = gridDim.x;
    const int np = (m + NB - 1) / NB;

    // columns k0..k0+nb of all rows -> P (thread <-> (column, row group), eight loads in flight)
    auto load_plain = [&](int k0, int nb) {
        const int JW = nb <= 8 ? 8 : nb <= 16 ? 16 : 32;
        const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
        for (int j = jl; j < nb; j += JW)
            for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; v[u] = i < m ? A[(size_t)i * m + k0 + j] : 0.0; }
#pragma unroll
                for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; if (i < m) P[i * ld + j] = v[u]; }
            }
    };
    // Gauss-Jordan steps on the panel in P (see k_gj_blocked), panel back to the matrix, swap list to swp_out
    auto eliminate = [&](int k0, int nb, int* swp_out) {
        const int JW = nb <= 8 ? 8 : nb <= 16 ? 16 : 32;
        const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
        for (int i = tid; i < m; i += NT) posof[i] = (i >= k0 && i < k0 + nb) ? i - k0 : -1;
        __syncthreads();
        for (int s = 0; s < nb; s++) {
            const int k = k0 + s;
            // pivot of column s among the rows k..m-1 (smallest row on ties)
            double best = -1.0;
            int brow = k;
            for (int i = k + tid; i < m; i += NT) {
                const double v = fabs(P[i * ld + s]);
                if (v > best) { best = v; brow = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double v = __shfl_xor_sync(0xffffffffu, best, o);
                const int rw = __shfl_xor_sync(0xffffffffu, brow, o);
                if (v > best || (v == best && rw < brow)) { best = v; brow = rw; }
            }
            if (lane == 0) { s_val[warp] = best; s_row[warp] = brow; }
            __syncthreads();
            // every warp reduces the NW warp results again with shuffles (no second barrier, no serial loop)
            best = lane < NW ? s_val[lane] : -1.0;
            brow = lane < NW ? s_row[lane] : m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double v = __shfl_xor_sync(0xffffffffu, best, o);
                const int rw = __shfl_xor_sync(0xffffffffu, brow, o);
                if (v > best || (v == best && rw < brow)) { best = v; brow = rw; }
            }
            const int p = brow;
            if (tid == 0) {
                piv[k] = p;
                spiv[s] = p;
                if (!(best > tiny)) bar[1] = 1u;     // numerically singular block
            }
            // The row that becomes the pivot row (rowbuf), the row it displaces (rowold), and the multiplier of every
            // row: column s as it stands after the swap, times 1 / pivot.  The pivot row itself carries -1 / pivot and
            // starts from zero, so that one formula,  new = base - f * rowbuf  (and -f in column s), serves all rows.
            const double pivinv = 1.0 / P[p * ld + s];
            for (int j = tid; j < nb; j += NT) { rowbuf[j] = P[p * ld + j]; rowold[j] = P[k * ld + j]; }
            for (int i = tid; i < m; i += NT) {
                double cv = P[i * ld + s];
                if (i == p) cv = P[k * ld + s];
                colbuf[i] = i == k ? -pivinv : cv * pivinv;
            }
            __syncthreads();
            const bool swapped = p != k;
            for (int j = jl; j < nb; j += JW) {
                const double rb = rowbuf[j], ro = rowold[j];
                const bool js = j == s;
                double* col = P + j;
                // every row by the same formula (the values it leaves in rows k and p are wrong) ...
#pragma unroll 4
                for (int i = rg; i < m; i += nrg) {
                    const double f = colbuf[i];
                    const double v = col[i * ld] - f * rb;
                    col[i * ld] = js ? -f : v;
                }
                // ... then the owner of rows k and p in this column (the thread that just wrote them) sets them right:
                // the pivot row starts from zero, the row that received the displaced row k from that row
                if (rg == k % nrg) { const double f = colbuf[k]; col[k * ld] = js ? -f : 0.0 - f * rb; }
                if (swapped && rg == p % nrg) { const double f = colbuf[p]; col[p * ld] = js ? -f : ro - f * rb; }
            }
            __syncthreads();
        }
        for (int j = jl; j < nb; j += JW)
            for (int i = rg; i < m; i += nrg) A[(size_t)i * m + k0 + j] = P[i * ld + j];
        if (tid == 0) {
            int nt = nb;
            for (int a = 0; a < nb; a++) { s_T[a] = k0 + a; s_sig[a] = a; }
            for (int s2 = 0; s2 < nb; s2++) {
                const int p = spiv[s2];
                int b = posof[p];
                if (b < 0) { b = nt; posof[p] = nt; s_T[nt] = p; s_sig[nt] = nt; nt++; }
                const int t = s_sig[s2]; s_sig[s2] = s_sig[b]; s_sig[b] = t;
            }
            swp_out[0] = nt;
            for (int a = 0; a < nt; a++) { swp_out[1 + a] = s_T[a]; swp_out[65 + a] = s_sig[a]; }
        }
    };

    if (blockIdx.x == 0) {
        load_plain(0, min(NB, m));
        __syncthreads();
        eliminate(0, min(NB, m), swp);
    }
    grid_barrier(bar, nblk, epoch);
    const int owners = (int)gridDim.x - 1;
    const int W = (m + owners - 1) / owners;                                // columns per owner CTA
    const int WP = W <= 8 ? 8 : W <= 16 ? 16 : 32;
    for (int n = 0; n < np; n++) {
        const int k0 = n * NB, nb = min(NB, m - k0);
        const int k1 = k0 + NB, nb1 = n + 1 < np ? min(NB, m - k1) : 0;     // the next panel (none after the last)
        const int* sw = swp + 160 * (n & 1);
        __syncthreads();
        if (tid == 0) s_nt = sw[0];
        if (tid < 64) { s_T[tid] = sw[1 + tid]; s_sig[tid] = sw[65 + tid]; }
        __syncthreads();
        const int nt = s_nt;
        if (blockIdx.x == 0) {
            if (nb1 > 0) {
                // ---- panel n applied to the columns of panel n+1, in shared memory; then that panel is eliminated ----
                load_plain(k1, nb1);
                __syncthreads();
                const int JW = nb1 <= 8 ? 8 : nb1 <= 16 ? 16 : 32;
                const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
                if (jl < nb1) for (int a = rg; a < nt; a += nrg) tmp[a * 32 + jl] = P[s_T[s_sig[a]] * ld + jl];
                __syncthreads();
                if (jl < nb1) for (int a = nb + rg; a < nt; a += nrg) P[s_T[a] * ld + jl] = tmp[a * 32 + jl];
                __syncthreads();
                if (jl < nb1) {
                    double ak[32];
#pragma unroll
                    for (int s2 = 0; s2 < 32; s2++) ak[s2] = s2 < nb ? tmp[s2 * 32 + jl] : 0.0;   // rows K_n of these columns
                    for (int i = rg; i < m; i += nrg) {
                        double a2 = (i >= k0 && i < k0 + nb) ? 0.0 : P[i * ld + jl];
                        const double* pr = A + (size_t)i * m + k0;     // P_new(n): panel n is never the last one here, nb = NB
#pragma unroll
                        for (int blk = 0; blk < 4; blk++)
                            if (nb > 8 * blk) {
#pragma unroll
                                for (int t2 = 0; t2 < 8; t2++) a2 += pr[8 * blk + t2] * ak[8 * blk + t2];
                            }
                        P[i * ld + jl] = a2;
                    }
                }
                __syncthreads();
                eliminate(k1, nb1, swp + 160 * ((n + 1) & 1));
            }
        } else {
            // ---- owner CTAs: panel n applied to their columns, the columns of panels n and n+1 excepted ----
            const int c0 = ((int)blockIdx.x - 1) * W, c1 = min(m, c0 + W);
            if (c0 < c1) {
                const int nb8 = (nb + 7) & ~7;
                {
                    const int JW = nb8 <= 8 ? 8 : nb8 <= 16 ? 16 : 32;
                    const int jl = tid & (JW - 1), rg = tid / JW, nrg = NT / JW;
                    if (jl < nb8)
                        for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                            double v[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; v[u] = (i < m && jl < nb) ? A[(size_t)i * m + k0 + jl] : 0.0; }
#pragma unroll
                            for (int u = 0; u < 8; u++) { const int i = i0 + u * nrg; if (i < m) P[i * ld + jl] = v[u]; }
                        }
                }
                __syncthreads();
                for (int cb = c0; cb < c1; cb += 32) {
                    const int jc = tid & (WP - 1), rg = tid / WP, nrg = NT / WP;
                    const int j = cb + jc;
                    const bool jok = jc < W && j < c1 && !(j >= k0 && j < k0 + nb) && !(j >= k1 && j < k1 + nb1);
                    for (int a = rg; a < nt; a += nrg) tmp[a * WP + jc] = jok ? A[(size_t)s_T[s_sig[a]] * m + j] : 0.0;
                    __syncthreads();
                    for (int a = nb + rg; a < nt; a += nrg) if (jok) A[(size_t)s_T[a] * m + j] = tmp[a * WP + jc];
                    __syncthreads();
                    if (jok) {
                        double ak[32];
#pragma unroll
                        for (int s2 = 0; s2 < 32; s2++) ak[s2] = s2 < nb ? tmp[s2 * WP + jc] : 0.0;
                        for (int i0 = rg; i0 < m; i0 += 8 * nrg) {
                            double acc[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                const int i = i0 + u * nrg;
                                acc[u] = (i < m && !(i >= k0 && i < k0 + nb)) ? A[(size_t)i * m + j] : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                const int i = i0 + u * nrg;
                                if (i < m) {
                                    const double* pr = P + (size_t)i * ld;
                                    double a2 = acc[u];
#pragma unroll
                                    for (int blk = 0; blk < 4; blk++)
                                        if (nb > 8 * blk) {
#pragma unroll
                                            for (int t2 = 0; t2 < 8; t2++) a2 += pr[8 * blk + t2] * ak[8 * blk + t2];
                                        }
                                    A[(size_t)i * m + j] = a2;
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        grid_barrier(bar, nblk, epoch);
    }
    if (blockIdx.x == 0 && tid == 0) {
        for (int c = 0; c < m; c++) colsrc[c] = c;
        for (int k = m - 1; k >= 0; k--) {
            const int p = piv[k];
            if (p != k) { const int t = colsrc[k]; colsrc[k] = colsrc[p]; colsrc[p] = t; }
        }
    }
    grid_barrier(bar, nblk, epoch);
    for (int i = blockIdx.x; i < m; i += gridDim.x)
        for (int j = tid; j < m; j += NT) out[(size_t)i * m + j] = A[(size_t)i * m + colsrc[j]];
}

// ---- substitution ----
// z = Sinv_j * y_j  (dense m x m times vector; one warp per row)
__global__ void __launch_bounds__(256) k_dense_matvec(int m, const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= m) return;
    const double* a = A + (long long)row * m;
    double acc = 0.0;
    for (int j = lane; j < m; j += 32) acc += a[j] * x[j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = acc;
}
// t_j = src_j - (block of line j towards line j + which) * v_{j + which}
__global__ void k_line_residual(int m, int line, int which, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                const double* __restrict__ vals, int prow, const double* __restrict__ src, const double* __restrict__ v,
                                double* __restrict__ t) {
    const int rl = blockIdx.x * blockDim.x + threadIdx.x;
    if (rl >= m) return;
    const long long r = (long long)line * m + rl;
    const long long c0 = (long long)(line + which) * m;
    double acc = src[r];
    if (r != prow)
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
            const long long c = col[e];
            if (c == prow || c < c0 || c >= c0 + m) continue;
            acc -= vals[e] * v[c];
        }
    t[r] = acc;
}
__global__ void k_pinned_residual(long long n, const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ vals,
                                  int prow, const double* __restrict__ b, const double* __restrict__ x, double* __restrict__ r) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double acc = b[row];
    if (row == prow) acc += x[row];
    else
        for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++)
            if (col[e] != prow) acc -= vals[e] * x[col[e]];
    r[row] = acc;
}
__global__ void k_vec_add(long long n, const double* __restrict__ a, double* __restrict__ x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += a[i];
}
__global__ void k_norm2(long long n, const double* __restrict__ a, double* __restrict__ out) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += a[i] * a[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

static int direct_factor(tfb_mat* mat, int prow) {
    tfb_ctx* c = mat->ctx;
    const int dof = c->desc.dof, nx = c->desc.nx, ny = c->desc.ny;
    const int m = dof * nx, nl = ny * c->desc.nz;
    tfb_direct_factor* f = (tfb_direct_factor*)mat->direct;
    if (f && f->version == mat->version && f->prow == prow) return 0;
    if (!f && c->direct_pool) {
        tfb_direct_factor* parked = (tfb_direct_factor*)c->direct_pool;
        c->direct_pool = nullptr;
        if (parked->m == m && parked->nl == nl) { f = parked; mat->direct = f; }
        else direct_release(parked);
    }
    if (!f) {
        f = new tfb_direct_factor();
        mat->direct = f;
        f->m = m; f->nl = nl;
        const size_t mm = (size_t)m * m;
        size_t freeb = 0, total = 0;
        TFB_CUDA(cudaMemGetInfo(&freeb, &total));
        TFB_CHECK(sizeof(double) * mm * (nl + 2) < freeb * 0.8, "the line inverses of the direct solve do not fit in device memory");
        TFB_CUDA(cudaMalloc(&f->Sinv, sizeof(double) * mm * nl));
        TFB_CUDA(cudaMalloc(&f->S, sizeof(double) * mm));
        TFB_CUDA(cudaMalloc(&f->W, sizeof(double) * mm));
        TFB_CUDA(cudaMalloc(&f->colbuf, sizeof(double) * m));
        TFB_CUDA(cudaMalloc(&f->rowbuf, sizeof(double) * m));
        TFB_CUDA(cudaMalloc(&f->piv, sizeof(int) * m));
        TFB_CUDA(cudaMalloc(&f->colsrc, sizeof(int) * m));
        TFB_CUDA(cudaMalloc(&f->swp, sizeof(int) * 320));
        TFB_CUDA(cudaMalloc(&f->bar, sizeof(unsigned) * 2));
        TFB_CUDA(cudaMalloc(&f->vb, sizeof(double) * c->n_local));
        TFB_CUDA(cudaMalloc(&f->vx, sizeof(double) * c->n_local));
        TFB_CUDA(cudaMalloc(&f->vdx, sizeof(double) * c->n_local));
        TFB_CUDA(cudaMalloc(&f->y, sizeof(double) * c->n_local));
        TFB_CUDA(cudaMalloc(&f->z, sizeof(double) * c->n_local));
        TFB_CUDA(cudaMalloc(&f->rr, sizeof(double) * (c->n_local + 2)));
    }
    cudaEvent_t e0, e1;
    TFB_CUDA(cudaEventCreate(&e0)); TFB_CUDA(cudaEventCreate(&e1));
    TFB_CUDA(cudaEventRecord(e0, c->stream));
    const size_t mm = (size_t)m * m;
    // the persistent elimination kernels need all their CTAs resident at once
    int per_sm = 0, sms = 0;
    TFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->desc.device));
    // blocked elimination (k_gj_blocked): the panel width is what fits in shared memory next to the column buffers; a block
    // that fits whole is one panel on one CTA.  TFB_DIRECT_UNBLOCKED=1 keeps the column-at-a-time kernel.
    // 512 threads: measured against 256 (AMOC 1122 ms, QG 559 ms) and 1024 with the update's operands in shared memory
    // (977 ms, 484 ms); 512 gives 889 ms, 454 ms -- more warps make the three CTA barriers of a step dearer, fewer leave
    // the LDS -> DFMA -> STS chains exposed
    constexpr int GJ_NT = 512;
    auto gj_bytes = [&](int nb) {
        const size_t pan = (size_t)m * (nb | 1) + 2048;
        return sizeof(double) * (pan + (size_t)m + 2 * (size_t)nb) + sizeof(int) * ((size_t)m + nb + 16);
    };
    const size_t smem_cap = 224 * 1024;
    int NB = 0;
    static int unblocked = -1;
    if (unblocked < 0) { const char* e = getenv("TFB_DIRECT_UNBLOCKED"); unblocked = (e && e[0] == '1') ? 1 : 0; }
    if (!unblocked) {
        if (gj_bytes(m) <= smem_cap) NB = m;
        else for (int nb = 32; nb >= 8; nb -= 8) if (gj_bytes(nb) <= smem_cap) { NB = nb; break; }
    }
    int gjb_grid = 1;
    if (NB > 0) {
        TFB_CUDA(cudaFuncSetAttribute(k_gj_blocked<GJ_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gj_bytes(NB)));
        TFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gj_blocked<GJ_NT>, GJ_NT, gj_bytes(NB)));
        TFB_CHECK(per_sm >= 1, "the elimination kernel does not fit on an SM");
        gjb_grid = NB >= m ? 1 : (int)std::max<long long>(1, std::min<long long>((long long)sms * per_sm, (m + 7) / 8));
    }
    // TFB_DIRECT_LOOKAHEAD=1: the look-ahead kernel.  Correct (same tests), but not the default: measured against
    // k_gj_blocked it gives AMOC 889 -> 847 ms, QG 454 -> 487 ms, heated cavity 46.1 -> 46.5 ms -- the panel elimination on
    // its one CTA is the critical path either way (about 4 us per elimination step at m = 1280), and what the look-ahead
    // hides (the update by the other CTAs, two barriers) it pays back as CTA 0's own update of the next panel's columns.
    static int lookahead = -1;
    if (lookahead < 0) { const char* e = getenv("TFB_DIRECT_LOOKAHEAD"); lookahead = (e && e[0] == '1') ? 1 : 0; }
    int gjl_grid = 0;
    if (NB > 0 && NB < m && lookahead) {
        TFB_CUDA(cudaFuncSetAttribute(k_gj_lookahead<GJ_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gj_bytes(NB)));
        TFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gj_lookahead<GJ_NT>, GJ_NT, gj_bytes(NB)));
        if (per_sm >= 1) gjl_grid = (int)std::min<long long>((long long)sms * per_sm, (m + 7) / 8 + 1);
        if (gjl_grid < 2) gjl_grid = 0;
    }
    TFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gauss_jordan<256>, 256, 0));
    const int gj_grid = (int)std::max<long long>(1, std::min<long long>((long long)sms * std::min(per_sm, 2), ((long long)mm + 256 * 8 - 1) / (256 * 8)));
    double amax = 0.0;   // scale for the singularity test: largest |value| of the matrix
    {
        std::vector<double> probe(std::min<size_t>((size_t)c->nnz, 4096));
        TFB_CUDA(cudaMemcpyAsync(probe.data(), mat->d_vals, sizeof(double) * probe.size(), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        for (double v : probe) amax = std::max(amax, fabs(v));
    }
    const double tiny = 1e-14 * std::max(amax, 1e-300);
    const unsigned rb = (unsigned)((m + 127) / 128);
    for (int j = 0; j < nl; j++) {
        TFB_CUDA(cudaMemsetAsync(f->S, 0, sizeof(double) * mm, c->stream));
        k_dense_block<<<rb, 128, 0, c->stream>>>(m, j, 0, c->d_row_ptr, c->d_col, mat->d_vals, prow, f->S);
        TFB_LAUNCHED();
        if (j > 0) {
            const double* Sprev = f->Sinv + (size_t)(j - 1) * mm;
            k_sparse_times_dense<<<dim3((unsigned)std::min(8, (m + 255) / 256), m), 256, 0, c->stream>>>(m, j, c->d_row_ptr, c->d_col, mat->d_vals, prow, Sprev, f->W);
            k_dense_times_sparse_sub<<<dim3(rb, m), 128, 0, c->stream>>>(m, j, c->d_row_ptr, c->d_col, mat->d_vals, prow, f->W, f->S);
            TFB_LAUNCHED(); TFB_LAUNCHED();
        }
        TFB_CUDA(cudaMemsetAsync(f->bar, 0, sizeof(unsigned) * 2, c->stream));
        if (gjl_grid > 0) k_gj_lookahead<GJ_NT><<<gjl_grid, GJ_NT, gj_bytes(NB), c->stream>>>(m, NB, f->S, f->piv, f->swp, f->colsrc, f->bar, f->Sinv + (size_t)j * mm, tiny);
        else if (NB > 0) k_gj_blocked<GJ_NT><<<gjb_grid, GJ_NT, gj_bytes(NB), c->stream>>>(m, NB, f->S, f->piv, f->swp, f->colsrc, f->bar, f->Sinv + (size_t)j * mm, tiny);
        else k_gauss_jordan<256><<<gj_grid, 256, 0, c->stream>>>(m, f->S, f->colbuf, f->rowbuf, f->piv, f->colsrc, f->bar, f->Sinv + (size_t)j * mm, tiny);
        TFB_LAUNCHED();
        TFB_CUDA(cudaGetLastError());
    }
    unsigned status[2] = {0, 0};
    TFB_CUDA(cudaMemcpyAsync(status, f->bar, sizeof(unsigned) * 2, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaEventRecord(e1, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&f->factor_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    TFB_CHECK(status[1] == 0, "direct solve: a line block is numerically singular");
    f->version = mat->version;
    f->prow = prow;
    return 0;
}

// x = (pinned matrix)^-1 b on the device, factors of `mat` in place
static int direct_substitute(tfb_mat* mat, int prow, const double* d_b, double* d_x) {
    tfb_ctx* c = mat->ctx;
    tfb_direct_factor* f = (tfb_direct_factor*)mat->direct;
    const int m = f->m, nl = f->nl;
    const size_t mm = (size_t)m * m;
    const unsigned rb = (unsigned)((m + 127) / 128), wb = (unsigned)((m + 7) / 8);
    // forward: y_j = b_j - L_j z_{j-1},  z_j = Sinv_j y_j
    for (int j = 0; j < nl; j++) {
        const double* src = d_b;
        if (j > 0) {
            k_line_residual<<<rb, 128, 0, c->stream>>>(m, j, -1, c->d_row_ptr, c->d_col, mat->d_vals, prow, d_b, f->z, f->y);
            src = f->y;
        }
        k_dense_matvec<<<wb, 256, 0, c->stream>>>(m, f->Sinv + (size_t)j * mm, src + (size_t)j * m, f->z + (size_t)j * m);
        TFB_LAUNCHED(); TFB_LAUNCHED();
    }
    // backward: x_j = z_j - Sinv_j (U_j x_{j+1})
    TFB_CUDA(cudaMemcpyAsync(d_x + (size_t)(nl - 1) * m, f->z + (size_t)(nl - 1) * m, sizeof(double) * m, cudaMemcpyDeviceToDevice, c->stream));
    for (int j = nl - 2; j >= 0; j--) {
        // y_j := U_j x_{j+1}  (as 0 - (-U x)): reuse k_line_residual with a zero source
        TFB_CUDA(cudaMemsetAsync(f->y + (size_t)j * m, 0, sizeof(double) * m, c->stream));
        k_line_residual<<<rb, 128, 0, c->stream>>>(m, j, +1, c->d_row_ptr, c->d_col, mat->d_vals, prow, f->y, d_x, f->y);   // y_j = -U_j x_{j+1}
        k_dense_matvec<<<wb, 256, 0, c->stream>>>(m, f->Sinv + (size_t)j * mm, f->y + (size_t)j * m, d_x + (size_t)j * m);  // = -Sinv U x
        k_vec_add<<<rb, 128, 0, c->stream>>>(m, f->z + (size_t)j * m, d_x + (size_t)j * m);
        TFB_LAUNCHED(); TFB_LAUNCHED(); TFB_LAUNCHED();
    }
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// Interface.solve through the direct path: host vectors, pressure pinned at `prow`.  One step of iterative refinement
// brings the answer to spsolve grade.  info->setup_ms = time of the factorisation (0 when the cached factors were reused).
extern "C" int tfb_direct_solve(tfb_mat* mat, const double* b, double* x, int prow, tfb_solve_info* info) {
    TFB_CHECK(mat && b && x, "null argument");
    tfb_ctx* c = mat->ctx;
    TFB_CHECK(c->nranks == 1, "the direct solve is single-GPU");
    TFB_CHECK(c->desc.nz == 1, "the direct solve is for 2-D grids (one line of cells per block)");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    const long long n = c->n_local;
    const bool had = mat->direct && ((tfb_direct_factor*)mat->direct)->version == mat->version && ((tfb_direct_factor*)mat->direct)->prow == prow;
    if (direct_factor(mat, prow)) return -1;
    tfb_direct_factor* f = (tfb_direct_factor*)mat->direct;
    cudaEvent_t e0, e1;
    TFB_CUDA(cudaEventCreate(&e0)); TFB_CUDA(cudaEventCreate(&e1));
    TFB_CUDA(cudaEventRecord(e0, c->stream));
    double *d_b = f->vb, *d_x = f->vx, *d_dx = f->vdx;
    TFB_CUDA(cudaMemcpyAsync(d_b, b, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    if (direct_substitute(mat, prow, d_b, d_x)) return -1;
    const unsigned nb = (unsigned)((n + 255) / 256);
    double norms[2] = {0.0, 0.0};
    for (int ref = 0; ref < 2; ref++) {
        k_pinned_residual<<<nb, 256, 0, c->stream>>>(n, c->d_row_ptr, c->d_col, mat->d_vals, prow, d_b, d_x, f->rr);
        TFB_LAUNCHED();
        if (ref == 1) break;
        if (direct_substitute(mat, prow, f->rr, d_dx)) return -1;
        k_vec_add<<<nb, 256, 0, c->stream>>>(n, d_dx, d_x);
        TFB_LAUNCHED();
    }
    TFB_CUDA(cudaMemsetAsync(f->rr + n, 0, sizeof(double) * 2, c->stream));
    k_norm2<<<64, 256, 0, c->stream>>>(n, f->rr, f->rr + n);
    k_norm2<<<64, 256, 0, c->stream>>>(n, d_b, f->rr + n + 1);
    TFB_LAUNCHED(); TFB_LAUNCHED();
    TFB_CUDA(cudaMemcpyAsync(norms, f->rr + n, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaMemcpyAsync(x, d_x, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaEventRecord(e1, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double relres = norms[1] > 0.0 ? sqrt(norms[0] / norms[1]) : sqrt(norms[0]);
    if (info) {
        info->iters = 1; info->relres = relres; info->converged = relres <= 1e-10;
        info->setup_ms = had ? 0.f : f->factor_ms; info->solve_ms = ms;
    }
    return relres <= 1e-8 ? 0 : 1;
}
