// Plane transforms of the fast-diagonalisation solves on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// A fast-diagonalisation solve applies  out = Qy' * R * Qx'^T  to every xy-plane R of a variable (forward: the
// transposed eigenvector matrices, backward: the matrices themselves).  These are two dense 128-wide contractions
// per plane -- 18 of the 30 cuBLAS fp64 GEMMs of round 1 -- and they feed a *preconditioner*, so they do not need
// fp64: this kernel runs them as a 3xTF32 split (a = a_hi + a_lo with tf32 halves; a*b ~ a_hi*b_hi + a_lo*b_hi +
// a_hi*b_lo, fp32 accumulation in TMEM), which keeps fp32 accuracy (plain TF32 makes IDR diverge, measured in
// round 1) at 3 tensor-core passes instead of the fp64 pipe.
//
//   step 1:  D1[a][j] = sum_i A1[a][i] * R[j][i]          A1 = Qx^T (forward) or Qx (backward), 128 x K1
//   step 2:  D2[b][a] = sum_j A2[b][j] * D1[a][j]         A2 = Qy^T (forward) or Qy (backward), 128 x K2
//
// Both are  D(128 x N) = A(128 x K) * B(N x K)^T  with K-major operands in shared memory (SWIZZLE_128B, 32 tf32 per
// row, or SWIZZLE_64B with 16): A comes pre-split and pre-swizzled from global memory by one bulk copy per K-chunk
// (tfb_tc_format_matrix lays it out), B is produced by the CTA -- for step 1 from the plane in global memory, for
// step 2 from the accumulator D1 in TMEM (lane a holds row a, so the transposition the second contraction needs is
// free: thread a writes row a of the K-major B operand).  Roles: warps 0-3 produce B chunks and run the epilogue
// (thread t <-> TMEM lane t), warp 4 issues the bulk copies of A and the MMAs.  Two CTAs per SM overlap one CTA's
// TMEM<->shared-memory phases with the other's MMAs.
//
// Descriptor formats follow the CUTLASS/CuTe sm_100 definitions (cute/arch/mma_sm100_desc.hpp: UMMA::SmemDescriptor,
// UMMA::InstrDescriptor; canonical K-major layouts in cute/atom/mma_traits_sm100.hpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <vector>

namespace tfbtc {

constexpr int MROWS = 128;             // UMMA M: rows of A / TMEM lanes
constexpr int PRODUCERS = 128;         // warps 0..3
constexpr int THREADS = 160;           // + warp 4 (A loads, MMA issue)
constexpr int MAXQ = 6;                // arrays (variables) per launch

template <int KC>
struct Geo {
    static_assert(KC == 32 || KC == 16, "K-chunk: 32 (SWIZZLE_128B) or 16 (SWIZZLE_64B)");
    static constexpr int ROWB = KC * 4;                 // bytes of one operand row inside a chunk
    static constexpr int SWZ_BITS = KC == 32 ? 3 : 2;
    static constexpr int SBO = 8 * ROWB;                // 8-row swizzle atom
    static constexpr int CHUNK = MROWS * ROWB;          // one 128-row operand chunk (hi or lo)
    static constexpr int STAGE = 4 * CHUNK;             // A_hi, A_lo, B_hi, B_lo
    static constexpr uint64_t LAYOUT = KC == 32 ? 2 : 4;   // UMMA::LayoutType SWIZZLE_128B / SWIZZLE_64B
    // byte offset of 16-byte unit `c` of row `r`: Swizzle<SWZ_BITS,4,3>
    __host__ __device__ static inline uint32_t unit(uint32_t r, uint32_t c) {
        const uint32_t o = r * ROWB + (c << 4);
        return o ^ (((o >> 7) & ((1u << SWZ_BITS) - 1u)) << 4);
    }
};

// tf32 split of an fp32 value: hi keeps the top 10 mantissa bits (what the tensor core reads), lo the rest
__host__ __device__ inline void split_tf32(float a, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
    hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
#else
    union { float f; uint32_t u; } v; v.f = a; v.u &= 0xffffe000u; hi = v.f;
#endif
    lo = a - hi;
}

// ---- host: lay a 128 x kpad matrix out as the kernel's A operand ------------------------------------------
// A[r][k] (row-major, m x kk valid, zero padded to 128 x kpad; kpad multiple of KC).  Output: per K-chunk the
// hi block followed by the lo block, each CHUNK bytes, rows swizzled.  Returns floats written.
template <int KC>
inline size_t tfb_tc_format_matrix(const double* A, int m, int kk, int lda, int kpad, std::vector<float>& out) {
    using G = Geo<KC>;
    const int nch = kpad / KC;
    out.assign((size_t)nch * 2 * (G::CHUNK / 4), 0.f);
    for (int ch = 0; ch < nch; ch++)
        for (int r = 0; r < MROWS; r++)
            for (int e = 0; e < KC; e++) {
                const int k = ch * KC + e;
                const double a = (r < m && k < kk) ? A[(size_t)r * lda + k] : 0.0;
                float hi, lo_unused;
                split_tf32((float)a, hi, lo_unused);
                float lo = (float)(a - (double)hi);
                const size_t base = (size_t)ch * 2 * (G::CHUNK / 4);
                const size_t off = (G::unit(r, e >> 2) >> 2) + (e & 3);
                out[base + off] = hi;
                out[base + G::CHUNK / 4 + off] = lo;
            }
    return out.size();
}

#ifdef __CUDACC__
// ---- PTX wrappers --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
// bounded wait: a wrong descriptor must not hang the GPU box -- after ~2 s of polling the CTA traps
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t a = s32(b);
    uint32_t done = 0;
    long long t0 = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && (spin & 255u) == 255u) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor of a K-major swizzled operand (version 1, base offset 0, LBO unused = 1)
template <int KC>
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    using G = Geo<KC>;
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(G::SBO >> 4) << 32) |
           ((uint64_t)1 << 46) | (G::LAYOUT << 61);
}
// UMMA instruction descriptor: tf32 x tf32 -> f32, both operands K-major, M = 128, N = n
__host__ __device__ inline uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
}

struct PlaneArgs {
    const float* in[MAXQ];       // per array: planes of ld_in * rows floats
    float* out[MAXQ];
    const float* A1[MAXQ];       // formatted matrices (tfb_tc_format_matrix), K1 = pad(n_in_cols), K2 = pad(n_in_rows)
    const float* A2[MAXQ];
    int narr, nplanes;
    int rows, cols;              // extents of a plane (ny, nx); cols contiguous
    long long plane_stride;      // floats between planes
    int k1pad, k2pad;            // padded contraction lengths (multiples of KC): >= cols, >= rows
    int n1, n2;                  // MMA N of step 1 (>= rows, multiple of 16) and of step 2 (>= cols, multiple of 16)
};

// write 16-byte unit c of row r of a K-major chunk: hi and lo halves of 4 values
template <int KC>
__device__ __forceinline__ void store_split(uint8_t* bhi, uint8_t* blo, uint32_t r, uint32_t c, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    const uint32_t o = Geo<KC>::unit(r, c);
    *reinterpret_cast<float4*>(bhi + o) = h;
    *reinterpret_cast<float4*>(blo + o) = l;
}

template <int KC, int STAGES>
__global__ void __launch_bounds__(THREADS, 2) tfb_fdm_plane_kernel(const PlaneArgs a) {
    using G = Geo<KC>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full_a[STAGES], bar_full_b[STAGES], bar_empty[STAGES], bar_d[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t TCOLS = 256;   // D1: columns [0,128), D2: [128,256)

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&bar_full_a[s], 1); mbar_init(&bar_full_b[s], PRODUCERS); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_d[0], 1); mbar_init(&bar_d[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) tmem_alloc(&tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const int nc1 = a.k1pad / KC, nc2 = a.k2pad / KC, per_item = nc1 + nc2;
    const long long items = (long long)a.narr * a.nplanes;
    long long my_items = 0;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) my_items++;
    const long long total = my_items * per_item;     // chunk uses of this CTA, numbered consecutively

    if (warp == 4) {
        if (lane == 0) {
            long long ld = 0;      // next chunk use whose A load has not been issued
            long long use = 0;
            long long n_item = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x, n_item++) {
                const int q = (int)(item / a.nplanes);
                for (int step = 0; step < 2; step++) {
                    const int nch = step ? nc2 : nc1;
                    const uint32_t idesc = instr_desc(step ? a.n2 : a.n1);
                    const uint32_t dcol = tmem + (step ? 128u : 0u);
                    for (int ch = 0; ch < nch; ch++, use++) {
                        // keep the A loads STAGES-2 chunk uses ahead (also across steps and items): loading into the slot of the
                        // use just issued would make this thread wait for its own MMAs
                        while (ld < total && ld < use + STAGES - 1) {
                            const int ls = (int)(ld % STAGES);
                            mbar_wait(&bar_empty[ls], (uint32_t)(((ld / STAGES) & 1) ^ 1));
                            // decode chunk use `ld` -> (item, step, chunk)
                            const long long li = ld / per_item;
                            const int lr = (int)(ld % per_item);
                            const long long litem = blockIdx.x + li * gridDim.x;
                            const int lq = (int)(litem / a.nplanes);
                            const float* src = lr < nc1 ? a.A1[lq] + (size_t)lr * 2 * (G::CHUNK / 4)
                                                        : a.A2[lq] + (size_t)(lr - nc1) * 2 * (G::CHUNK / 4);
                            mbar_expect_tx(&bar_full_a[ls], 2 * G::CHUNK);
                            bulk_g2s(smem + (size_t)ls * G::STAGE, src, 2 * G::CHUNK, &bar_full_a[ls]);
                            ld++;
                        }
                        const int s = (int)(use % STAGES);
                        const uint32_t ph = (uint32_t)((use / STAGES) & 1);
                        mbar_wait(&bar_full_a[s], ph);
                        mbar_wait(&bar_full_b[s], ph);
                        tc_fence_after();
                        const uint32_t sa = s32(smem + (size_t)s * G::STAGE);
#pragma unroll
                        for (int kk = 0; kk < KC / 8; kk++) {
                            const uint64_t ah = smem_desc<KC>(sa + kk * 32), al = smem_desc<KC>(sa + G::CHUNK + kk * 32);
                            const uint64_t bh = smem_desc<KC>(sa + 2 * G::CHUNK + kk * 32), bl = smem_desc<KC>(sa + 3 * G::CHUNK + kk * 32);
                            mma_tf32(dcol, ah, bh, idesc, (ch | kk) ? 1u : 0u);
                            mma_tf32(dcol, al, bh, idesc, 1u);
                            mma_tf32(dcol, ah, bl, idesc, 1u);
                        }
                        mma_commit(&bar_empty[s]);                 // stage reusable when these MMAs have read it
                        if (ch == nch - 1) mma_commit(&bar_d[step]);   // accumulator complete
                    }
                }
                (void)q;
            }
        }
        __syncwarp();
    } else {
        long long use = 0;
        uint32_t dphase = 0;
        const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);
        for (long long item = blockIdx.x; item < items; item += gridDim.x, dphase ^= 1) {
            const int q = (int)(item / a.nplanes);
            const long long plane = item % a.nplanes;
            const float* in = a.in[q] + plane * a.plane_stride;
            float* out = a.out[q] + plane * a.plane_stride;
            // ---- step 1: B chunk = columns [ch*KC, +KC) of the plane, one row per thread ----
            for (int ch = 0; ch < a.k1pad / KC; ch++, use++) {
                const int s = (int)(use % STAGES);
                mbar_wait(&bar_empty[s], (uint32_t)(((use / STAGES) & 1) ^ 1));
                uint8_t* bhi = smem + (size_t)s * G::STAGE + 2 * G::CHUNK;
                uint8_t* blo = bhi + G::CHUNK;
                if (tid < a.n1) {
                    const float* row = in + (long long)tid * a.cols;
                    float4 v[KC / 4];
#pragma unroll
                    for (int c = 0; c < KC / 4; c++) {
                        const int col = ch * KC + 4 * c;
                        if (tid < a.rows && col + 3 < a.cols && ((((uintptr_t)(row + col)) & 15) == 0)) {
                            v[c] = *reinterpret_cast<const float4*>(row + col);
                        } else {
                            v[c].x = (tid < a.rows && col + 0 < a.cols) ? row[col + 0] : 0.f;
                            v[c].y = (tid < a.rows && col + 1 < a.cols) ? row[col + 1] : 0.f;
                            v[c].z = (tid < a.rows && col + 2 < a.cols) ? row[col + 2] : 0.f;
                            v[c].w = (tid < a.rows && col + 3 < a.cols) ? row[col + 3] : 0.f;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < KC / 4; c++) store_split<KC>(bhi, blo, tid, c, v[c]);
                }
                fence_async_smem();
                mbar_arrive(&bar_full_b[s]);
            }
            // ---- step 2: B chunk = columns [ch*KC, +KC) of D1 (TMEM lane = row of B) ----
            mbar_wait(&bar_d[0], dphase);
            tc_fence_after();
            for (int ch = 0; ch < a.k2pad / KC; ch++, use++) {
                const int s = (int)(use % STAGES);
                mbar_wait(&bar_empty[s], (uint32_t)(((use / STAGES) & 1) ^ 1));
                uint8_t* bhi = smem + (size_t)s * G::STAGE + 2 * G::CHUNK;
                uint8_t* blo = bhi + G::CHUNK;
#pragma unroll
                for (int h = 0; h < KC / 16; h++) {
                    uint32_t v[16];
                    if (ch * KC + h * 16 < a.n1) {      // columns beyond the MMA's N were never written
                        tmem_ld16(tmem + lane_addr + (uint32_t)(ch * KC + h * 16), v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e++) v[e] = 0u;
                    }
                    if (tid < a.n2) {
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            store_split<KC>(bhi, blo, tid, h * 4 + c,
                                            make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                        __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])));
                    }
                }
                tc_fence_before();
                fence_async_smem();
                mbar_arrive(&bar_full_b[s]);
            }
            // ---- epilogue: D2[b][a] -> out[b][a] ----
            mbar_wait(&bar_d[1], dphase);
            tc_fence_after();
            for (int c0 = 0; c0 < a.n2; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_addr + 128u + (uint32_t)c0, v);
                tmem_ld_wait();
                if (tid < a.rows) {
                    float* orow = out + (long long)tid * a.cols + c0;
                    if (c0 + 15 < a.cols && ((((uintptr_t)orow) & 15) == 0)) {
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            reinterpret_cast<float4*>(orow)[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                                             __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e++)
                            if (c0 + e < a.cols) orow[e] = __uint_as_float(v[e]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_free(tmem, TCOLS);
}

template <int KC, int STAGES>
inline size_t plane_kernel_smem() { return (size_t)STAGES * Geo<KC>::STAGE + 1024; }

// ---------------------------------------------------------------------------------------------------------
// z direction: after the x/y transforms every horizontal mode (a, b) is left with ONE tridiagonal system
//     [coef * ((lx_a + ly_b) Mz + Kz) + shift * Mz] u = r          (Kz tridiagonal, Mz diagonal, both 1-D)
// so the z direction is a Thomas sweep per mode -- O(nz) work instead of the two dense nz x nz transforms of the
// full diagonalisation, and the form a z-slab partition can solve with interface unknowns only.  The pivots depend
// on (mode, k) but not on the right-hand side: tfb_thomas_setup_kernel stores 1/pivot and the eliminated upper
// diagonal once per parameter set, the solve is then two fused multiply-adds per unknown, in place.
// Arrays are SoA planes: element (k, mode) at k * modes + mode.  A vanishing pivot (the constant mode of an
// all-Neumann operator) gets 1/pivot = 0, i.e. that unknown is pinned to zero.
// ---------------------------------------------------------------------------------------------------------
struct ThomasVar {
    const double* lx;      // eigenvalues along x (ex entries) and y (indexed by global j)
    const double* ly;
    const double* zk;      // 4 x nz table: lower, diagonal, upper of Kz, then Mz
    float* inv;            // [mz][modes]
    float* cp;             // [mz][modes]
    double coef, shift;
    int k0;                // first global plane of the block that is factored (z-slabs: the local block only)
    int mz;                // unknowns of the block along z
    int mx, my;            // active modes along x, y (others are written as zero)
    int pin_last;          // 1: a mode with lx + ly = 0 (constant mode of an all-Neumann operator) gets its last unknown pinned
    double mu_eps;
};

__global__ void tfb_thomas_setup_kernel(ThomasVar v, int ex, int ey, int jofs, int nz, double thresh) {
    const long long modes = (long long)ex * ey;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= modes) return;
    const int i = (int)(m % ex), j = jofs + (int)(m / ex);
    const bool active = i < v.mx && j < v.my;
    const double mu = active ? v.lx[i] + v.ly[j] : 0.0;
    const bool pin = v.pin_last && fabs(mu) <= v.mu_eps;
    double cprev = 0.0;
    for (int kl = 0; kl < v.mz; kl++) {
        const int k = v.k0 + kl;
        const double lo = kl > 0 ? v.coef * v.zk[k] : 0.0;
        const double dg = v.coef * (mu * v.zk[3 * nz + k] + v.zk[nz + k]) + v.shift * v.zk[3 * nz + k];
        const double up = v.coef * v.zk[2 * nz + k];
        const double den = dg - lo * cprev;
        double inv = 0.0, cp = 0.0;
        if (active && fabs(den) > thresh && !(pin && kl == v.mz - 1)) { inv = 1.0 / den; cp = (kl + 1 < v.mz) ? up * inv : 0.0; }
        v.inv[(long long)kl * modes + m] = (float)inv;
        v.cp[(long long)kl * modes + m] = (float)cp;
        cprev = cp;
    }
}

struct ThomasArgs {
    float* x[MAXQ];            // in: right-hand sides, out: solutions (in place)
    const float* inv[MAXQ];
    const float* cp[MAXQ];
    const double* zk[MAXQ];    // for the sub-diagonal
    double coef[MAXQ];
    int mz[MAXQ];
    int narr, nz, k0;          // k0: global plane of local plane 0 (z-slabs solve their own block)
    long long modes;
    float* iface;              // optional [narr][2][modes]: first and last unknown of every local solution
};

// one thread per (array, mode); the loads of a block of TB planes are issued before the dependent chain runs
template <int TB>
__global__ void __launch_bounds__(128) tfb_thomas_kernel(const ThomasArgs a) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (m >= a.modes) return;
    float* __restrict__ x = a.x[q] + m;
    const float* __restrict__ inv = a.inv[q] + m;
    const float* __restrict__ cp = a.cp[q] + m;
    const double* __restrict__ lo = a.zk[q] + a.k0;
    const double coef = a.coef[q];
    const int mz = a.mz[q];
    const long long st = a.modes;
    // software pipeline: the loads of block b+1 are issued before the dependent chain of block b runs
    float rp = 0.f;
    float xv[TB], iv[TB], lv[TB], xn_[TB], in_[TB], ln_[TB];
    auto load_fwd = [&](int k0, float (&xx)[TB], float (&ii)[TB], float (&ll)[TB]) {
#pragma unroll
        for (int t = 0; t < TB; t++) {
            const int k = k0 + t;
            const bool ok = k < mz;
            xx[t] = ok ? x[(long long)k * st] : 0.f;
            ii[t] = ok ? inv[(long long)k * st] : 0.f;
            ll[t] = (ok && k > 0) ? (float)(coef * lo[k]) : 0.f;
        }
    };
    load_fwd(0, xv, iv, lv);
    for (int k0 = 0; k0 < mz; k0 += TB) {
        if (k0 + TB < mz) load_fwd(k0 + TB, xn_, in_, ln_);
#pragma unroll
        for (int t = 0; t < TB; t++) {
            rp = (xv[t] - lv[t] * rp) * iv[t];
            if (k0 + t < mz) x[(long long)(k0 + t) * st] = rp;
        }
#pragma unroll
        for (int t = 0; t < TB; t++) { xv[t] = xn_[t]; iv[t] = in_[t]; lv[t] = ln_[t]; }
    }
    float xn = 0.f, xlast = 0.f;
    auto load_bwd = [&](int k1, float (&xx)[TB], float (&cc)[TB]) {
#pragma unroll
        for (int t = 0; t < TB; t++) {
            const int k = k1 - 1 - t;
            const bool ok = k >= 0;
            xx[t] = ok ? x[(long long)k * st] : 0.f;
            cc[t] = ok ? cp[(long long)k * st] : 0.f;
        }
    };
    load_bwd(mz, xv, iv);
    for (int k1 = mz; k1 > 0; k1 -= TB) {
        if (k1 - TB > 0) load_bwd(k1 - TB, xn_, in_);
#pragma unroll
        for (int t = 0; t < TB; t++) {
            if (k1 - 1 - t >= 0) {
                xn = xv[t] - iv[t] * xn;
                x[(long long)(k1 - 1 - t) * st] = xn;
                if (k1 - 1 - t == mz - 1) xlast = xn;
            }
        }
#pragma unroll
        for (int t = 0; t < TB; t++) { xv[t] = xn_[t]; iv[t] = in_[t]; }
    }
    if (a.iface) {
        a.iface[((long long)q * 2 + 0) * st + m] = xn;       // first unknown (computed last)
        a.iface[((long long)q * 2 + 1) * st + m] = xlast;
    }
    // planes mz .. nz-1 (wall unknowns) are not touched here
}

// ---------------------------------------------------------------------------------------------------------
// z-slabs: the tridiagonal systems run through every slab.  Each rank factors and solves ITS block T_g (above), so
//     x_g = y_g - l_{g-1} * v'_g - f_{g+1} * w'_g,      y_g = T_g^-1 r_g,
// with the "spikes" v'_g = lo_g T_g^-1 e_first, w'_g = up_g T_g^-1 e_last (lo_g, up_g: the couplings to the neighbour
// slabs) and l_{g-1}, f_{g+1} the last / first unknowns of the neighbours' SOLUTIONS.  Those follow from a reduced
// system in the 2G interface unknowns (f_g, l_g) per mode,
//     f_g + A_g l_{g-1} + B_g f_{g+1} = yf_g,   l_g + C_g l_{g-1} + D_g f_{g+1} = yl_g,
//     A = v'_first, B = w'_first, C = v'_last, D = w'_last,
// whose matrix depends only on the operator: tfb_spike_weights_kernel inverts it once per parameter set and keeps the
// two rows every rank needs, so a solve is: local sweeps, ONE all-gather of (yf, yl) per array (2 floats per mode and
// rank), two short dot products and one correction pass -- instead of two all-to-all transposes of the whole array.
// ---------------------------------------------------------------------------------------------------------
struct SpikeVar {
    const float* inv; const float* cp; const double* zk;
    float* v; float* w;        // [mz][modes]: the scaled spikes
    float* coef4;              // [4][modes]: A, B, C, D of this rank (all-gathered afterwards)
    double coef;
    int k0, mz, nz_active;     // block offset, local unknowns, global unknowns of this variable along z
    long long modes;
};
__global__ void __launch_bounds__(128) tfb_spike_setup_kernel(SpikeVar s) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= s.modes || s.mz <= 0) return;
    const long long st = s.modes;
    const float* inv = s.inv + m;
    const float* cp = s.cp + m;
    // v = T^-1 e_first: forward sweep of e_first, backward sweep; w = T^-1 e_last: only the last forward entry is non-zero
    float* v = s.v + m;
    float* w = s.w + m;
    double rp = 0.0;
    for (int k = 0; k < s.mz; k++) {
        const double lok = k > 0 ? s.coef * s.zk[s.k0 + k] : 0.0;
        rp = ((k == 0 ? 1.0 : 0.0) - lok * rp) * (double)inv[(long long)k * st];
        v[(long long)k * st] = (float)rp;
    }
    double xv = 0.0, xw = 0.0;
    for (int k = s.mz - 1; k >= 0; k--) {
        xv = (double)v[(long long)k * st] - (double)cp[(long long)k * st] * xv;
        xw = (k == s.mz - 1 ? (double)inv[(long long)k * st] : 0.0) - (double)cp[(long long)k * st] * xw;
        v[(long long)k * st] = (float)xv;
        w[(long long)k * st] = (float)xw;
    }
}
// scale the spikes by the couplings to the neighbour slabs and publish this rank's four coefficients
__global__ void __launch_bounds__(128) tfb_spike_scale_kernel(SpikeVar s, double lo_g, double up_g) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= s.modes) return;
    const long long st = s.modes;
    for (int k = 0; k < s.mz; k++) {
        s.v[(long long)k * st + m] = (float)(lo_g * (double)s.v[(long long)k * st + m]);
        s.w[(long long)k * st + m] = (float)(up_g * (double)s.w[(long long)k * st + m]);
    }
    const bool any = s.mz > 0;
    s.coef4[0 * st + m] = any ? s.v[m] : 0.f;                                     // A: v'_first
    s.coef4[1 * st + m] = any ? s.w[m] : 0.f;                                     // B: w'_first
    s.coef4[2 * st + m] = any ? s.v[(long long)(s.mz - 1) * st + m] : 0.f;        // C: v'_last
    s.coef4[3 * st + m] = any ? s.w[(long long)(s.mz - 1) * st + m] : 0.f;        // D: w'_last
}
// rows of the inverse of the reduced system that rank `me` needs: weights[0] -> l_{me-1}, weights[1] -> f_{me+1}; each is a
// vector over the 2G gathered values ordered (yf_0, yl_0, yf_1, yl_1, ...).  coef_all: [G][4][modes].
#define TFB_SPIKE_MAXG 16
__global__ void __launch_bounds__(64) tfb_spike_weights_kernel(int G, int me, long long modes, const float* __restrict__ coef_all,
                                                               float* __restrict__ weights) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= modes) return;
    const int n = 2 * G;
    double M[2 * TFB_SPIKE_MAXG][2 * TFB_SPIKE_MAXG];
    double R[2 * TFB_SPIKE_MAXG][2];      // right-hand sides of M^T w = e_p, e_q
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++) M[r][c] = r == c ? 1.0 : 0.0;
    for (int g = 0; g < G; g++) {
        const double A = coef_all[((long long)g * 4 + 0) * modes + m], B = coef_all[((long long)g * 4 + 1) * modes + m];
        const double C = coef_all[((long long)g * 4 + 2) * modes + m], D = coef_all[((long long)g * 4 + 3) * modes + m];
        // M^T is what is factored: entry (row r, col c) of M goes to M[c][r]
        if (g > 0) { M[2 * g - 1][2 * g] = A; M[2 * g - 1][2 * g + 1] = C; }
        if (g < G - 1) { M[2 * g + 2][2 * g] = B; M[2 * g + 2][2 * g + 1] = D; }
    }
    const int p = 2 * (me - 1) + 1, q = 2 * (me + 1);
    for (int r = 0; r < n; r++) { R[r][0] = (me > 0 && r == p) ? 1.0 : 0.0; R[r][1] = (me < G - 1 && r == q) ? 1.0 : 0.0; }
    // Gaussian elimination without pivoting (the system is diagonally dominant: |spike tips| < 1); band of width 3
    for (int k = 0; k < n; k++) {
        const double piv = M[k][k];
        const double ip = fabs(piv) > 1e-12 ? 1.0 / piv : 0.0;
        const int rmax = k + 3 < n ? k + 3 : n - 1;
        for (int r = k + 1; r <= rmax; r++) {
            const double fct = M[r][k] * ip;
            if (fct == 0.0) continue;
            for (int c = k; c <= (k + 3 < n ? k + 3 : n - 1); c++) M[r][c] -= fct * M[k][c];
            R[r][0] -= fct * R[k][0]; R[r][1] -= fct * R[k][1];
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        const double piv = M[k][k];
        const double ip = fabs(piv) > 1e-12 ? 1.0 / piv : 0.0;
        double a0 = R[k][0], a1 = R[k][1];
        for (int c = k + 1; c <= (k + 3 < n ? k + 3 : n - 1); c++) { a0 -= M[k][c] * R[c][0]; a1 -= M[k][c] * R[c][1]; }
        R[k][0] = a0 * ip; R[k][1] = a1 * ip;
    }
    for (int r = 0; r < n; r++) {
        weights[((long long)0 * n + r) * modes + m] = (float)R[r][0];
        weights[((long long)1 * n + r) * modes + m] = (float)R[r][1];
    }
}
// x = y - l_{g-1} v' - f_{g+1} w' with (l_{g-1}, f_{g+1}) = weights . gathered interface values
struct SpikeFixArgs {
    float* x[MAXQ];
    const float* v[MAXQ]; const float* w[MAXQ];
    const float* weights[MAXQ];   // [2][2G][modes]
    int mz[MAXQ];
    const float* gathered;        // [G][narr][2][modes]
    int narr, G;
    long long modes;
};
__global__ void __launch_bounds__(128) tfb_spike_fix_kernel(const SpikeFixArgs a) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;
    if (m >= a.modes) return;
    const long long st = a.modes;
    const int n = 2 * a.G;
    double lp = 0.0, fn = 0.0;
    for (int g = 0; g < a.G; g++) {
        const float* y = a.gathered + ((long long)(g * a.narr + q) * 2) * st + m;
        const double yf = y[0], yl = y[st];
        lp += (double)a.weights[q][((long long)0 * n + 2 * g) * st + m] * yf + (double)a.weights[q][((long long)0 * n + 2 * g + 1) * st + m] * yl;
        fn += (double)a.weights[q][((long long)1 * n + 2 * g) * st + m] * yf + (double)a.weights[q][((long long)1 * n + 2 * g + 1) * st + m] * yl;
    }
    const float lpf = (float)lp, fnf = (float)fn;
    float* x = a.x[q] + m;
    const float* v = a.v[q] + m;
    const float* w = a.w[q] + m;
    for (int k = 0; k < a.mz[q]; k++) x[(long long)k * st] -= lpf * v[(long long)k * st] + fnf * w[(long long)k * st];
}

// interleaved fp64 vector -> fp32 SoA arrays of `nv` variables:  comp[v][cell] = r[cell*dof + var[v]] - sub[...]
struct DeintArgs {
    float* comp[MAXQ];
    int var[MAXQ];
    int nv, dof;
    long long ncell;
};
__global__ void __launch_bounds__(256) tfb_deint_kernel(const DeintArgs a, const double* __restrict__ r, const double* __restrict__ sub) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < a.ncell; cell += (long long)gridDim.x * blockDim.x) {
#pragma unroll 4
        for (int v = 0; v < a.nv; v++) {
            const long long row = cell * a.dof + a.var[v];
            a.comp[v][cell] = (float)(sub ? r[row] - sub[row] : r[row]);
        }
    }
}
// fp32 SoA arrays -> rows of the interleaved fp64 vector; unknowns outside the active extents (wall-normal
// boundary velocities, row = -1 * u) get -r
struct IntArgs {
    const float* comp[MAXQ];
    int var[MAXQ];
    int mx[MAXQ], my[MAXQ], mz[MAXQ];
    int nv, dof, nx, ny, k0;
    long long ncell;
};
__global__ void __launch_bounds__(256) tfb_int_kernel(const IntArgs a, const double* __restrict__ r, const double* __restrict__ sub,
                                                      double* __restrict__ z) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < a.ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(cell % a.nx), j = (int)((cell / a.nx) % a.ny), k = a.k0 + (int)(cell / ((long long)a.nx * a.ny));
#pragma unroll 4
        for (int v = 0; v < a.nv; v++) {
            const long long row = cell * a.dof + a.var[v];
            const bool wall = i >= a.mx[v] || j >= a.my[v] || k >= a.mz[v];
            z[row] = wall ? -(sub ? r[row] - sub[row] : r[row]) : (double)a.comp[v][cell];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused head and tail of the scaled-mass block preconditioner (velocity-pressure problems):
//   head:  dp = gamma r_p / |cell|  (pinned cell: -r_p),  comp[v] = (float)(r_v - (G dp)_v)
//   tail:  z_v = FDM result (wall-normal boundary unknowns: -r_v),  z_p = dp
// G is the gradient block of the matrix in a two-slot form: row (cell, v) couples to the pressure of the same cell
// (slot 0) and of the next cell along axis v (slot 1) -- tfb_gell_build_kernel copies the values out of the compact
// sub-matrix and counts entries that do not fit (then the caller keeps the general path).
// ---------------------------------------------------------------------------------------------------------
struct GEll {
    float* val;            // [dim][2][ncell]
    int* misfit;           // device counter
};
__global__ void __launch_bounds__(256) tfb_gell_build_kernel(long long ncell, int dof, int dim, int nx, int ny, long long cell0,
                                                             const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                             const double* __restrict__ vals, GEll g) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    const long long stride[3] = {1, nx, (long long)nx * ny};
    for (int v = 0; v < dim; v++) {
        const long long row = cell * dof + v;
        float s0 = 0.f, s1 = 0.f;
        for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) {
            const long long pc = (long long)col[e] / dof - cell0;      // local cell of the pressure unknown
            if (pc == cell) s0 += (float)vals[e];
            else if (pc == cell + stride[v]) s1 += (float)vals[e];
            else if (vals[e] != 0.0) atomicAdd(g.misfit, 1);
        }
        g.val[((long long)v * 2 + 0) * ncell + cell] = s0;
        g.val[((long long)v * 2 + 1) * ncell + cell] = s1;
    }
}

struct PreArgs {
    float* comp[3];
    float* dp;             // [ncell + ghost plane above]: the pressure update, also read by the tail
    const float* gval;     // GEll::val
    const double* hx; const double* hy; const double* hz;
    double gamma;
    long long ncell, pin_local;   // pin_local: local cell of the pinned pressure or -1
    int dof, dim, nx, ny, k0;
};
// pass 1: dp (one thread per cell, plus the plane above the slab when it exists in `r_above`)
__global__ void __launch_bounds__(256) tfb_tc_dp_kernel(const PreArgs a, const double* __restrict__ r) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < a.ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(cell % a.nx), j = (int)((cell / a.nx) % a.ny), k = a.k0 + (int)(cell / ((long long)a.nx * a.ny));
        const double rp = r[cell * a.dof + a.dim];
        a.dp[cell] = (float)(cell == a.pin_local ? -rp : a.gamma * rp / ((a.hx[i] * a.hy[j]) * a.hz[k]));
    }
}
// pass 2: comp[v] = r_v - g0 dp(cell) - g1 dp(next cell along v)
__global__ void __launch_bounds__(256) tfb_tc_pre_kernel(const PreArgs a, const double* __restrict__ r) {
    const long long stride[3] = {1, a.nx, (long long)a.nx * a.ny};
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < a.ncell; cell += (long long)gridDim.x * blockDim.x) {
        const float d0 = a.dp[cell];
#pragma unroll
        for (int v = 0; v < 3; v++) {
            if (v >= a.dim) break;
            const float g0 = a.gval[((long long)v * 2 + 0) * a.ncell + cell], g1 = a.gval[((long long)v * 2 + 1) * a.ncell + cell];
            double acc = r[cell * a.dof + v] - (double)g0 * (double)d0;
            if (g1 != 0.f) acc -= (double)g1 * (double)a.dp[cell + stride[v]];
            a.comp[v][cell] = (float)acc;
        }
    }
}
// tail: all rows of z
__global__ void __launch_bounds__(256) tfb_tc_post_kernel(const IntArgs a, const float* __restrict__ dp, int pvar,
                                                          const double* __restrict__ r, double* __restrict__ z) {
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < a.ncell; cell += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(cell % a.nx), j = (int)((cell / a.nx) % a.ny), k = a.k0 + (int)(cell / ((long long)a.nx * a.ny));
#pragma unroll 4
        for (int v = 0; v < a.nv; v++) {
            const long long row = cell * a.dof + a.var[v];
            const bool wall = i >= a.mx[v] || j >= a.my[v] || k >= a.mz[v];
            z[row] = wall ? -r[row] : (double)a.comp[v][cell];
        }
        z[cell * a.dof + pvar] = (double)dp[cell];
    }
}

// Vectorised head / tail for the 3-D velocity-pressure layout (dof = 4): blockIdx.y is the plane, one thread per cell moves
// the cell's four fp64 values as two 16-byte accesses (a warp covers 1 KB contiguously) and the SoA arrays coalesced;
// index arithmetic is 32-bit with one division per cell.  The head recomputes the pressure update of the neighbouring
// cells from r_p directly (no separate dp pass; `r` must have its halo plane above the slab in place on z-slabs).
// ihx/ihy/ihz are reciprocal cell widths (1 / |cell| = ihx ihy ihz).
struct Pre4Args {
    float* comp[3];
    float* dp;
    const float* gval;
    const double* ihx; const double* ihy; const double* ihz;
    double gamma;
    long long ncell, pin_local;
    int nx, ny, k0;
};
__global__ void __launch_bounds__(256) tfb_tc_pre4_kernel(const Pre4Args a, const double* __restrict__ r) {
    const unsigned plane_cells = (unsigned)a.nx * (unsigned)a.ny;
    const unsigned kl = blockIdx.y;
    const long long cell0 = (long long)kl * plane_cells;
    const double ihz = a.ihz[a.k0 + kl], ihz1 = a.ihz[a.k0 + kl + 1];
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(r);
    float* __restrict__ c0 = a.comp[0];
    float* __restrict__ c1 = a.comp[1];
    float* __restrict__ c2 = a.comp[2];
    for (unsigned pc = blockIdx.x * blockDim.x + threadIdx.x; pc < plane_cells; pc += gridDim.x * blockDim.x) {
        const unsigned j = pc / (unsigned)a.nx, i = pc - j * (unsigned)a.nx;
        const long long cell = cell0 + pc;
        const double2 uv = r2[cell * 2], wp = r2[cell * 2 + 1];
        const double ihxy = a.ihx[i] * a.ihy[j];
        auto dp_of = [&](long long cc, double rp, double ivol) { return (double)(float)(cc == a.pin_local ? -rp : a.gamma * rp * ivol); };
        const double d0 = dp_of(cell, wp.y, ihxy * ihz);
        a.dp[cell] = (float)d0;
        const float* g = a.gval + cell;
        const float gu0 = g[0], gu1 = g[a.ncell], gv0 = g[2 * a.ncell], gv1 = g[3 * a.ncell], gw0 = g[4 * a.ncell], gw1 = g[5 * a.ncell];
        double au = uv.x - (double)gu0 * d0, av = uv.y - (double)gv0 * d0, aw = wp.x - (double)gw0 * d0;
        if (gu1 != 0.f) au -= (double)gu1 * dp_of(cell + 1, r[(cell + 1) * 4 + 3], (a.ihx[i + 1] * a.ihy[j]) * ihz);
        if (gv1 != 0.f) av -= (double)gv1 * dp_of(cell + a.nx, r[(cell + a.nx) * 4 + 3], (a.ihx[i] * a.ihy[j + 1]) * ihz);
        if (gw1 != 0.f) aw -= (double)gw1 * dp_of(cell + plane_cells, r[(cell + plane_cells) * 4 + 3], ihxy * ihz1);
        c0[cell] = (float)au; c1[cell] = (float)av; c2[cell] = (float)aw;
    }
}
__global__ void __launch_bounds__(256) tfb_tc_post4_kernel(const IntArgs a, const float* __restrict__ dp,
                                                           const double* __restrict__ r, double* __restrict__ z) {
    const unsigned plane_cells = (unsigned)a.nx * (unsigned)a.ny;
    const unsigned kl = blockIdx.y;
    const int k = a.k0 + (int)kl;
    const long long cell0 = (long long)kl * plane_cells;
    double2* __restrict__ z2 = reinterpret_cast<double2*>(z);
    const float* __restrict__ c0 = a.comp[0];
    const float* __restrict__ c1 = a.comp[1];
    const float* __restrict__ c2 = a.comp[2];
    for (unsigned pc = blockIdx.x * blockDim.x + threadIdx.x; pc < plane_cells; pc += gridDim.x * blockDim.x) {
        const unsigned j = pc / (unsigned)a.nx, i = pc - j * (unsigned)a.nx;
        const long long cell = cell0 + pc;
        const bool wu = (int)i >= a.mx[0] || (int)j >= a.my[0] || k >= a.mz[0];
        const bool wv = (int)i >= a.mx[1] || (int)j >= a.my[1] || k >= a.mz[1];
        const bool ww = (int)i >= a.mx[2] || (int)j >= a.my[2] || k >= a.mz[2];
        double2 uv, wp;
        uv.x = wu ? -r[cell * 4 + 0] : (double)c0[cell];
        uv.y = wv ? -r[cell * 4 + 1] : (double)c1[cell];
        wp.x = ww ? -r[cell * 4 + 2] : (double)c2[cell];
        wp.y = (double)dp[cell];
        z2[cell * 2] = uv;
        z2[cell * 2 + 1] = wp;
    }
}
#endif  // __CUDACC__

}  // namespace tfbtc
