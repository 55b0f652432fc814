// libtfb200 core: context management, fixed-pattern construction, assembly launches.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include "tfb_assemble.cuh"
#include "tfb_spmv_march.cuh"

struct TfbWiden { __host__ __device__ long long operator()(int v) const { return (long long)v; } };

thread_local std::string g_tfb_err;
int64_t g_tfb_launches = 0;

int tfb_fail(const char* file, int line, const char* what, const char* detail) {
    char buf[1024];
    snprintf(buf, sizeof buf, "%s:%d: %s: %s", file, line, what, detail ? detail : "");
    g_tfb_err = buf;
    cudaGetLastError();   // do not let this failure resurface at the next launch check
    return -1;
}

extern "C" const char* tfb_last_error(void) { return g_tfb_err.c_str(); }
extern "C" int64_t tfb_launch_count(void) { return g_tfb_launches; }

extern "C" int tfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* tfb_config_name(int config) {
#define X(C) if (config == C::ID) return C::NAME;
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return nullptr;
}

TfbGrid tfb_ctx::grid() const {
    TfbGrid g;
    g.nx = desc.nx; g.ny = desc.ny; g.nz = desc.nz; g.dim = desc.dim; g.dof = desc.dof;
    g.zfold = desc.nz == 1;
    for (int a = 0; a < 3; a++) g.met[a] = d_met[a];
    g.cor = d_cor;
    for (int f = 0; f < TFB_MAX_FORCE; f++) { g.fval[f] = d_fval[f]; g.fdir[f] = fdir[f]; }
    return g;
}

static int config_dof(int config) {
#define X(C) if (config == C::ID) return C::DOF;
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return -1;
}

extern "C" int tfb_create(const tfb_desc* d, tfb_ctx** out) {
    TFB_CHECK(d && out, "null argument");
    TFB_CHECK(tfb_config_name(d->config) != nullptr, "unknown kernel config");
    TFB_CHECK(config_dof(d->config) == d->dof, "dof does not match the kernel config");
    TFB_CHECK(d->nx >= 2 && d->ny >= 2 && d->nz >= 1, "grid must be at least 2x2x1");
    TFB_CHECK(d->k0 >= 0 && d->k1 <= d->nz && d->k0 < d->k1, "bad z-slab");
    TFB_CHECK((int64_t)d->nx * d->ny * d->nz * d->dof < (int64_t)INT32_MAX, "more than 2^31 unknowns");
    TFB_CUDA(cudaSetDevice(d->device));
    tfb_ctx* c = new tfb_ctx();
    c->desc = *d;
    c->nzl = d->k1 - d->k0;
    c->plane_rows = (int64_t)d->nx * d->ny * d->dof;
    c->n_local = c->plane_rows * c->nzl;
    c->n_global = c->plane_rows * d->nz;
    c->row0 = c->plane_rows * d->k0;
    TFB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const int n[3] = {d->nx, d->ny, d->nz};
    for (int a = 0; a < 3; a++) {
        TFB_CUDA(cudaMalloc(&c->d_met[a], sizeof(double) * TFB_NMET * n[a]));
        TFB_CUDA(cudaMemcpy(c->d_met[a], d->met[a], sizeof(double) * TFB_NMET * n[a], cudaMemcpyHostToDevice));
    }
    TFB_CUDA(cudaMalloc(&c->d_cor, sizeof(double) * 2 * d->ny));
    TFB_CUDA(cudaMemcpy(c->d_cor, d->cor, sizeof(double) * 2 * d->ny, cudaMemcpyHostToDevice));
    TFB_CUDA(cudaMalloc(&c->d_state, sizeof(double) * c->plane_rows * (c->nzl + 2)));
    TFB_CUDA(cudaMemset(c->d_state, 0, sizeof(double) * c->plane_rows * (c->nzl + 2)));
    TFB_CUDA(cudaMalloc(&c->d_rhs, sizeof(double) * c->n_local));
    TFB_CUDA(cudaMalloc(&c->d_frc_static, sizeof(double) * c->n_local));
    for (int e = 0; e < 16; e++) TFB_CUDA(cudaEventCreate(&c->ev[e]));     // the others are created on first use
    memset(&c->prm, 0, sizeof c->prm);
    c->desc.met[0] = c->desc.met[1] = c->desc.met[2] = nullptr;
    c->desc.cor = nullptr;
    *out = c;
    int rc = tfb_build_pattern(c);
    if (rc != 0) { tfb_destroy(c); *out = nullptr; return rc; }
    return 0;
}

extern "C" void tfb_destroy(tfb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->desc.device);
    for (int a = 0; a < 3; a++) cudaFree(c->d_met[a]);
    cudaFree(c->d_cor);
    for (int f = 0; f < TFB_MAX_FORCE; f++) cudaFree(c->d_fval[f]);
    cudaFree(c->d_frc_static);
    cudaFree(c->d_state);
    cudaFree(c->d_rhs);
    cudaFree(c->d_row_ptr);
    cudaFree(c->d_col);
    cudaFree(c->d_flush);
    cudaFree(c->d_massdiag);
    for (double* p : c->vals_pool) cudaFree(p);
    tfb_solver_free(c->solver);
    tfb_direct_pool_free(c);
    for (int e = 0; e < TFB_EVENT_SLOTS; e++) if (c->ev[e]) cudaEventDestroy(c->ev[e]);
    for (auto e : c->ev_comm) if (e) cudaEventDestroy(e);
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int tfb_set_params(tfb_ctx* c, const tfb_params* prm, const double* const* fval,
                              const int8_t* fdir, const double* frc_static) {
    TFB_CHECK(c && prm, "null argument");
    static_assert(sizeof(tfb_params) == sizeof(TfbParams), "ABI struct mismatch");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    memcpy(&c->prm, prm, sizeof(TfbParams));
    c->have_params = true;
    for (int f = 0; f < TFB_MAX_FORCE; f++) {
        c->fdir[f] = fdir ? fdir[f] : 0;
        const double* src = fval ? fval[f] : nullptr;
        if (src) {
            int a = c->fdir[f];
            // in-plane extent of the face, restricted to the owned planes where z is in-plane
            size_t n1 = a == 0 ? c->desc.ny : c->desc.nx;
            size_t n2 = a == 2 ? c->desc.ny : c->desc.nz;
            if (!c->d_fval[f]) TFB_CUDA(cudaMalloc(&c->d_fval[f], sizeof(double) * n1 * n2));
            TFB_CUDA(cudaMemcpyAsync(c->d_fval[f], src, sizeof(double) * n1 * n2, cudaMemcpyHostToDevice, c->stream));
        } else if (c->d_fval[f]) {
            cudaFree(c->d_fval[f]);
            c->d_fval[f] = nullptr;
        }
    }
    c->has_frc_static = frc_static != nullptr;
    if (frc_static)
        TFB_CUDA(cudaMemcpyAsync(c->d_frc_static, frc_static, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int tfb_sizes(tfb_ctx* c, int64_t* n_local, int64_t* nnz_local, int64_t* n_global, int64_t* row0) {
    TFB_CHECK(c, "null ctx");
    if (n_local) *n_local = c->n_local;
    if (nnz_local) *nnz_local = c->nnz;
    if (n_global) *n_global = c->n_global;
    if (row0) *row0 = c->row0;
    return 0;
}

template <class Cfg>
static int build_pattern_t(tfb_ctx* c) {
    const long long nrows = c->n_local;
    TfbGrid g = c->grid();
    int* d_counts = nullptr;
    TFB_CUDA(cudaMalloc(&d_counts, sizeof(int) * (nrows + 1)));
    TFB_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * (nrows + 1), c->stream));
    const int bs = 256;
    const unsigned nb = (unsigned)((nrows + bs - 1) / bs);
    tfb_count_kernel<Cfg><<<nb, bs, 0, c->stream>>>(g, c->desc.k0, nrows, d_counts);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    {   // the device pattern is int32: refuse grids whose slab holds 2^31 or more structural non-zeros (a 330^3 cavity
        // still fits in 180 GB) instead of letting the scan wrap around
        long long* d_total = nullptr;
        void* t2 = nullptr;
        size_t t2_bytes = 0;
        TFB_CUDA(cudaMalloc(&d_total, sizeof(long long)));
        auto as64 = thrust::make_transform_iterator(d_counts, TfbWiden());
        TFB_CUDA(cub::DeviceReduce::Sum(nullptr, t2_bytes, as64, d_total, nrows, c->stream));
        TFB_CUDA(cudaMalloc(&t2, t2_bytes));
        TFB_CUDA(cub::DeviceReduce::Sum(t2, t2_bytes, as64, d_total, nrows, c->stream));
        long long total = 0;
        TFB_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(t2); cudaFree(d_total);
        if (total >= (long long)INT32_MAX) { cudaFree(d_counts); return tfb_fail(__FILE__, __LINE__, "tfb_build_pattern", "2^31 or more structural non-zeros per slab: split the grid over more GPUs"); }
    }
    TFB_CUDA(cudaMalloc(&c->d_row_ptr, sizeof(int) * (nrows + 1)));
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    TFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, c->d_row_ptr, nrows + 1, c->stream));
    TFB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    TFB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_counts, c->d_row_ptr, nrows + 1, c->stream));
    TFB_LAUNCHED();
    int nnz = 0;
    TFB_CUDA(cudaMemcpyAsync(&nnz, c->d_row_ptr + nrows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(tmp);
    cudaFree(d_counts);
    TFB_CHECK(nnz > 0, "empty pattern");
    c->nnz = nnz;
    TFB_CUDA(cudaMalloc(&c->d_col, sizeof(int) * (size_t)nnz));
    tfb_fill_cols_kernel<Cfg><<<nb, bs, 0, c->stream>>>(g, c->desc.k0, nrows, c->d_row_ptr, c->d_col);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_pattern = true;
    return 0;
}

int tfb_build_pattern(tfb_ctx* c) {
    // nnz per slab can be counted in 64 bits on the host side later; the device pattern uses int32.
#define X(C) if (c->desc.config == C::ID) return build_pattern_t<C>(c);
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return tfb_fail(__FILE__, __LINE__, "tfb_build_pattern", "unknown config");
}

extern "C" int tfb_get_pattern(tfb_ctx* c, int64_t* row_ptr, int64_t* col_idx) {
    TFB_CHECK(c && c->have_pattern, "no pattern");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    std::vector<int> tmp((size_t)std::max<int64_t>(c->n_local + 1, c->nnz));
    if (row_ptr) {
        TFB_CUDA(cudaMemcpy(tmp.data(), c->d_row_ptr, sizeof(int) * (c->n_local + 1), cudaMemcpyDeviceToHost));
        for (int64_t r = 0; r <= c->n_local; r++) row_ptr[r] = tmp[r];
    }
    if (col_idx) {
        TFB_CUDA(cudaMemcpy(tmp.data(), c->d_col, sizeof(int) * c->nnz, cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < c->nnz; e++) col_idx[e] = tmp[e];
    }
    return 0;
}

// versions are unique across matrices: the solver's caches are keyed on (address, version) and a new tfb_mat can
// land on the address of a destroyed one
uint64_t tfb_next_version() { static uint64_t v = 0; return ++v; }

extern "C" int tfb_mat_create(tfb_ctx* c, tfb_mat** out) {
    TFB_CHECK(c && out && c->have_pattern, "no pattern");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    tfb_mat* m = new tfb_mat();
    m->ctx = c;
    // +2: the structured SpMV fetches whole 16-byte pairs and may read one element past the last span
    if (!c->vals_pool.empty()) {
        m->d_vals = c->vals_pool.back();
        c->vals_pool.pop_back();
    } else {
        cudaError_t e = cudaMalloc(&m->d_vals, sizeof(double) * ((size_t)c->nnz + 2));
        if (e != cudaSuccess) { delete m; return tfb_fail(__FILE__, __LINE__, "cudaMalloc(values)", cudaGetErrorString(e)); }
    }
    TFB_CUDA(cudaMemset(m->d_vals, 0, sizeof(double) * ((size_t)c->nnz + 2)));
    *out = m;
    return 0;
}

extern "C" void tfb_mat_destroy(tfb_mat* m) {
    if (!m) return;
    tfb_ctx* c = m->ctx;
    cudaSetDevice(c->desc.device);
    tfb_direct_free(m);
    if (m->d_vals && c->vals_pool.size() < 2) {
        cudaDeviceSynchronize();            // what cudaFree did implicitly: nothing in flight still reads the buffer
        c->vals_pool.push_back(m->d_vals);
    } else {
        cudaFree(m->d_vals);
    }
    delete m;
}

extern "C" int tfb_mat_get_values(tfb_mat* m, double* out) {
    TFB_CHECK(m && out, "null argument");
    TFB_CUDA(cudaSetDevice(m->ctx->desc.device));
    TFB_CUDA(cudaMemcpyAsync(out, m->d_vals, sizeof(double) * (size_t)m->ctx->nnz, cudaMemcpyDeviceToHost, m->ctx->stream));
    TFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return 0;
}

extern "C" int tfb_mat_set_values(tfb_mat* m, const double* in) {
    TFB_CHECK(m && in, "null argument");
    TFB_CUDA(cudaSetDevice(m->ctx->desc.device));
    TFB_CUDA(cudaMemcpyAsync(m->d_vals, in, sizeof(double) * (size_t)m->ctx->nnz, cudaMemcpyHostToDevice, m->ctx->stream));
    TFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    m->version = tfb_next_version();
    return 0;
}

// ------------------------------- assembly launches -------------------------------
template <class Cfg, bool DO_J, bool DO_F, int TJ, int MINB>
static int launch_assemble_v(tfb_ctx* c, tfb_mat* m) {
    TfbAsmArgs a;
    a.g = c->grid();
    a.prm = c->prm;
    a.state = c->d_state;
    a.frc_static = c->has_frc_static ? c->d_frc_static : nullptr;
    a.row_ptr = c->d_row_ptr;
    a.vals = m ? m->d_vals : nullptr;
    a.rhs = c->d_rhs;
    a.k0 = c->desc.k0;
    a.nzl = c->nzl;
    a.kc0 = 0;
    a.kstep = 1;
    a.kofs0 = 0; a.klim = c->nzl;
    const bool ends_only = c->win1 == -2;     // the first and the last plane of the slab (the ones that read the halo)
    if (ends_only) a.kstep = c->nzl - 1;
    constexpr int LINE_CAP = TFB_TI * TfbTile<Cfg>::cell_slots() + 2;
    size_t smem = sizeof(double) * (((TfbTile<Cfg>::template state_doubles<TJ>() + 1) & ~1) + (DO_J ? TJ * LINE_CAP : 0));
    auto kern = tfb_assemble_kernel<Cfg, DO_J, DO_F, TJ, MINB>;
    static unsigned configured_devices = 0;       // the attribute is per device
    if (!((configured_devices >> (c->desc.device & 31)) & 1u)) {
        TFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_devices |= 1u << (c->desc.device & 31);
    }
    dim3 block(32, Cfg::DOF, TJ);
    dim3 grid((c->desc.nx + TFB_TI - 1) / TFB_TI, (c->desc.ny + TJ - 1) / TJ, ends_only ? 2 : c->nzl);
    kern<<<grid, block, smem, c->stream>>>(a);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

#define TFB_KCH 16   // planes per z-chunk of the marching kernel (also the host pipeline granularity)

template <class Cfg, bool DO_J, bool DO_F, int TJ, int KCH, int MINB, int EXP = 0, int OPT = 0>
static int launch_march_v(tfb_ctx* c, tfb_mat* m) {
    TfbAsmArgs a;
    a.g = c->grid();
    a.prm = c->prm;
    a.state = c->d_state;
    a.frc_static = c->has_frc_static ? c->d_frc_static : nullptr;
    a.row_ptr = c->d_row_ptr;
    a.vals = m ? m->d_vals : nullptr;
    a.rhs = c->d_rhs;
    a.k0 = c->desc.k0;
    a.nzl = c->nzl;
    a.kstep = KCH;
    a.kofs0 = c->win1 >= 0 ? c->win0 : 0;        // plane window [win0, win1) of this launch (pieces of the host pipeline,
    a.klim = c->win1 >= 0 ? c->win1 : c->nzl;    // interior planes of a z-slab); the whole slab otherwise
    a.kc0 = 0;
    const int nlaunch = (a.klim - a.kofs0 + a.kstep - 1) / a.kstep;
    size_t smem = sizeof(double) * TfbMarch<Cfg, TJ>::smem_doubles(DO_J);
    auto kern = tfb_assemble_march_kernel<Cfg, DO_J, DO_F, TJ, KCH, MINB, EXP, OPT>;
    static unsigned configured_devices = 0;       // the attribute is per device
    if (!((configured_devices >> (c->desc.device & 31)) & 1u)) {
        TFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_devices |= 1u << (c->desc.device & 31);
    }
    dim3 block(32, Cfg::DOF, TJ);
    dim3 grid((c->desc.nx + TFB_TI - 1) / TFB_TI, (c->desc.ny + TJ - 1) / TJ, nlaunch);
    kern<<<grid, block, smem, c->stream>>>(a);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

// Kernel shape per configuration.  The alternatives that were measured against the defaults (tile shapes, occupancy caps,
// ablations) are only compiled with -DTFB_ASM_EXPERIMENTS (TFB_ASM_EXPERIMENTS=1 python -m transiflow_b200.build) and
// selected with TFB_ASM_VARIANT; tools/asm_time.py times them.
template <class Cfg, bool DO_J, bool DO_F>
static int launch_assemble_t(tfb_ctx* c, tfb_mat* m) {
#ifdef TFB_ASM_EXPERIMENTS
    static int variant = -1;
    if (variant < 0) {
        const char* e = getenv("TFB_ASM_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    if constexpr (Cfg::ID == 1 && DO_J && DO_F) {   // Cfg_ldc3d
        switch (variant) {
        case 16: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2>(c, m);      // round-1 default
        case 17: return launch_march_v<Cfg, DO_J, DO_F, 4, 16, 1>(c, m);
        case 24: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 3>(c, m);
        case 18: return launch_march_v<Cfg, DO_J, DO_F, 2, 32, 2>(c, m);
        case 19: return launch_march_v<Cfg, DO_J, DO_F, 2, 8, 2>(c, m);
        case 20: return launch_march_v<Cfg, DO_J, DO_F, 3, 16, 1>(c, m);
        case 22: return launch_march_v<Cfg, DO_J, DO_F, 1, 16, 4>(c, m);
        case 41: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2, 1>(c, m);   // every warp on the BC-free path (timing only)
        case 42: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2, 2>(c, m);   // no bulk stores
        case 44: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2, 4>(c, m);   // no row arithmetic
        case 43: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2, 3>(c, m);
        case 51: return launch_march_v<Cfg, DO_J, DO_F, 2, 16, 2, 0, 1>(c, m);
        case 53: return launch_march_v<Cfg, DO_J, DO_F, 1, 16, 3, 0, 1>(c, m);
        default: break;
        }
    }
    if constexpr (!Cfg::FLAT && Cfg::DOF >= 5 && DO_J && DO_F) {
        if (c->win1 != -2) {
            if (variant == 31) return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 2>(c, m);
            if (variant == 32) return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 1>(c, m);      // round-1 default
            if (variant == 34) return launch_march_v<Cfg, DO_J, DO_F, 1, TFB_KCH, 3>(c, m);
            if (variant == 36) return launch_march_v<Cfg, DO_J, DO_F, 1, TFB_KCH, 2>(c, m);
            if (variant == 38) return launch_march_v<Cfg, DO_J, DO_F, 1, TFB_KCH, 2, 0, 1>(c, m);
            if (variant == 40) return launch_march_v<Cfg, DO_J, DO_F, 3, TFB_KCH, 1, 0, 1>(c, m);
            if (variant == 45) return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 1, 2, 1>(c, m);   // no bulk stores
            if (variant == 46) return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 1, 4, 1>(c, m);   // no row arithmetic
            if (variant == 47) return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 1, 1, 1>(c, m);   // all warps BC-free
        }
    }
#endif
    if constexpr (!Cfg::FLAT) {
        // the two planes of a z-slab that read the halo, after the exchange that ran next to the interior planes: one launch
        // of the plane-tile kernel (a marching CTA would fill its ring for a single plane)
        if (c->win1 == -2) return launch_assemble_v<Cfg, DO_J, DO_F, 2, 2>(c, m);
        // true 3-D grids: z-marching kernel (ring of state planes, software prefetch)
        if constexpr (Cfg::DOF >= 5) {
            // dof 5: 168 registers, no spills, one CTA of ten warps per SM; the warp -> (equation, line) table evens the
            // four schedulers out (0.477 -> 0.453 ms at 128^3)
            return launch_march_v<Cfg, DO_J, DO_F, 2, TFB_KCH, 1, 0, 1>(c, m);
        }
        // dof 4: one line per CTA, four CTAs per SM, equations rotated over the schedulers with the CTA index
        // (two lines x two CTAs: 0.238 ms, one line x four CTAs: 0.246 ms, with the rotation: 0.231 ms at 128^3)
        else return launch_march_v<Cfg, DO_J, DO_F, 1, TFB_KCH, 4, 0, 1>(c, m);
    } else {
        return launch_assemble_v<Cfg, DO_J, DO_F, (Cfg::DOF >= 5 ? 3 : 4), 1>(c, m);
    }
}

template <class Cfg>
static int launch_assemble(tfb_ctx* c, tfb_mat* m, bool do_j, bool do_f) {
    if (do_j && do_f) return launch_assemble_t<Cfg, true, true>(c, m);
    if (do_j) return launch_assemble_t<Cfg, true, false>(c, m);
    return launch_assemble_t<Cfg, false, true>(c, m);
}

static int dispatch_assemble(tfb_ctx* c, tfb_mat* m, int do_j, int do_f);

extern "C" int tfb_state_upload(tfb_ctx* c, const double* state) {
    TFB_CHECK(c && state, "null argument");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    TFB_CUDA(cudaMemcpyAsync(c->d_state + c->plane_rows, state, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    c->state_uploads++;
    return 0;
}

extern "C" int64_t tfb_upload_count(tfb_ctx* c) { return c ? c->state_uploads : -1; }

// 64-bit wrapping sum and xor of the words of a host vector, split over a few threads (memory-bandwidth bound: about
// 1 ms for the 67 MB state of a 128^3 cavity, against 3-8 ms for its upload)
extern "C" int tfb_host_checksum(const double* p, int64_t n, uint64_t out[2]) {
    TFB_CHECK(p && n >= 0 && out, "bad arguments");
    const uint64_t* w = reinterpret_cast<const uint64_t*>(p);
    const int nt = n < (1 << 16) ? 1 : (int)std::min<int64_t>(16, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> sum(nt, 0), x(nt, 0);
    auto work = [&](int t) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, xx = 0;
        int64_t i = a;
        for (; i + 4 <= b; i += 4) { s0 += w[i]; s1 += w[i + 1] * 3u; s2 += w[i + 2] * 5u; s3 += w[i + 3] * 7u; xx ^= w[i] ^ w[i + 1] ^ w[i + 2] ^ w[i + 3]; }
        for (; i < b; i++) { s0 += w[i]; xx ^= w[i]; }
        sum[t] = s0 + s1 + s2 + s3; x[t] = xx;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    out[0] = out[1] = 0;
    for (int t = 0; t < nt; t++) { out[0] += sum[t] * (uint64_t)(2 * t + 1); out[1] ^= x[t]; }
    return 0;
}

extern "C" int tfb_assemble_resident(tfb_ctx* c, tfb_mat* m, int do_j, int do_f) {
    TFB_CHECK(c && c->have_params, "tfb_set_params has not been called");
    TFB_CHECK(do_j || do_f, "nothing to do");
    TFB_CHECK(!do_j || (m && m->ctx == c), "matrix does not belong to this context");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (do_j) { m->version = tfb_next_version(); m->shift = 0.0; }
    if (c->nranks > 1) {
        const bool marching = c->desc.dim == 3 && c->desc.nz > 1;
        if (marching && c->nzl >= 4 && tfb_overlap_enabled(0)) {
            // the interior planes do not read the halo: exchange it on a side stream while they are assembled, then the
            // two planes next to it (one launch of the plane-tile kernel).  OPT-IN, see tfb_overlap_enabled (tfb_comm.cu)
            if (tfb_comm_stream(c)) return -1;
            TFB_CUDA(cudaEventRecord(c->ev_comm[0], c->stream));
            TFB_CUDA(cudaStreamWaitEvent(c->s_comm, c->ev_comm[0], 0));
            if (tfb_halo_exchange_on(c, c->d_state, c->s_comm)) return -1;
            TFB_CUDA(cudaEventRecord(c->ev_comm[1], c->s_comm));
            c->win0 = 1; c->win1 = c->nzl - 1;
            int rc = dispatch_assemble(c, m, do_j, do_f);
            TFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_comm[1], 0));
            c->win0 = 0; c->win1 = -2;      // both end planes, one launch
            if (!rc) rc = dispatch_assemble(c, m, do_j, do_f);
            c->win0 = 0; c->win1 = -1;
            return rc;
        }
        int rc = tfb_halo_exchange(c, c->d_state);
        if (rc) return rc;
    }
    return dispatch_assemble(c, m, do_j, do_f);
}

extern "C" int tfb_rhs_download(tfb_ctx* c, double* out) {
    TFB_CHECK(c && out, "null argument");
    TFB_CUDA(cudaMemcpyAsync(out, c->d_rhs, sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int tfb_sync(tfb_ctx* c) {
    TFB_CHECK(c, "null ctx");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int tfb_rhs(tfb_ctx* c, const double* state, double* out) {
    int rc = tfb_state_upload(c, state);
    if (rc) return rc;
    rc = tfb_assemble_resident(c, nullptr, 0, 1);
    if (rc) return rc;
    return tfb_rhs_download(c, out);
}

static int dispatch_assemble_impl(tfb_ctx* c, tfb_mat* m, int do_j, int do_f) {
#define X(C) if (c->desc.config == C::ID) return launch_assemble<C>(c, m, do_j != 0, do_f != 0);
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return tfb_fail(__FILE__, __LINE__, "dispatch_assemble", "unknown config");
}
static int dispatch_assemble(tfb_ctx* c, tfb_mat* m, int do_j, int do_f) { return dispatch_assemble_impl(c, m, do_j, do_f); }

// Pieces of the pipelined host path, as plane boundaries b[0] = 0 < b[1] < ... = nzl.  Every piece costs a fixed ~40 us
// (copy set-up on both engines, event hand-overs), the first upload and the last download are not overlapped with
// anything, so the pieces are small at both ends and large in the middle: 4, 12, 16, 32, ..., 32, 16, 12, 4 planes
// (measured at 128^3, uniform pieces: 2 planes 2.40 ms, 4: 2.21, 8: 1.95, 16: 1.83, 32: 1.83, 64: 2.08 per call).
// TFB_PIPE_PLANES=n forces uniform pieces of n planes.
static std::vector<int> tfb_pipe_pieces(int nzl) {
    static int uniform = -1;
    if (uniform < 0) { const char* e = getenv("TFB_PIPE_PLANES"); uniform = e ? std::max(1, atoi(e)) : 0; }
    std::vector<int> b{0};
    if (const char* e = getenv("TFB_PIPE_PIECES")) {          // explicit piece sizes "4,12,16,...": tuning only
        int k = 0;
        for (const char* q = e; *q && k < nzl;) {
            k = std::min(nzl, k + std::max(1, atoi(q)));
            b.push_back(k);
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        if (b.back() != nzl) b.push_back(nzl);
        return b;
    }
    if (uniform > 0 || nzl < 64) {
        const int step = uniform > 0 ? uniform : 8;
        for (int k = step; k < nzl; k += step) b.push_back(k);
        b.push_back(nzl);
        return b;
    }
    const int ramp[3] = {4, 12, 16};
    std::vector<int> tail;
    int lo = 0, hi = nzl;
    for (int r = 0; r < 3 && hi - lo >= 2 * ramp[r]; r++) {
        lo += ramp[r]; b.push_back(lo);
        hi -= ramp[r]; tail.push_back(hi);
    }
    while (hi - lo > 32) { lo += 32; b.push_back(lo); }
    if (hi > lo) b.push_back(hi);
    for (int t = (int)tail.size() - 2; t >= 0; t--) b.push_back(tail[t]);
    b.push_back(nzl);
    b.erase(std::unique(b.begin(), b.end()), b.end());
    return b;
}

// the piece boundaries for a slab of nzl planes (host-only; tests/test_cabi_cpu.py checks the grading without a device)
extern "C" int tfb_pipe_pieces_of(int nzl, int* out, int cap) {
    if (nzl < 1 || !out || cap < 2) return -1;
    const std::vector<int> b = tfb_pipe_pieces(nzl);
    if ((int)b.size() > cap) return -1;
    for (size_t i = 0; i < b.size(); i++) out[i] = b[i];
    return (int)b.size();
}

// Host-buffer path for 3-D grids (one GPU or a z-slab): the upload of the state, the assembly and the download
// of F(x) are pipelined over z-pieces on three streams (PCIe is full duplex), so the call costs
// about max(H2D, D2H) instead of H2D + kernel + D2H.
static int jacobian_pipelined(tfb_ctx* c, const double* state, tfb_mat* m, double* rhs_out) {
    const std::vector<int> b = tfb_pipe_pieces(c->nzl);
    const int nch = (int)b.size() - 1;
    if (!c->s_h2d) {
        TFB_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        TFB_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        for (int e = 0; e < TFB_MAX_CHUNKS; e++) {
            TFB_CUDA(cudaEventCreateWithFlags(&c->ev_up[e], cudaEventDisableTiming));
            TFB_CUDA(cudaEventCreateWithFlags(&c->ev_k[e], cudaEventDisableTiming));
        }
    }
    // order against earlier work on the compute stream (e.g. a solve still reading the old values)
    TFB_CUDA(cudaEventRecord(c->ev_k[TFB_MAX_CHUNKS - 1], c->stream));
    TFB_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_k[TFB_MAX_CHUNKS - 1], 0));
    const size_t pr = (size_t)c->plane_rows;
    int lo = 0, hi = c->nzl;      // planes the pieces still have to upload
    if (c->nranks > 1) {
        // z-slab: the first and the last owned plane go up first and are exchanged with the neighbours on the compute
        // stream (NCCL) while the bulk of the slab is still on its way
        TFB_CUDA(cudaMemcpyAsync(c->d_state + pr, state, sizeof(double) * pr, cudaMemcpyHostToDevice, c->s_h2d));
        TFB_CUDA(cudaMemcpyAsync(c->d_state + pr * c->nzl, state + pr * (c->nzl - 1), sizeof(double) * pr,
                                 cudaMemcpyHostToDevice, c->s_h2d));
        TFB_CUDA(cudaEventRecord(c->ev_up[TFB_MAX_CHUNKS - 2], c->s_h2d));
        TFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up[TFB_MAX_CHUNKS - 2], 0));
        if (tfb_halo_exchange(c, c->d_state)) return -1;
        lo = 1; hi = c->nzl - 1;
    }
    // upload pieces are shifted by one plane: piece ch ends with the first plane of piece ch+1, which is the
    // only plane of the next piece the kernel of piece ch reads
    for (int ch = 0; ch < nch; ch++) {
        const int u0 = std::max(lo, ch == 0 ? 0 : b[ch] + 1), u1 = std::min(hi, std::min(b[ch + 1] + 1, c->nzl));
        if (u1 > u0)
            TFB_CUDA(cudaMemcpyAsync(c->d_state + pr * (u0 + 1), state + pr * u0, sizeof(double) * pr * (u1 - u0),
                                     cudaMemcpyHostToDevice, c->s_h2d));
        TFB_CUDA(cudaEventRecord(c->ev_up[ch], c->s_h2d));
    }
    m->version = tfb_next_version();
    m->shift = 0.0;
    c->state_uploads++;
    for (int ch = 0; ch < nch; ch++) {
        const int p0 = b[ch], p1 = b[ch + 1];
        TFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_up[ch], 0));
        c->win0 = p0; c->win1 = p1;                 // plane window of this launch (marching chunks inside it)
        int rc = dispatch_assemble(c, m, 1, rhs_out != nullptr);
        c->win0 = 0; c->win1 = -1;
        if (rc) return rc;
        if (rhs_out) {
            TFB_CUDA(cudaEventRecord(c->ev_k[ch], c->stream));
            TFB_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_k[ch], 0));
            TFB_CUDA(cudaMemcpyAsync(rhs_out + pr * p0, c->d_rhs + pr * p0, sizeof(double) * pr * (p1 - p0),
                                     cudaMemcpyDeviceToHost, c->s_d2h));
        }
    }
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    if (rhs_out) TFB_CUDA(cudaStreamSynchronize(c->s_d2h));
    return 0;
}

extern "C" int tfb_jacobian(tfb_ctx* c, const double* state, tfb_mat* m, double* rhs_out) {
    TFB_CHECK(c && state && m && m->ctx == c, "bad arguments");
    TFB_CHECK(c->have_params, "tfb_set_params has not been called");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    const int nch = c->desc.nz > 1 && c->desc.dim == 3 && c->nzl >= 4 ? (int)tfb_pipe_pieces(c->nzl).size() - 1 : 0;
    if (nch >= 2 && nch < TFB_MAX_CHUNKS - 1 && !getenv("TFB_NO_PIPELINE"))
        return jacobian_pipelined(c, state, m, rhs_out);
    int rc = tfb_state_upload(c, state);
    if (rc) return rc;
    rc = tfb_assemble_resident(c, m, 1, rhs_out != nullptr);
    if (rc) return rc;
    if (rhs_out) return tfb_rhs_download(c, rhs_out);
    return tfb_sync(c);
}

extern "C" int tfb_mass_diag(tfb_ctx* c, double* diag_out) {
    TFB_CHECK(c && diag_out, "null argument");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (!c->d_massdiag) TFB_CUDA(cudaMalloc(&c->d_massdiag, sizeof(double) * c->n_local));   // kept: callers ask repeatedly
    double* d = c->d_massdiag;
    const int bs = 256;
    tfb_mass_kernel<<<(unsigned)((c->n_local + bs - 1) / bs), bs, 0, c->stream>>>(c->grid(), c->desc.k0, c->n_local, d);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    TFB_CUDA(cudaMemcpyAsync(diag_out, d, sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------- timing helpers -------------------------------
extern "C" int tfb_event_record(tfb_ctx* c, int slot) {
    TFB_CHECK(c && slot >= 0 && slot < TFB_EVENT_SLOTS, "bad slot");
    if (!c->ev[slot]) TFB_CUDA(cudaEventCreate(&c->ev[slot]));
    TFB_CUDA(cudaEventRecord(c->ev[slot], c->stream));
    return 0;
}

extern "C" int tfb_event_elapsed_ms(tfb_ctx* c, int a, int b, float* ms) {
    TFB_CHECK(c && ms && a >= 0 && a < TFB_EVENT_SLOTS && b >= 0 && b < TFB_EVENT_SLOTS && c->ev[a] && c->ev[b], "bad slot");
    TFB_CUDA(cudaEventSynchronize(c->ev[b]));
    TFB_CUDA(cudaEventElapsedTime(ms, c->ev[a], c->ev[b]));
    return 0;
}

__global__ void tfb_flush_kernel(double4* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_double4(1.0, 2.0, 3.0, 4.0);
}

extern "C" int tfb_flush_l2(tfb_ctx* c) {
    TFB_CHECK(c, "null ctx");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    if (!c->d_flush) {
        c->flush_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
        TFB_CUDA(cudaMalloc(&c->d_flush, c->flush_bytes));
    }
    tfb_flush_kernel<<<148 * 8, 256, 0, c->stream>>>((double4*)c->d_flush, c->flush_bytes / sizeof(double4));
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfb_pinned_alloc(size_t bytes, void** out) {
    TFB_CHECK(out, "null argument");
    TFB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return 0;
}

extern "C" int tfb_pinned_free(void* p) {
    if (p) TFB_CUDA(cudaFreeHost(p));
    return 0;
}

// ------------------------------- structured SpMV (tfb_spmv_march.cuh) -------------------------------
template <class Cfg, int TJ>
static int launch_spmv_march_t(tfb_ctx* c, const tfb_mat* m, const double* x_global_base, int kvalid0, int kvalid1, double* y,
                               int prow, unsigned rowmask, unsigned colmask, const double* rowscale) {
    // TJ: lines per CTA.  Two by default (dof 5 with three measured slower: 514 us against 441 us at 128^3)
    TfbSpmvArgs a;
    a.g = c->grid();
    a.x = x_global_base;
    a.kvalid0 = kvalid0; a.kvalid1 = kvalid1;
    a.row_ptr = c->d_row_ptr;
    a.vals = m->d_vals;
    a.y = y;
    a.rowscale = rowscale;
    a.k0 = c->desc.k0; a.nzl = c->nzl;
    a.pvar = -1; a.prow_cell_i = a.prow_cell_j = a.prow_cell_k = -1;
    if (prow >= 0) {
        const long long cell = prow / c->desc.dof;
        a.pvar = prow % c->desc.dof;
        a.prow_cell_i = (int)(cell % c->desc.nx);
        a.prow_cell_j = (int)((cell / c->desc.nx) % c->desc.ny);
        a.prow_cell_k = (int)(cell / ((long long)c->desc.nx * c->desc.ny));
    }
    a.rowmask = rowmask; a.colmask = colmask;
    if (c->plane_nnz < 0) {
        // non-zeros of one plane away from the z walls (their CSR layouts are shifted copies of each other): difference
        // of the first offsets of two consecutive planes without wall flags, if the slab holds such a pair
        c->plane_nnz = 0;
        const int nz = c->desc.nz, kfar2 = tfb_far2_index(nz);
        for (int kl = 0; kl + 1 < c->nzl; kl++) {
            const int k = c->desc.k0 + kl;
            auto flagged = [&](int kk) { return kk == 0 || kk == nz - 1 || kk == kfar2; };
            if (flagged(k) || flagged(k + 1)) continue;
            int two[2] = {0, 0};
            TFB_CUDA(cudaMemcpy(&two[0], c->d_row_ptr + (size_t)kl * c->plane_rows, sizeof(int), cudaMemcpyDeviceToHost));
            TFB_CUDA(cudaMemcpy(&two[1], c->d_row_ptr + (size_t)(kl + 1) * c->plane_rows, sizeof(int), cudaMemcpyDeviceToHost));
            c->plane_nnz = two[1] - two[0];
            break;
        }
    }
    a.plane_nnz = c->plane_nnz;
    a.kofs0 = c->win1 >= 0 ? c->win0 : 0;
    a.klim = c->win1 >= 0 ? c->win1 : c->nzl;
    size_t smem = sizeof(double) * TfbMarch<Cfg, TJ>::smem_doubles(true);
    // the attribute is per device: set it when this instantiation meets a device for the first time
    static unsigned configured_devices = 0;
    if (!((configured_devices >> (c->desc.device & 31)) & 1u)) {
        TFB_CUDA(cudaFuncSetAttribute(tfb_spmv_march_kernel<Cfg, TJ, TFB_KCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TFB_CUDA(cudaFuncSetAttribute(tfb_spmv_march_kernel<Cfg, TJ, TFB_KCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_devices |= 1u << (c->desc.device & 31);
    }
    dim3 block(32, Cfg::DOF, TJ);
    dim3 grid((c->desc.nx + TFB_TI - 1) / TFB_TI, (c->desc.ny + TJ - 1) / TJ, (a.klim - a.kofs0 + TFB_KCH - 1) / TFB_KCH);
    if (rowmask || colmask) tfb_spmv_march_kernel<Cfg, TJ, TFB_KCH, true><<<grid, block, smem, c->stream>>>(a);
    else tfb_spmv_march_kernel<Cfg, TJ, TFB_KCH, false><<<grid, block, smem, c->stream>>>(a);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    return 0;
}

template <class Cfg>
static int launch_spmv_march(tfb_ctx* c, const tfb_mat* m, const double* x_global_base, int kvalid0, int kvalid1, double* y,
                             int prow, unsigned rowmask, unsigned colmask, const double* rowscale) {
#ifdef TFB_ASM_EXPERIMENTS
    static int tj = -1;
    if (tj < 0) { const char* e = getenv("TFB_SPMV_TJ"); tj = e ? atoi(e) : 2; }
    if constexpr (Cfg::DOF == 4) {
        if (tj == 3) return launch_spmv_march_t<Cfg, 3>(c, m, x_global_base, kvalid0, kvalid1, y, prow, rowmask, colmask, rowscale);
        if (tj == 1) return launch_spmv_march_t<Cfg, 1>(c, m, x_global_base, kvalid0, kvalid1, y, prow, rowmask, colmask, rowscale);
    }
#endif
    return launch_spmv_march_t<Cfg, 2>(c, m, x_global_base, kvalid0, kvalid1, y, prow, rowmask, colmask, rowscale);
}

// returns 1 when the configuration has no structured kernel (2-D / folded grids): caller uses the CSR kernel
int tfb_spmv_structured(tfb_ctx* c, const tfb_mat* m, const double* x_global_base, int kvalid0, int kvalid1, double* y,
                        int prow, unsigned rowmask, unsigned colmask, const double* rowscale) {
#define X(C) if (c->desc.config == C::ID) { if constexpr (C::FLAT) return 1; else return launch_spmv_march<C>(c, m, x_global_base, kvalid0, kvalid1, y, prow, rowmask, colmask, rowscale); }
    TFB_FOR_EACH_CONFIG(X)
#undef X
    return 1;
}

// dst = src + alpha * diag(d) on the fixed pattern (device-side matrix arithmetic for time stepping)
__global__ void tfb_add_diag_kernel(long long nrows, long long row0, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                    const double* __restrict__ src, double* __restrict__ dst, double alpha,
                                    const double* __restrict__ d, int* __restrict__ missing) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const int e0 = row_ptr[row], e1 = row_ptr[row + 1];
    const int gcol = (int)(row + row0);
    bool found = false;
    for (int e = e0; e < e1; e++) {
        double v = src[e];
        if (col[e] == gcol) { v += alpha * d[row]; found = true; }
        dst[e] = v;
    }
    if (!found && d[row] != 0.0) atomicAdd(missing, 1);
}

extern "C" int tfb_mat_add_diag(tfb_mat* dst, const tfb_mat* src, double alpha, const double* d) {
    TFB_CHECK(dst && src && d && dst->ctx == src->ctx, "bad arguments");
    tfb_ctx* c = dst->ctx;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    double* dd = nullptr;
    int* dmiss = nullptr;
    TFB_CUDA(cudaMalloc(&dd, sizeof(double) * c->n_local));
    TFB_CUDA(cudaMalloc(&dmiss, sizeof(int)));
    TFB_CUDA(cudaMemsetAsync(dmiss, 0, sizeof(int), c->stream));
    TFB_CUDA(cudaMemcpyAsync(dd, d, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    tfb_add_diag_kernel<<<(unsigned)((c->n_local + 255) / 256), 256, 0, c->stream>>>(c->n_local, c->row0, c->d_row_ptr, c->d_col,
                                                                                   src->d_vals, dst->d_vals, alpha, dd, dmiss);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    int miss = 0;
    TFB_CUDA(cudaMemcpyAsync(&miss, dmiss, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(dd);
    cudaFree(dmiss);
    dst->version = tfb_next_version();
    dst->shift = src->shift;
    TFB_CHECK(miss == 0, "a row with a non-zero diagonal update has no structural diagonal");
    return 0;
}

extern "C" int tfb_mat_set_shift(tfb_mat* m, double shift) {
    TFB_CHECK(m, "null argument");
    m->shift = shift;
    return 0;
}
