// Internal definitions of libtfb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/tfb200.h"
#include "tfb_cell.h"

extern thread_local std::string g_tfb_err;
extern int64_t g_tfb_launches;

int tfb_fail(const char* file, int line, const char* what, const char* detail);

#define TFB_CUDA(call)                                                                 \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) return tfb_fail(__FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define TFB_CHECK(cond, msg)                                                           \
    do {                                                                               \
        if (!(cond)) return tfb_fail(__FILE__, __LINE__, #cond, msg);                  \
    } while (0)
#define TFB_LAUNCHED() (g_tfb_launches++)

#define TFB_MAX_RANKS 16
#define TFB_MAX_CHUNKS 80
struct tfb_solver_state;  // tfb_solver.cu

#define TFB_EVENT_SLOTS 1040   // 16 general-purpose timers + 512 (start, stop) pairs for un-synchronised timing loops
struct tfb_ctx {
    tfb_desc desc;
    int nzl;                    // owned planes
    int64_t plane_rows;         // nx*ny*dof
    int64_t n_local, n_global, row0;
    int64_t nnz = 0;
    int64_t state_uploads = 0;  // host -> device copies of the state (tfb_upload_count)
    int plane_nnz = -1;         // structural non-zeros of a plane away from the z walls (-1: not determined yet)
    cudaStream_t stream = nullptr;
    // geometry on the device
    double* d_met[3] = {nullptr, nullptr, nullptr};
    double* d_cor = nullptr;
    double* d_fval[TFB_MAX_FORCE] = {};
    int8_t fdir[TFB_MAX_FORCE] = {};
    double* d_frc_static = nullptr;
    bool has_frc_static = false;
    TfbParams prm;
    bool have_params = false;
    // state with one ghost plane below and above: (nzl+2) planes
    double* d_state = nullptr;
    double* d_rhs = nullptr;
    // fixed pattern
    int* d_row_ptr = nullptr;   // n_local+1
    int* d_col = nullptr;       // nnz, GLOBAL columns
    bool have_pattern = false;
    // scratch for L2 flush
    double* d_massdiag = nullptr;   // tfb_mass_diag result buffer
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
    cudaEvent_t ev[TFB_EVENT_SLOTS] = {};
    // pipelined host path of tfb_jacobian: copy streams, per-chunk events, z-chunk window of a launch
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up[TFB_MAX_CHUNKS] = {}, ev_k[TFB_MAX_CHUNKS] = {};
    int win0 = 0, win1 = -1;    // plane window [win0, win1) of the next assembly launch (win1 < 0: the whole slab)
    // z-slabs: the halo exchange runs on its own stream next to the interior planes of the kernel that needs it
    void* direct_pool = nullptr;   // parked work space of the 2-D direct solve (tfb_direct.cu)
    cudaStream_t s_comm = nullptr;
    cudaEvent_t ev_comm[2] = {nullptr, nullptr};
    tfb_solver_state* solver = nullptr;   // FDM operators, Krylov work space (tfb_solver.cu)
    // value buffers of destroyed matrices, kept for the next tfb_mat_create: a Newton loop makes one Jacobian per
    // step and cudaMalloc/cudaFree of ~1 GB next to a 60 GB Krylov basis cost up to 0.6 s per call (measured)
    std::vector<double*> vals_pool;
    // multi-GPU
    int nranks = 1, rank = 0;
    int slab_k0[TFB_MAX_RANKS + 1] = {};   // first plane of every rank's slab (after tfb_comm_init)
    void* nccl_comm = nullptr;
    TfbGrid grid() const;
};

struct tfb_mat {
    tfb_ctx* ctx;
    double* d_vals = nullptr;
    uint64_t version = 0;       // bumped whenever values change (invalidates the preconditioner)
    double shift = 0.0;         // the matrix is (an assembled Jacobian) + shift * (mass matrix): tfb_mat_set_shift
    void* direct = nullptr;     // cached factors of the 2-D direct solve (tfb_direct.cu), keyed on `version`
};

uint64_t tfb_next_version();
int tfb_build_pattern(tfb_ctx* ctx);
int tfb_spmv_structured(tfb_ctx* c, const tfb_mat* m, const double* x_global_base, int kvalid0, int kvalid1, double* y,
                        int prow, unsigned rowmask, unsigned colmask, const double* rowscale);
int tfb_halo_exchange(tfb_ctx* ctx, double* d_vec_with_ghosts);
int tfb_halo_exchange_on(tfb_ctx* ctx, double* d_vec_with_ghosts, cudaStream_t stream);
int tfb_comm_stream(tfb_ctx* c);
// TFB_OVERLAP: "1" = the halo exchange of the assembly (which = 0) and of the operator products (which = 1) runs on a side
// stream next to the interior planes; "asm" / "spmv" select one of the two
bool tfb_overlap_enabled(int which);
int tfb_allreduce_sum(tfb_ctx* c, double* d_buf, int count);
int tfb_alltoallv(tfb_ctx* c, const double* send, const long long* scount, const long long* sdispl,
                  double* recv, const long long* rcount, const long long* rdispl);
int tfb_halo_up_f32(tfb_ctx* c, const float* first_plane, float* ghost_above, size_t count);
int tfb_halo_up_f64(tfb_ctx* c, const double* first_plane, double* ghost_above, size_t count);
int tfb_allgather_f32(tfb_ctx* c, const float* send, float* recv, size_t count);
void tfb_solver_free(tfb_solver_state* s);
void tfb_direct_free(tfb_mat* mat);
void tfb_direct_pool_free(tfb_ctx* c);
