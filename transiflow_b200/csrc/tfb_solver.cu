// Krylov solver of the Newton step (placeholder until the preconditioned FGMRES lands).
#include "tfb_internal.h"

struct tfb_solver_state { int unused; };
void tfb_solver_free(tfb_solver_state* s) { delete s; }

__global__ void tfb_spmv_kernel(long long nrows, const int* __restrict__ row_ptr, const int* __restrict__ col,
                                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double s = 0.0;
    for (int e = row_ptr[row]; e < row_ptr[row + 1]; e++) s += vals[e] * x[col[e]];
    y[row] = s;
}

extern "C" int tfb_spmv(tfb_mat* m, const double* x, double* y) {
    TFB_CHECK(m && x && y, "null argument");
    tfb_ctx* c = m->ctx;
    TFB_CHECK(c->nranks == 1, "host-vector spmv is single-GPU only");
    TFB_CUDA(cudaSetDevice(c->desc.device));
    double *dx = nullptr, *dy = nullptr;
    TFB_CUDA(cudaMalloc(&dx, sizeof(double) * c->n_local));
    TFB_CUDA(cudaMalloc(&dy, sizeof(double) * c->n_local));
    TFB_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * c->n_local, cudaMemcpyHostToDevice, c->stream));
    const int bs = 256;
    tfb_spmv_kernel<<<(unsigned)((c->n_local + bs - 1) / bs), bs, 0, c->stream>>>(c->n_local, c->d_row_ptr, c->d_col, m->d_vals, dx, dy);
    TFB_LAUNCHED();
    TFB_CUDA(cudaGetLastError());
    TFB_CUDA(cudaMemcpyAsync(y, dy, sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(dx);
    cudaFree(dy);
    return 0;
}

extern "C" int tfb_solve(tfb_mat*, const double*, double*, const tfb_solve_opts*, tfb_solve_info*) {
    return tfb_fail(__FILE__, __LINE__, "tfb_solve", "not implemented yet");
}
