// NCCL plumbing for z-slab partitioned runs (one process per GPU).  libnccl is loaded with
// dlopen so that single-GPU use has no NCCL dependency at all.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "tfb_internal.h"

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_t_;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

static struct {
    void* handle;
    int (*GetUniqueId)(ncclUniqueId_t*);
    int (*CommInitRank)(ncclComm_t_*, int, ncclUniqueId_t, int);
    int (*CommDestroy)(ncclComm_t_);
    int (*Send)(const void*, size_t, int, int, ncclComm_t_, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, ncclComm_t_, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t);
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t_, cudaStream_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char* (*GetErrorString)(int);
} nccl;

static int load_nccl() {
    if (nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (nccl.handle) break;
    }
    if (!nccl.handle) return tfb_fail(__FILE__, __LINE__, "dlopen(libnccl.so.2)", dlerror());
#define SYM(f) *(void**)(&nccl.f) = dlsym(nccl.handle, "nccl" #f); if (!nccl.f) return tfb_fail(__FILE__, __LINE__, "dlsym", "nccl" #f);
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(AllGather) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    return 0;
}

// TFB_TRACE=1: every NCCL operation is announced on stderr when it is enqueued (which stream, how much); TFB_TRACE=2
// also waits for it.  For chasing ordering problems between ranks; off by default.
static int trace_level() {
    static int lvl = -1;
    if (lvl < 0) { const char* e = getenv("TFB_TRACE"); lvl = e ? atoi(e) : 0; }
    return lvl;
}
static long g_trace_seq = 0;
#define TFB_TRACE_OP(ctx_, what, count, str_)                                                                  \
    do {                                                                                                       \
        if (trace_level() > 0) {                                                                               \
            fprintf(stderr, "[tfb r%d] #%ld %s count %lld on %s\n", (ctx_)->rank, g_trace_seq++, what,          \
                    (long long)(count), (str_) == (ctx_)->stream ? "main" : "side");                           \
            fflush(stderr);                                                                                    \
        }                                                                                                      \
    } while (0)
#define TFB_TRACE_DONE(ctx_, str_)                                                                             \
    do {                                                                                                       \
        if (trace_level() > 1) {                                                                               \
            TFB_CUDA(cudaStreamSynchronize(str_));                                                             \
            fprintf(stderr, "[tfb r%d] #%ld done\n", (ctx_)->rank, g_trace_seq - 1);                           \
            fflush(stderr);                                                                                    \
        }                                                                                                      \
    } while (0)

#define TFB_NCCL(call)                                                                     \
    do {                                                                                   \
        int r_ = (call);                                                                   \
        if (r_ != 0) return tfb_fail(__FILE__, __LINE__, #call, nccl.GetErrorString(r_));  \
    } while (0)

extern "C" int tfb_nccl_unique_id(uint8_t id[128]) {
    if (load_nccl()) return -1;
    ncclUniqueId_t u;
    TFB_NCCL(nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 0;
}

extern "C" int tfb_comm_init(tfb_ctx* c, int nranks, int rank, const uint8_t id[128]) {
    TFB_CHECK(c && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
    c->nranks = nranks;
    c->rank = rank;
    if (nranks == 1) return 0;
    if (load_nccl()) return -1;
    TFB_CUDA(cudaSetDevice(c->desc.device));
    ncclUniqueId_t u;
    memcpy(u.internal, id, 128);
    ncclComm_t_ comm;
    TFB_NCCL(nccl.CommInitRank(&comm, nranks, u, rank));
    c->nccl_comm = comm;
    // every rank learns every slab: k0 of rank r in slab_k0[r], slab_k0[nranks] = nz
    TFB_CHECK(nranks <= TFB_MAX_RANKS, "too many ranks");
    double* d = nullptr;
    TFB_CUDA(cudaMalloc(&d, sizeof(double) * nranks));
    std::vector<double> h(nranks, 0.0);
    h[rank] = (double)c->desc.k0;
    TFB_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(double) * nranks, cudaMemcpyHostToDevice, c->stream));
    TFB_NCCL(nccl.AllReduce(d, d, (size_t)nranks, NCCL_FLOAT64, NCCL_SUM, comm, c->stream));
    TFB_CUDA(cudaMemcpyAsync(h.data(), d, sizeof(double) * nranks, cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    for (int r = 0; r < nranks; r++) c->slab_k0[r] = (int)h[r];
    c->slab_k0[nranks] = c->desc.nz;
    for (int r = 0; r < nranks; r++)
        TFB_CHECK(c->slab_k0[r] < c->slab_k0[r + 1], "z-slabs must be ordered by rank and non-empty");
    return 0;
}

// personalised all-to-all of doubles (counts / displacements in elements), own block by memcpy
int tfb_alltoallv_bytes(tfb_ctx* c, const void* send, const long long* scount, const long long* sdispl, void* recv,
                        const long long* rcount, const long long* rdispl, int elem_bytes) {
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    TFB_CHECK(elem_bytes == 8 || elem_bytes == 4, "element size");
    ncclComm_t_ comm = (ncclComm_t_)c->nccl_comm;
    const int dt = elem_bytes == 8 ? NCCL_FLOAT64 : NCCL_FLOAT32;
    const char* sp = (const char*)send;
    char* rp = (char*)recv;
    TFB_TRACE_OP(c, "alltoallv", scount[c->rank], c->stream);
    TFB_NCCL(nccl.GroupStart());
    for (int r = 0; r < c->nranks; r++) {
        if (r == c->rank) continue;
        if (scount[r] > 0) TFB_NCCL(nccl.Send(sp + sdispl[r] * elem_bytes, (size_t)scount[r], dt, r, comm, c->stream));
        if (rcount[r] > 0) TFB_NCCL(nccl.Recv(rp + rdispl[r] * elem_bytes, (size_t)rcount[r], dt, r, comm, c->stream));
    }
    TFB_NCCL(nccl.GroupEnd());
    TFB_CUDA(cudaMemcpyAsync(rp + rdispl[c->rank] * elem_bytes, sp + sdispl[c->rank] * elem_bytes,
                             (size_t)elem_bytes * scount[c->rank], cudaMemcpyDeviceToDevice, c->stream));
    TFB_LAUNCHED();
    return 0;
}

int tfb_alltoallv(tfb_ctx* c, const double* send, const long long* scount, const long long* sdispl,
                  double* recv, const long long* rcount, const long long* rdispl) {
    return tfb_alltoallv_bytes(c, send, scount, sdispl, recv, rcount, rdispl, 8);
}

// One-layer halo exchange of a slab vector stored with ghost planes:
// [ghost below | owned planes | ghost above], plane = nx*ny*dof doubles.
// Rank r owns planes directly above rank r-1 (z-slab order = rank order).
int tfb_halo_exchange(tfb_ctx* c, double* v) { return tfb_halo_exchange_on(c, v, c->stream); }

// OPT-IN ONLY.  Measured on 2 x B200 at 128^3 per rank (round 2): the side stream buys nothing -- the exchange kernel finds
// no free SM while the interior-plane kernel keeps refilling them and runs in its tail (assembly 0.249 ms with, 0.250 ms
// without, 0.231 ms on one GPU) -- and bench.py hung intermittently with it (3 of 11 runs, with and without a
// high-priority side stream, never with the trace of TFB_TRACE on), so the default exchanges the halo in front of the kernel.
bool tfb_overlap_enabled(int which) {
    static int mode = -1;     // bit 0: assembly, bit 1: operator products
    if (mode < 0) {
        const char* e = getenv("TFB_OVERLAP");
        mode = !e ? 0 : !strcmp(e, "asm") ? 1 : !strcmp(e, "spmv") ? 2 : !strcmp(e, "0") ? 0 : 3;
    }
    return (mode >> which) & 1;
}

// side stream + events for exchanges that run next to the interior part of a kernel
int tfb_comm_stream(tfb_ctx* c) {
    if (c->s_comm) return 0;
    TFB_CUDA(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
    for (auto& e : c->ev_comm) TFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}

int tfb_halo_exchange_on(tfb_ctx* c, double* v, cudaStream_t stream) {
    if (c->nranks <= 1) return 0;
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    ncclComm_t_ comm = (ncclComm_t_)c->nccl_comm;
    const size_t pl = (size_t)c->plane_rows;
    double* first_owned = v + pl;
    double* last_owned = v + pl * c->nzl;
    double* ghost_lo = v;
    double* ghost_hi = v + pl * (c->nzl + 1);
    TFB_TRACE_OP(c, "halo exchange", pl, stream);
    TFB_NCCL(nccl.GroupStart());
    if (c->rank > 0) {
        TFB_NCCL(nccl.Send(first_owned, pl, NCCL_FLOAT64, c->rank - 1, comm, stream));
        TFB_NCCL(nccl.Recv(ghost_lo, pl, NCCL_FLOAT64, c->rank - 1, comm, stream));
    }
    if (c->rank < c->nranks - 1) {
        TFB_NCCL(nccl.Send(last_owned, pl, NCCL_FLOAT64, c->rank + 1, comm, stream));
        TFB_NCCL(nccl.Recv(ghost_hi, pl, NCCL_FLOAT64, c->rank + 1, comm, stream));
    }
    TFB_NCCL(nccl.GroupEnd());
    TFB_LAUNCHED();
    TFB_TRACE_DONE(c, stream);
    return 0;
}

int tfb_allreduce_sum(tfb_ctx* c, double* d_buf, int count) {
    if (c->nranks <= 1) return 0;
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    TFB_TRACE_OP(c, "allreduce", count, c->stream);
    TFB_NCCL(nccl.AllReduce(d_buf, d_buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (ncclComm_t_)c->nccl_comm, c->stream));
    TFB_LAUNCHED();
    TFB_TRACE_DONE(c, c->stream);
    return 0;
}

// recv[r * count .. (r+1) * count) = send of rank r  (fp32 elements)
int tfb_allgather_f32(tfb_ctx* c, const float* send, float* recv, size_t count) {
    if (c->nranks <= 1) {
        TFB_CUDA(cudaMemcpyAsync(recv, send, sizeof(float) * count, cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    TFB_TRACE_OP(c, "allgather", count, c->stream);
    TFB_NCCL(nccl.AllGather(send, recv, count, NCCL_FLOAT32, (ncclComm_t_)c->nccl_comm, c->stream));
    TFB_LAUNCHED();
    TFB_TRACE_DONE(c, c->stream);
    return 0;
}

// one-directional halo of an fp32 plane: every rank sends `count` floats at `first_plane` to the rank below and receives
// the first plane of the rank above into `ghost_above`
int tfb_halo_up_f32(tfb_ctx* c, const float* first_plane, float* ghost_above, size_t count) {
    if (c->nranks <= 1) return 0;
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    ncclComm_t_ comm = (ncclComm_t_)c->nccl_comm;
    TFB_TRACE_OP(c, "halo up f32", count, c->stream);
    TFB_NCCL(nccl.GroupStart());
    if (c->rank > 0) TFB_NCCL(nccl.Send(first_plane, count, NCCL_FLOAT32, c->rank - 1, comm, c->stream));
    if (c->rank < c->nranks - 1) TFB_NCCL(nccl.Recv(ghost_above, count, NCCL_FLOAT32, c->rank + 1, comm, c->stream));
    TFB_NCCL(nccl.GroupEnd());
    TFB_LAUNCHED();
    return 0;
}

// the same for an fp64 plane (halo above the slab of a vector that has its slack planes in place)
int tfb_halo_up_f64(tfb_ctx* c, const double* first_plane, double* ghost_above, size_t count) {
    if (c->nranks <= 1) return 0;
    TFB_CHECK(c->nccl_comm, "tfb_comm_init has not been called");
    ncclComm_t_ comm = (ncclComm_t_)c->nccl_comm;
    TFB_TRACE_OP(c, "halo up f64", count, c->stream);
    TFB_NCCL(nccl.GroupStart());
    if (c->rank > 0) TFB_NCCL(nccl.Send(first_plane, count, NCCL_FLOAT64, c->rank - 1, comm, c->stream));
    if (c->rank < c->nranks - 1) TFB_NCCL(nccl.Recv(ghost_above, count, NCCL_FLOAT64, c->rank + 1, comm, c->stream));
    TFB_NCCL(nccl.GroupEnd());
    TFB_LAUNCHED();
    return 0;
}
