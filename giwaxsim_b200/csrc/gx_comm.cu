// Multi-GPU exchange behind the C ABI: one NCCL communicator per process (one process per GPU), the
// two collectives of the path issued on the caller's stream, right behind the kernels that produce
// their inputs - no host synchronisation in between.
//
//   stage A   voxelgrids.py:502-503 (`voxel_grid[...] += ...` by every worker): partial sums of the
//             ranks -> gx_comm_reduce_scatter_f32 (each rank ends with the total of ITS slab of
//             columns), counts -> gx_comm_all_reduce (u32, 0.65 MB), then every rank finalises only
//             its slab (gx_voxel_finalize with a column range) and gx_comm_all_gather completes iq.
//   stage B   detector.py:298 (`det_ints += ...`): gx_comm_all_reduce of the fp64 partial images.
//
// NCCL is resolved at run time from the libnccl.so.2 the process has already loaded (PyTorch bundles
// it), so the library neither links against NCCL nor needs its headers; the handful of prototypes used
// are restated below (nccl.h 2.27: stable since 2.0).
#include <dlfcn.h>
#include <string.h>
#include "gx_common.cuh"

typedef struct { char internal[128]; } gx_nccl_id;         // ncclUniqueId
typedef void *gx_nccl_comm;                                // ncclComm_t
enum { GX_NCCL_SUM = 0, GX_NCCL_UINT32 = 3, GX_NCCL_FLOAT32 = 7, GX_NCCL_FLOAT64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(gx_nccl_id *);
    int (*CommInitRank)(gx_nccl_comm *, int, gx_nccl_id, int);
    int (*CommDestroy)(gx_nccl_comm);
    const char *(*GetErrorString)(int);
    int (*AllReduce)(const void *, void *, size_t, int, int, gx_nccl_comm, cudaStream_t);
    int (*ReduceScatter)(const void *, void *, size_t, int, int, gx_nccl_comm, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, gx_nccl_comm, cudaStream_t);
    bool ok;
};
static NcclApi g_nccl;

static int nccl_api()
{
    if (g_nccl.ok) return GX_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        gx_set_error("gx_comm: libnccl.so.2 cannot be loaded (%s)", dlerror());
        return GX_ERR_UNSUPPORTED;
    }
#define GX_SYM(field, name)                                                          \
    do {                                                                             \
        *reinterpret_cast<void **>(&g_nccl.field) = dlsym(h, name);                  \
        if (!g_nccl.field) {                                                         \
            gx_set_error("gx_comm: %s missing from libnccl", name);                  \
            return GX_ERR_UNSUPPORTED;                                               \
        }                                                                            \
    } while (0)
    GX_SYM(GetUniqueId, "ncclGetUniqueId");
    GX_SYM(CommInitRank, "ncclCommInitRank");
    GX_SYM(CommDestroy, "ncclCommDestroy");
    GX_SYM(GetErrorString, "ncclGetErrorString");
    GX_SYM(AllReduce, "ncclAllReduce");
    GX_SYM(ReduceScatter, "ncclReduceScatter");
    GX_SYM(AllGather, "ncclAllGather");
#undef GX_SYM
    g_nccl.ok = true;
    return GX_OK;
}

#define GX_NCCL(call)                                                                         \
    do {                                                                                      \
        int r_ = (call);                                                                      \
        if (r_ != 0) {                                                                        \
            gx_set_error("%s: %s -> %s", __func__, #call, g_nccl.GetErrorString(r_));         \
            return GX_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

struct GxComm {
    gx_nccl_comm comm;
    int rank, world;
};

extern "C" int gx_comm_unique_id(void *h_id128)
{
    GX_REQUIRE(h_id128 != NULL, "NULL id buffer");
    if (int e = nccl_api()) return e;
    gx_nccl_id id;
    GX_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(h_id128, &id, sizeof(id));
    return GX_OK;
}

extern "C" int gx_comm_init(const void *h_id128, int rank, int world, void **out_comm)
{
    GX_REQUIRE(h_id128 && out_comm && world >= 1 && rank >= 0 && rank < world, "bad arguments");
    if (int e = nccl_api()) return e;
    gx_nccl_id id;
    memcpy(&id, h_id128, sizeof(id));
    GxComm *c = new GxComm;
    c->rank = rank; c->world = world; c->comm = NULL;
    int r = g_nccl.CommInitRank(&c->comm, world, id, rank);      // uses the calling thread's current device
    if (r != 0) {
        gx_set_error("gx_comm_init: ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
        delete c;
        return GX_ERR_CUDA;
    }
    *out_comm = c;
    return GX_OK;
}

extern "C" int gx_comm_destroy(void *comm)
{
    if (!comm) return GX_OK;
    GxComm *c = static_cast<GxComm *>(comm);
    if (g_nccl.ok && c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return GX_OK;
}

static int nccl_type(int dtype, int *out)
{
    switch (dtype) {
    case GX_DTYPE_F32: *out = GX_NCCL_FLOAT32; return GX_OK;
    case GX_DTYPE_U32: *out = GX_NCCL_UINT32; return GX_OK;
    case GX_DTYPE_F64: *out = GX_NCCL_FLOAT64; return GX_OK;
    }
    gx_set_error("gx_comm: unknown dtype %d", dtype);
    return GX_ERR_INVALID;
}

extern "C" int gx_comm_all_reduce(void *comm, void *d_buf, int64_t count, int dtype, void *stream)
{
    GX_REQUIRE(comm && d_buf && count >= 0, "bad arguments");
    GxComm *c = static_cast<GxComm *>(comm);
    int t;
    if (int e = nccl_type(dtype, &t)) return e;
    if (count == 0 || c->world == 1) return GX_OK;
    GX_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, t, GX_NCCL_SUM, c->comm, gx_stream(stream)));
    return GX_OK;
}

// d_buf holds world * count_per_rank fp32 partial sums; on return the slab
// [rank * count_per_rank, (rank + 1) * count_per_rank) of THIS rank's d_buf holds the sum over ranks
// (in place: the other slabs keep their partial values).
extern "C" int gx_comm_reduce_scatter_f32(void *comm, float *d_buf, int64_t count_per_rank, void *stream)
{
    GX_REQUIRE(comm && d_buf && count_per_rank >= 0, "bad arguments");
    GxComm *c = static_cast<GxComm *>(comm);
    if (count_per_rank == 0 || c->world == 1) return GX_OK;
    GX_NCCL(g_nccl.ReduceScatter(d_buf, d_buf + (size_t)c->rank * count_per_rank, (size_t)count_per_rank,
                                 GX_NCCL_FLOAT32, GX_NCCL_SUM, c->comm, gx_stream(stream)));
    return GX_OK;
}

// every rank contributes its slab of d_buf (world * count_per_rank fp32), in place
extern "C" int gx_comm_all_gather_f32(void *comm, float *d_buf, int64_t count_per_rank, void *stream)
{
    GX_REQUIRE(comm && d_buf && count_per_rank >= 0, "bad arguments");
    GxComm *c = static_cast<GxComm *>(comm);
    if (count_per_rank == 0 || c->world == 1) return GX_OK;
    GX_NCCL(g_nccl.AllGather(d_buf + (size_t)c->rank * count_per_rank, d_buf, (size_t)count_per_rank,
                             GX_NCCL_FLOAT32, c->comm, gx_stream(stream)));
    return GX_OK;
}
