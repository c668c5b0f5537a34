// NCCL plumbing for the row-sharded path.  NCCL is dlopen'ed on first
// use so that the library loads (and its symbols can be checked) on a machine without
// NCCL/GPUs, and so that inside a PyTorch process the already-loaded libnccl.so.2 is
// shared instead of a second copy.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace gsi {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool loaded = false;
};

static NcclApi& nccl() {
    static NcclApi api;
    if (!api.loaded) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        GSI_REQUIRE(h != nullptr, GSI_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define GSI_SYM(field, name)                                                                       \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));                         \
        GSI_REQUIRE(api.field != nullptr, GSI_ERR_NCCL, std::string("NCCL symbol missing: ") + name);
        GSI_SYM(GetUniqueId, "ncclGetUniqueId")
        GSI_SYM(CommInitRank, "ncclCommInitRank")
        GSI_SYM(CommDestroy, "ncclCommDestroy")
        GSI_SYM(AllGather, "ncclAllGather")
        GSI_SYM(AllReduce, "ncclAllReduce")
        GSI_SYM(Broadcast, "ncclBroadcast")
        GSI_SYM(GroupStart, "ncclGroupStart")
        GSI_SYM(GroupEnd, "ncclGroupEnd")
        GSI_SYM(GetErrorString, "ncclGetErrorString")
#undef GSI_SYM
        api.loaded = true;
    }
    return api;
}

#define GSI_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess)                                                                      \
            throw gsi::Error(GSI_ERR_NCCL, std::string("NCCL error: ") + nccl().GetErrorString(_r) + \
                                               " (" #expr ")");                                     \
    } while (0)

void comm_unique_id(void* out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    GSI_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, 128);
}

void comm_init(gsi_ctx* ctx, const void* unique_id) {
    GSI_REQUIRE(unique_id != nullptr, GSI_ERR_INVALID_ARGUMENT, "world > 1 needs a NCCL unique id");
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclComm_t comm;
    GSI_NCCL(nccl().CommInitRank(&comm, ctx->world, id, ctx->rank));
    ctx->nccl_comm = comm;
}

void comm_destroy(gsi_ctx* ctx) {
    if (ctx->nccl_comm) {
        nccl().CommDestroy(reinterpret_cast<ncclComm_t>(ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
}

void comm_allgather(gsi_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    if (ctx->world == 1) {
        if (send != recv) GSI_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return;
    }
    GSI_NCCL(nccl().AllGather(send, recv, bytes_per_rank, ncclChar, reinterpret_cast<ncclComm_t>(ctx->nccl_comm),
                              ctx->stream));
}

void comm_allreduce_sum(gsi_ctx* ctx, double* buf, size_t count) {
    if (ctx->world == 1) return;
    GSI_NCCL(nccl().AllReduce(buf, buf, count, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(ctx->nccl_comm),
                              ctx->stream));
}

void comm_broadcast(gsi_ctx* ctx, double* buf, size_t count, int root) {
    if (ctx->world == 1) return;
    GSI_NCCL(nccl().Broadcast(buf, buf, count, ncclDouble, root, reinterpret_cast<ncclComm_t>(ctx->nccl_comm),
                              ctx->stream));
}

void comm_allgatherv(gsi_ctx* ctx, double* full, const int64_t* offsets, const int64_t* counts) {
    if (ctx->world == 1) return;
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->nccl_comm);
    GSI_NCCL(nccl().GroupStart());
    for (int r = 0; r < ctx->world; ++r) {
        if (counts[r] == 0) continue;
        GSI_NCCL(nccl().Broadcast(full + offsets[r], full + offsets[r], (size_t)counts[r], ncclDouble, r, comm,
                                  ctx->stream));
    }
    GSI_NCCL(nccl().GroupEnd());
}

}  // namespace gsi
