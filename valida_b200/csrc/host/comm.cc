// Multi-GPU plumbing: one process per GPU, an NCCL communicator per context, collectives enqueued on the
// context's stream (trace columns shard across the GPUs of one box for the LDE, rows for the
// Keccak leaves / tree layers, and the pieces meet again through NCCL all-gathers over NVLink).
// NCCL is resolved at run time (dlopen of libnccl.so.2 — the copy torch already loaded when the caller is a
// torchrun rank) so that a single-GPU user of the library needs no NCCL at all.
#include "../ctx.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstring>

namespace {

struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl n;
    if (n.lib || !n.why.empty()) return n;
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) { n.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return n; }
    auto sym = [&](const char* name) { void* p = dlsym(n.lib, name); if (!p && n.why.empty()) n.why = std::string("libnccl lacks ") + name; return p; };
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
    n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    if (!n.why.empty()) { n.lib = nullptr; }
    return n;
}

#define VG_NCCL(ctx, expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess) VG_FAIL(ctx, "%s failed: %s", #expr, nccl().GetErrorString(_r)); } while (0)

}  // namespace

void vg_shard_range(uint64_t total, int nranks, int rank, uint64_t* begin, uint64_t* end) {
    // the first (total % nranks) ranks take one extra unit
    const uint64_t q = total / (uint64_t)nranks, r = total % (uint64_t)nranks, k = (uint64_t)rank;
    *begin = k * q + (k < r ? k : r);
    *end = *begin + q + (k < r ? 1 : 0);
}

int32_t vg_comm_group_begin(vgpu_ctx* ctx) { VG_NCCL(ctx, nccl().GroupStart()); return 0; }
int32_t vg_comm_group_end(vgpu_ctx* ctx) { VG_NCCL(ctx, nccl().GroupEnd()); return 0; }
int32_t vg_comm_allgather_inplace(vgpu_ctx* ctx, uint32_t* buf, uint64_t words_per_rank) {
    VG_NCCL(ctx, nccl().AllGather(buf + (uint64_t)ctx->comm_rank * words_per_rank, buf, words_per_rank, ncclUint32, (ncclComm_t)ctx->nccl, ctx->stream));
    return 0;
}
int32_t vg_comm_bcast(vgpu_ctx* ctx, uint32_t* buf, uint64_t words, int root) {
    if (!words) return 0;
    VG_NCCL(ctx, nccl().Broadcast(buf, buf, words, ncclUint32, root, (ncclComm_t)ctx->nccl, ctx->stream));
    return 0;
}
void vg_comm_free(vgpu_ctx* ctx) {
    if (ctx->nccl) { nccl().CommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
    ctx->comm_size = 1; ctx->comm_rank = 0; ctx->sharding = false;
}

extern "C" {

int32_t vgpu_comm_unique_id(uint8_t out[VGPU_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == VGPU_COMM_ID_BYTES, "ncclUniqueId size");
    if (!out || !nccl().lib) return -1;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return -1;
    std::memcpy(out, &id, sizeof id);
    return 0;
}

int32_t vgpu_comm_init(vgpu_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t unique_id[VGPU_COMM_ID_BYTES]) {
    if (!ctx) return -1;
    if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) VG_FAIL(ctx, "comm_init: bad rank %d of %d", rank, nranks);
    if (nranks & (nranks - 1)) VG_FAIL(ctx, "comm_init: the number of ranks must be a power of two (tree layers are split evenly), got %d", nranks);
    if (!nccl().lib) VG_FAIL(ctx, "comm_init: %s", nccl().why.c_str());
    vg_comm_free(ctx);
    VG_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof id);
    ncclComm_t comm = nullptr;
    VG_NCCL(ctx, nccl().CommInitRank(&comm, nranks, id, rank));
    ctx->nccl = comm; ctx->comm_size = nranks; ctx->comm_rank = rank; ctx->sharding = nranks > 1;
    return 0;
}

int32_t vgpu_comm_set_sharding(vgpu_ctx* ctx, int32_t on) {
    if (!ctx) return -1;
    if (on && !ctx->nccl) VG_FAIL(ctx, "comm_set_sharding: vgpu_comm_init has not been called");
    ctx->sharding = on != 0;
    return 0;
}

// the share of a tree layer of `len` nodes that rank `rank` derives itself (merkle.cu share_of); *split = 0 when the
// layer is shorter than the communicator and every rank computes all of it
void vgpu_tree_share(uint64_t len, int32_t nranks, int32_t rank, uint64_t* begin, uint64_t* count, int32_t* split) {
    if (nranks > 1 && len >= (uint64_t)nranks) { *count = len / (uint64_t)nranks; *begin = *count * (uint64_t)rank; *split = 1; }
    else { *begin = 0; *count = len; *split = 0; }
}

void vgpu_shard_range(uint64_t total, int32_t nranks, int32_t rank, uint64_t* begin, uint64_t* end) { vg_shard_range(total, nranks, rank, begin, end); }

}  // extern "C"
