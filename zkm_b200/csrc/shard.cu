// In-segment sharding runtime: the NCCL communicator of the ranks that prove one segment together (shard.cuh).
#include "shard.cuh"
#include "batch.cuh"
#include <dlfcn.h>
#include <cstring>

namespace zkm {

namespace {
// the handful of NCCL entry points used, resolved with dlsym (types follow nccl.h 2.x; the ABI of these calls is stable)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint64 = 5 };
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
Shard g_shard;

template <class F>
void bind(F& f, const char* name) {
    f = (F)dlsym(g_nccl.handle, name);
    if (!f) throw std::runtime_error(std::string("zkm_b200: NCCL symbol not found: ") + name);
}
void load_nccl() {
    if (g_nccl.handle) return;
    // RTLD_NOLOAD first: reuse the copy torch.distributed (or the host application) already mapped
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error(std::string("zkm_b200: in-segment sharding needs NCCL (libnccl.so.2): ") + dlerror());
    g_nccl.handle = h;
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
    bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommDestroy, "ncclCommDestroy");
    bind(g_nccl.AllGather, "ncclAllGather");
    bind(g_nccl.Broadcast, "ncclBroadcast");
    bind(g_nccl.GetErrorString, "ncclGetErrorString");
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != 0) throw std::runtime_error(std::string("zkm_b200: ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
}  // namespace

const Shard& shard() { return g_shard; }

void shard_unique_id(unsigned char out[128]) {
    load_nccl();
    ncclUniqueId id;
    nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}

void shard_init(int rank, int world, const unsigned char id[128]) {
    ZKM_CHECK(world == 1 || world == 2 || world == 4 || world == 8, "in-segment sharding supports 1, 2, 4 or 8 ranks (the 4 cosets of the rate-4 LDE, halved once)");
    ZKM_CHECK(rank >= 0 && rank < world, "shard rank out of range");
    shard_shutdown();
    if (world == 1) return;
    ZKM_CHECK(id, "null NCCL unique id");
    Ctx& c = ctx();
    ZKM_CUDA(cudaSetDevice(c.device));
    load_nccl();
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    nccl_check(g_nccl.CommInitRank(&g_comm, world, uid, rank), "ncclCommInitRank");
    g_shard.rank = rank;
    g_shard.world = world;
}

void shard_shutdown() {
    if (g_comm) {
        if (ctx_ready()) cudaStreamSynchronize(ctx().stream);
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_shard = Shard();
}

void shard_all_gather(const u64* d_send, u64* d_recv, size_t count_per_rank, cudaStream_t s) {
    ZKM_CHECK(g_comm, "shard communicator not initialised");
    nccl_check(g_nccl.AllGather(d_send, d_recv, count_per_rank, ncclUint64, g_comm, s), "ncclAllGather");
}
void shard_broadcast(u64* d_buf, size_t count, int root, cudaStream_t s) {
    ZKM_CHECK(g_comm, "shard communicator not initialised");
    nccl_check(g_nccl.Broadcast(d_buf, d_buf, count, ncclUint64, root, g_comm, s), "ncclBroadcast");
}

}  // namespace zkm
