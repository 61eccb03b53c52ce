// Gradient all-reduce over NCCL behind the C ABI (cvae_comm_*): one communicator per process (one process per
// GPU), NVLink / NVSwitch inside one box.  libnccl.so.2 is bound at run time with dlopen -- the library PyTorch already
// loaded when there is one, the system's otherwise -- so libcvae.so has no link-time dependency on it and the
// single-GPU paths never touch it.  The rendezvous stays with the host: rank 0 calls cvae_comm_unique_id and ships the
// 128 bytes to the other ranks however it likes (the Python side broadcasts them over torch.distributed).
//
// Replaces nothing in the reference (it is single-GPU); it is the collective of the data-parallel training step
// (DESIGN.md section 4): sum the flat fp32 gradient buffer in place, Adam applies 1 / world.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace cvae {

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, const void*, int) = nullptr;   // ncclUniqueId is passed BY VALUE (128 bytes): see call site
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct UniqueId { char internal[128]; };                              // == ncclUniqueId
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);

static NcclApi g_nccl;
static void* g_comm = nullptr;
static int g_world = 0, g_rank = -1;
static std::mutex g_comm_mu;

static bool load_nccl() {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // already in the process (PyTorch's)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, const void*, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return false;
    g_nccl.handle = h;
    return true;
}
static const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"; }

}  // namespace cvae

using namespace cvae;

// 128 bytes identifying a new communicator; called by ONE rank, the bytes go to every rank's cvae_comm_init.
extern "C" int cvae_comm_unique_id(void* id128) {
    CVAE_REQUIRE(id128 != nullptr, CVAE_EINVAL, "comm_unique_id: null buffer");
    std::lock_guard<std::mutex> lock(g_comm_mu);
    CVAE_REQUIRE(load_nccl(), CVAE_ECUDA, "comm: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "no detail");
    const int rc = g_nccl.GetUniqueId(id128);
    CVAE_REQUIRE(rc == 0, CVAE_ECUDA, "ncclGetUniqueId: %s", nccl_err(rc));
    return CVAE_OK;
}

// Collective over all `world` ranks (blocks until every rank has called it).  The current CUDA device is the rank's GPU.
extern "C" int cvae_comm_init(int rank, int world, const void* id128) {
    CVAE_REQUIRE(id128 != nullptr && world >= 1 && rank >= 0 && rank < world, CVAE_EINVAL, "comm_init: rank %d of %d", rank, world);
    std::lock_guard<std::mutex> lock(g_comm_mu);
    CVAE_REQUIRE(g_comm == nullptr, CVAE_EINVAL, "comm_init: a communicator already exists (cvae_comm_destroy first)");
    CVAE_REQUIRE(load_nccl(), CVAE_ECUDA, "comm: libnccl.so.2 not found");
    UniqueId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = ((CommInitRankFn)g_nccl.CommInitRank)(&g_comm, world, id, rank);
    if (rc != 0) g_comm = nullptr;
    CVAE_REQUIRE(rc == 0, CVAE_ECUDA, "ncclCommInitRank: %s", nccl_err(rc));
    g_world = world;
    g_rank = rank;
    return CVAE_OK;
}

extern "C" int cvae_comm_world(void) { return g_comm ? g_world : 0; }

// In-place sum of `count` floats over the ranks, asynchronous on `stream` (every rank, same count, same order of calls).
extern "C" int cvae_comm_allreduce_sum(float* buf, int64_t count, void* stream) {
    CVAE_REQUIRE(g_comm != nullptr, CVAE_EINVAL, "comm_allreduce_sum: no communicator (cvae_comm_init)");
    CVAE_REQUIRE(buf != nullptr && count > 0, CVAE_EINVAL, "comm_allreduce_sum: empty buffer");
    const int rc = g_nccl.AllReduce(buf, buf, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, g_comm, (cudaStream_t)stream);
    CVAE_REQUIRE(rc == 0, CVAE_ECUDA, "ncclAllReduce: %s", nccl_err(rc));
    return CVAE_OK;
}

extern "C" int cvae_comm_destroy(void) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
        g_world = 0;
        g_rank = -1;
    }
    return CVAE_OK;
}
