// dist.cu — data parallelism over the GPUs of one box (nothing like it in the reference).
//
// One process per GPU.  The caller (bench.py / tests, via torch.distributed or any other out-of-band
// channel) broadcasts rank 0's NCCL unique id and calls ppo_b200_dist_init on every rank.  The path
// has exactly two exchange steps:
//   * sum all-reduce of the flat fp32 gradient of one net per minibatch (one contiguous arena, in
//     place, on the library's stream so it is ordered between backward and Adam), and
//   * all-gather of one (mean, M2, n) float64 triple per iteration for the advantage statistics.
// NCCL is bound with dlopen so single-GPU users never need it at load time; when torch is already in
// the process its bundled libnccl.so.2 is the one that gets picked up.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "internal.h"

namespace b200 {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1, g_shard_mode = 0;

#define NCCL_CHECK(expr)                                                                   \
    do {                                                                                   \
        ncclResult_t r__ = (expr);                                                         \
        if (r__ != ncclSuccess) B200_FATAL("%s -> %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
    } while (0)

static void load_nccl() {
    if (g_nccl.handle) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) B200_FATAL("cannot load libnccl.so.2: %s", dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(g_nccl.handle, s);
        if (!p) B200_FATAL("libnccl lacks %s", s);
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
}

bool dist_active() { return g_comm != nullptr && g_world > 1; }
int dist_rank() { return g_rank; }
int dist_world() { return g_world; }
int dist_shard_mode() { return g_shard_mode; }

void dist_allreduce_sum(float* buf, size_t count) {
    if (!dist_active() || count == 0) return;
    NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, ncclFloat, ncclSum, g_comm, stream()));
    ++g_launches;
}

void dist_allgather_doubles(const double* send, double* recv, int count_per_rank) {
    if (!dist_active()) {
        CUDA_CHECK(cudaMemcpyAsync(recv, send, count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
        return;
    }
    NCCL_CHECK(g_nccl.AllGather(send, recv, count_per_rank, ncclDouble, g_comm, stream()));
    ++g_launches;
}

}  // namespace b200

using namespace b200;

extern "C" {

void ppo_b200_dist_unique_id(char id[PPO_B200_NCCL_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= PPO_B200_NCCL_ID_BYTES, "id buffer too small");
    load_nccl();
    ncclUniqueId uid;
    NCCL_CHECK(g_nccl.GetUniqueId(&uid));
    memset(id, 0, PPO_B200_NCCL_ID_BYTES);
    memcpy(id, &uid, sizeof(uid));
}

void ppo_b200_dist_init(const char id[PPO_B200_NCCL_ID_BYTES], int rank, int world_size) {
    ensure_device();
    if (g_comm) B200_FATAL("ppo_b200_dist_init called twice");
    g_rank = rank;
    g_world = world_size;
    if (world_size <= 1) return;
    load_nccl();
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_CHECK(g_nccl.CommInitRank(&g_comm, world_size, uid, rank));
}

void ppo_b200_dist_finalize(void) {
    if (g_comm) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        NCCL_CHECK(g_nccl.CommDestroy(g_comm));
        g_comm = nullptr;
    }
    g_rank = 0;
    g_world = 1;
}

int ppo_b200_dist_rank(void) { return g_rank; }
int ppo_b200_dist_world(void) { return g_world; }
void ppo_b200_dist_set_shard_mode(int mode) { g_shard_mode = mode; }

}  // extern "C"
