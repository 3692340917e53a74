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
#include <sched.h>

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

// ---- peer arena: NVLink peer memory for the fused small-net gradient exchange -----------------------------
// For the reference-width nets the gradient is 18 KB: an NCCL all-reduce is pure launch + protocol latency
// (three extra launches per minibatch).  Instead every rank owns one cudaMalloc'd RECEIVE buffer
//     u64 recv[2 parities][source rank][element]         (element = {epoch tag : 32 | fp32 bits : 32})
// exported with cudaIpcGetMemHandle; the handles travel once through the NCCL communicator and every rank maps
// every other buffer (cudaIpcOpenMemHandle = peer stores over NVLink).  The slab-reduce + Adam kernel
// (fused_mlp.cu: fused_reduce_adam_kernel with a PeerView) reduces its own slabs, PUSHES each element together
// with the epoch tag into lane [my rank] of every peer's buffer with one 64-bit store (value and flag are one
// naturally atomic word, so no fence and no separate flag round trip: one NVLink hop of latency), then polls its
// own buffer until every source lane carries the tag and adds the G values in rank order before Adam.  Every rank
// adds the same numbers in the same order, so the replicas stay bit-identical.  The two parities alternate per
// exchange: a lane is rewritten two exchanges later, which its owner can only reach after the reader has
// finished the kernel that consumed it (stream order), see DESIGN.md §7.
static PeerArena g_peer;

static void peer_setup() {
    if (g_peer.tried) return;
    g_peer.tried = true;
    const char* off = getenv("PPO_B200_PEER");
    if (off && off[0] == '0') return;
    if (g_world > kPeerMaxRanks) return;
    const size_t bytes = 2 * kPeerMaxRanks * kPeerCap * sizeof(unsigned long long);
    char* local = nullptr;
    CUDA_CHECK(cudaMalloc(&local, bytes));
    CUDA_CHECK(cudaMemset(local, 0, bytes));
    cudaIpcMemHandle_t mine;
    CUDA_CHECK(cudaIpcGetMemHandle(&mine, local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    char *d_send = nullptr, *d_recv = nullptr;
    CUDA_CHECK(cudaMalloc(&d_send, 64));
    CUDA_CHECK(cudaMalloc(&d_recv, 64 * g_world));
    CUDA_CHECK(cudaMemcpy(d_send, &mine, 64, cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.AllGather(d_send, d_recv, 64, ncclChar, g_comm, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    std::vector<cudaIpcMemHandle_t> all(g_world);
    CUDA_CHECK(cudaMemcpy(all.data(), d_recv, 64 * g_world, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaFree(d_send));
    CUDA_CHECK(cudaFree(d_recv));
    bool ok = true;
    for (int r = 0; r < g_world; r++) {
        if (r == g_rank) { g_peer.base[r] = local; continue; }
        void* ptr = nullptr;
        const cudaError_t err = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) { (void)cudaGetLastError(); ok = false; break; }
        g_peer.base[r] = static_cast<char*>(ptr);
    }
    // everybody must agree (a rank that cannot map its peers would otherwise wait for flags nobody writes)
    int* d_ok = nullptr;
    CUDA_CHECK(cudaMalloc(&d_ok, sizeof(int)));
    const int mine_ok = ok ? 1 : 0;
    CUDA_CHECK(cudaMemcpy(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, g_comm, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    int all_ok = 0;
    CUDA_CHECK(cudaMemcpy(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaFree(d_ok));
    g_peer.local = local;
    g_peer.ready = all_ok == 1;
    g_peer.epoch = 0;
    if (!g_peer.ready && g_rank == 0) fprintf(stderr, "ppo_b200: peer memory not available, small-net gradients go through NCCL\n");
}

static void peer_teardown() {
    if (g_peer.local) {
        for (int r = 0; r < g_world; r++)
            if (r != g_rank && g_peer.base[r]) (void)cudaIpcCloseMemHandle(g_peer.base[r]);
        CUDA_CHECK(cudaFree(g_peer.local));
    }
    g_peer = PeerArena{};
}

bool dist_peer_ready() {
    if (!dist_active()) return false;
    peer_setup();
    return g_peer.ready;
}

// PeerView of the next exchange (advances the epoch), or ready == 0 when peer memory is not in use.
PeerView dist_peer_next(size_t vec_floats) {
    PeerView v{};
    if (!dist_active()) return v;
    peer_setup();
    if (!g_peer.ready || vec_floats > kPeerCap) return v;
    const unsigned long long epoch = ++g_peer.epoch;
    if ((epoch & 0xffffffffull) == 0) B200_FATAL("peer exchange epoch wrapped");
    const size_t parity_off = (epoch & 1) * kPeerMaxRanks * kPeerCap;
    v.ready = 1;
    v.world = g_world;
    v.rank = g_rank;
    v.epoch = (unsigned int)epoch;
    v.my_recv = reinterpret_cast<unsigned long long*>(g_peer.local) + parity_off;
    for (int r = 0; r < g_world; r++)
        v.peer_recv[r] = reinterpret_cast<unsigned long long*>(g_peer.base[r]) + parity_off + (size_t)g_rank * kPeerCap;
    return v;
}

PeerView dist_peer_reserve(size_t vec_floats, int count) {
    PeerView v{};
    if (!dist_active() || count <= 0) return v;
    peer_setup();
    if (!g_peer.ready || vec_floats > kPeerCap) return v;
    const unsigned long long first = g_peer.epoch + 1;
    g_peer.epoch += (unsigned long long)count;
    if ((first >> 32) != (g_peer.epoch >> 32) || (first & 0xffffffffull) == 0) B200_FATAL("peer exchange epoch wrapped");
    v.ready = 1;
    v.world = g_world;
    v.rank = g_rank;
    v.epoch = (unsigned int)first;
    v.parity_stride = kPeerMaxRanks * kPeerCap;
    v.my_recv = reinterpret_cast<unsigned long long*>(g_peer.local);
    for (int r = 0; r < g_world; r++)
        v.peer_recv[r] = reinterpret_cast<unsigned long long*>(g_peer.base[r]) + (size_t)g_rank * kPeerCap;
    return v;
}

long long dist_spin_limit() {
    static long long cached = 0;
    if (!cached) {
        const char* e = getenv("PPO_B200_SPIN_TIMEOUT_S");
        const double sec = e ? atof(e) : 120.0;
        cached = (long long)(std::max(1.0, sec) * 2.0e9);
    }
    return cached;
}

// One process per GPU: keep the calling thread (and with it the first-touch placement of the pinned host mirrors it
// allocates next) on the CPUs of the NUMA node the GPU hangs off, so eight ranks' host<->device copies do not cross the
// socket interconnect.  Best effort: silently skipped when sysfs has no answer (containers, single-node hosts) or when
// PPO_B200_NUMA_BIND=0.
static void bind_to_gpu_numa_node() {
    const char* e = getenv("PPO_B200_NUMA_BIND");
    if (e && e[0] == '0') return;
    int dev = 0;
    char bus[64];
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return;
    char list[1024] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return;
    cpu_set_t cur, want;
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return;
    CPU_ZERO(&want);
    int n_set = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {      // "0-55,112-167"
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &cur)) { CPU_SET(c, &want); n_set++; }
    }
    if (n_set > 0) sched_setaffinity(0, sizeof(want), &want);
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
    bind_to_gpu_numa_node();
    load_nccl();
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_CHECK(g_nccl.CommInitRank(&g_comm, world_size, uid, rank));
}

void ppo_b200_dist_finalize(void) {
    if (g_comm) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        // every rank must be past its last peer read before anybody unmaps / frees an arena
        if (g_peer.ready) {
            int* d_tok = nullptr;
            CUDA_CHECK(cudaMalloc(&d_tok, sizeof(int)));
            CUDA_CHECK(cudaMemset(d_tok, 0, sizeof(int)));
            NCCL_CHECK(g_nccl.AllReduce(d_tok, d_tok, 1, ncclInt, ncclSum, g_comm, stream()));
            CUDA_CHECK(cudaStreamSynchronize(stream()));
            CUDA_CHECK(cudaFree(d_tok));
        }
        peer_teardown();
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
