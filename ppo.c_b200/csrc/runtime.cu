// runtime.cu — device/stream plumbing exported through include/ppo_b200.h ("runtime plumbing").
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200 {

// ---- event-pair profiling of every launch (see B200_LAUNCH) ------------------------------------
bool g_profiling = false;
struct ProfRec { const char* name; cudaEvent_t beg, end; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static cudaEvent_t pool_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    return e;
}
void profile_mark(const char* name, bool begin) {
    if (begin) {
        ProfRec r{name, pool_event(), pool_event()};
        CUDA_CHECK(cudaEventRecord(r.beg, stream()));
        g_prof.push_back(r);
    } else {
        CUDA_CHECK(cudaEventRecord(g_prof.back().end, stream()));
    }
}

unsigned long long g_launches = 0;
static cudaStream_t g_stream = nullptr;
static bool g_stream_external = false;
static bool g_ready = false;
static int g_sms = kNumSMs;

struct Scratch { void* p = nullptr; size_t bytes = 0; };
static Scratch g_scratch[kScratchSlots];

void ensure_device() {
    if (g_ready) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        B200_FATAL("no usable CUDA device (%s): this library has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    g_ready = true;
}

cudaStream_t stream() {
    ensure_device();
    if (!g_stream) {
        // a *blocking* stream: legacy default-stream work of the caller (plain cudaMemcpy, as the
        // reference's callers use) orders against our kernels exactly as it did against the
        // reference's default-stream launches.
        CUDA_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamDefault));
    }
    return g_stream;
}

int num_sms() { ensure_device(); return g_sms; }

void* scratch(ScratchSlot slot, size_t bytes) {
    Scratch& s = g_scratch[slot];
    if (bytes > s.bytes) {
        if (s.p) { CUDA_CHECK(cudaStreamSynchronize(stream())); CUDA_CHECK(cudaFree(s.p)); }
        size_t cap = bytes + bytes / 4 + 256;
        CUDA_CHECK(cudaMalloc(&s.p, cap));
        CUDA_CHECK(cudaMemsetAsync(s.p, 0, cap, stream()));
        s.bytes = cap;
    }
    return s.p;
}
void* scratch_zeroed_once(ScratchSlot slot, size_t bytes) { return scratch(slot, bytes); }

}  // namespace b200

using namespace b200;

extern "C" {

int ppo_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void ppo_b200_set_device(int device) {
    CUDA_CHECK(cudaSetDevice(device));
}

void ppo_b200_set_stream(void* cuda_stream) {
    ensure_device();
    if (g_stream && !g_stream_external) { CUDA_CHECK(cudaStreamSynchronize(g_stream)); CUDA_CHECK(cudaStreamDestroy(g_stream)); }
    g_stream = static_cast<cudaStream_t>(cuda_stream);
    g_stream_external = cuda_stream != nullptr;
}

void* ppo_b200_malloc(size_t bytes) { return dmalloc<char>(bytes); }
void ppo_b200_free(void* dptr) { if (dptr) CUDA_CHECK(cudaFree(dptr)); }
void* ppo_b200_malloc_host(size_t bytes) { return hmalloc_pinned<char>(bytes); }
void ppo_b200_free_host(void* hptr) { if (hptr) CUDA_CHECK(cudaFreeHost(hptr)); }

void ppo_b200_h2d(void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void ppo_b200_d2h(void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void ppo_b200_memset(void* dst, int value, size_t bytes) {
    CUDA_CHECK(cudaMemsetAsync(dst, value, bytes, stream()));
}
void ppo_b200_sync(void) { CUDA_CHECK(cudaStreamSynchronize(stream())); }
unsigned long long ppo_b200_launch_count(void) { return g_launches; }
void ppo_b200_profile_begin(void) {
    for (auto& r : g_prof) { g_event_pool.push_back(r.beg); g_event_pool.push_back(r.end); }
    g_prof.clear();
    g_profiling = true;
}

// Stops profiling and writes one line per kernel: "name count total_ms\n".  Returns bytes written.
int ppo_b200_profile_end(char* out, int out_bytes) {
    g_profiling = false;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, r.beg, r.end));
        std::string name(r.name);
        const size_t lt = name.find('<');          // strip template arguments' namespace noise
        (void)lt;
        auto& a = agg[name];
        a.first += 1;
        a.second += ms;
        g_event_pool.push_back(r.beg);
        g_event_pool.push_back(r.end);
    }
    g_prof.clear();
    std::string s;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        s += line;
    }
    if (out && out_bytes > 0) {
        const int n = (int)std::min<size_t>(s.size(), (size_t)out_bytes - 1);
        memcpy(out, s.data(), n);
        out[n] = 0;
        return n;
    }
    return 0;
}

const char* ppo_b200_version(void) { return "ppo.c_b200 0.1 (sm_100a)"; }
void openblas_set_num_threads(int n) { (void)n; }

}  // extern "C"
