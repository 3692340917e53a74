// runtime.cu — device/stream plumbing exported through include/ppo_b200.h ("runtime plumbing").
#include "common.cuh"

namespace b200 {

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
const char* ppo_b200_version(void) { return "ppo.c_b200 0.1 (sm_100a)"; }
void openblas_set_num_threads(int n) { (void)n; }

}  // extern "C"
