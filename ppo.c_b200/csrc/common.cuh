// common.cuh — shared plumbing for the sm_100a kernels behind include/ppo_b200.h.
//
// Error convention: the reference checks nothing in release and sync+exit(1)s in
// debug (include/cuda_helper.h:4-19).  Here every CUDA/NCCL return is checked and a failure prints
// file:line and aborts.  There is deliberately NO CPU fallback anywhere in this directory.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ppo_b200.h"

#define B200_FATAL(...)                                                         \
    do {                                                                        \
        fprintf(stderr, "ppo_b200 fatal (%s:%d): ", __FILE__, __LINE__);         \
        fprintf(stderr, __VA_ARGS__);                                           \
        fprintf(stderr, "\n*** FAILED - ABORTING\n");                           \
        abort();                                                                \
    } while (0)

#define CUDA_CHECK(expr)                                                        \
    do {                                                                        \
        cudaError_t err__ = (expr);                                             \
        if (err__ != cudaSuccess) B200_FATAL("%s -> %s", #expr, cudaGetErrorString(err__)); \
    } while (0)

namespace b200 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- runtime state (runtime.cu) ---------------------------------------------------------------
cudaStream_t stream();            // the stream every kernel of this library is launched on
void ensure_device();             // aborts when no CUDA device is usable
extern unsigned long long g_launches;
int num_sms();

// Grow-only device scratch, keyed by slot, so hot paths never cudaMalloc (the reference mallocs per
// minibatch, src/loss.cu:51, src/ppo.cu:150).
enum ScratchSlot { kScratchGae = 0, kScratchGaeV, kScratchGaeVNext, kScratchLoss, kScratchStage,
                   kScratchStage2, kScratchStage3, kScratchPartials, kScratchMisc, kScratchSkinny, kScratchColsum, kScratchSlots };
void* scratch(ScratchSlot slot, size_t bytes);
void* scratch_zeroed_once(ScratchSlot slot, size_t bytes);  // zero-filled when (re)allocated only

template <typename T>
T* dmalloc(size_t count) {
    ensure_device();
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 4));
    return static_cast<T*>(p);
}
template <typename T>
T* hmalloc_pinned(size_t count) {
    ensure_device();
    void* p = nullptr;
    CUDA_CHECK(cudaHostAlloc(&p, count * sizeof(T) > 0 ? count * sizeof(T) : 4, cudaHostAllocPortable | cudaHostAllocMapped));
    return static_cast<T*>(p);
}

// Every kernel launch of the library goes through this macro: launch counter (bench.py's
// gpu_launches) and, when ppo_b200_profile_begin() is active, a CUDA-event pair per launch on the
// launching stream so per-kernel durations can be reported without a profiler attached.
void profile_mark(const char* name, bool begin);
extern bool g_profiling;
#define B200_LAUNCH(kernel, grid, block, smem, ...)                              \
    do {                                                                         \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, true);            \
        kernel<<<(grid), (block), (smem), ::b200::stream()>>>(__VA_ARGS__);      \
        ++::b200::g_launches;                                                    \
        CUDA_CHECK(cudaGetLastError());                                          \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, false);           \
    } while (0)

// Same, through cudaLaunchKernelEx with programmatic dependent launch: when `pdl` is true the kernel may
// start while its predecessor in the stream is still running; it must execute griddepcontrol.wait before
// touching anything the predecessor writes.
#define B200_LAUNCH_PDL(kernel, grid, block, smem, pdl, ...)                                    \
    do {                                                                                        \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, true);                           \
        cudaLaunchConfig_t cfg__ = {};                                                          \
        cfg__.gridDim = dim3(grid); cfg__.blockDim = dim3(block);                               \
        cfg__.dynamicSmemBytes = (smem); cfg__.stream = ::b200::stream();                       \
        cudaLaunchAttribute attr__[1];                                                          \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
        attr__[0].val.programmaticStreamSerializationAllowed = (pdl) ? 1 : 0;                   \
        cfg__.attrs = attr__; cfg__.numAttrs = 1;                                               \
        CUDA_CHECK(cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__));                            \
        ++::b200::g_launches;                                                                   \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, false);                          \
    } while (0)

// Cooperative launch (every CTA co-resident: the kernel may use grid-wide barriers).  `args_ptr` = address of the one
// by-value argument struct.
#define B200_LAUNCH_COOP(kernel, grid, block, smem, args_ptr)                                    \
    do {                                                                                        \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, true);                           \
        void* kargs__[1] = {(void*)(args_ptr)};                                                 \
        CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(block), kargs__, (smem), ::b200::stream())); \
        ++::b200::g_launches;                                                                   \
        if (::b200::g_profiling) ::b200::profile_mark(#kernel, false);                          \
    } while (0)

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// streaming 128-bit loads/stores that do not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

enum Act { kActNone = 0, kActRelu = 1, kActTanh = 2 };
inline int act_code(const char* name) {
    if (name && strcmp(name, "relu") == 0) return kActRelu;
    if (name && strcmp(name, "tanh") == 0) return kActTanh;
    return kActNone;  // src/activation_function.cu:49-55: anything else is the identity
}
// tanh in ~14 branch-free instructions, two of them MUFU (libdevice tanhf is ~50 with a divergent branch; at 2x64 the 64x64
// tanh epilogues were costing more issue slots than the K=3 layer they follow):
//   |x| >= 1/16: 1 - 2 / (e^{2|x|} + 1)  with ex2.approx / rcp.approx: absolute error <= ~1.5e-7, i.e. <= 2.4e-6 relative
//   |x| <  1/16: x (1 - x^2/3 + 2 x^4/15)                              : relative error < 4e-9 (next term 17 x^6 / 315)
// [EXT] tanh has no counterpart in the reference (src/activation_function.cu:46-58); the oracle's is glibc tanhf.
__device__ __forceinline__ float tanh_fast(float x) {
    const float ax = fabsf(x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.88539008177792681f));      // e^{2|x|}
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
    const float x2 = x * x;
    const float small = x * fmaf(x2, fmaf(x2, 0.13333333333f, -0.33333333333f), 1.0f);
    return ax < 0.0625f ? small : big;
}
__device__ __forceinline__ float act_apply(float x, int act) {
    if (act == kActRelu) return x > 0.f ? x : 0.f;
    if (act == kActTanh) return tanh_fast(x);
    return x;
}
// derivative expressed through the POST-activation value y (src/activation_function.cu:11-15)
__device__ __forceinline__ float act_grad(float y, float g, int act) {
    if (act == kActRelu) return y > 0.f ? g : 0.f;
    if (act == kActTanh) return g * (1.f - y * y);
    return g;
}

}  // namespace b200
