// ubench.cu — measured fp32 FMA ceilings of this GPU, used by bench.py as the roofline denominator of the fp32 kernels
// (MEASURED_PEAKS.json holds only HBM and bf16-tensor figures).  Three pure-FFMA streams (checked in SASS):
//   form 0  FFMA  R, R, c[..], c[..]   two constant-bank operands          -> the fp32 pipe's real peak
//   form 1  FFMA  R, R, R, R           8x4 register outer-product tile     -> scalar register-operand ceiling
//   form 2  FFMA2 R, R, R, R           same tile, rows packed in pairs     -> the form every MLP inner loop of this library uses
#include "common.cuh"

namespace b200 {

__global__ void ubench_ffma_const(float* out, float a, float b, int iters) {
    float acc[16];
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ubench_ffma_reg(float* out, const float* in, int iters) {
    float acc[8][4], a[8], w[4];
    for (int r = 0; r < 8; r++) { a[r] = in[threadIdx.x + r * 32]; for (int c = 0; c < 4; c++) acc[r][c] = 0.f; }
    for (int c = 0; c < 4; c++) w[c] = in[threadIdx.x + 256 + c * 32];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], w[c], acc[r][c]);
    }
    float s = 0;
    for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ubench_ffma2_reg(float* out, const float* in, int iters) {
    float2 acc[4][4], a[4], w[4];
    for (int r = 0; r < 4; r++) {
        a[r] = make_float2(in[threadIdx.x + r * 64], in[threadIdx.x + r * 64 + 32]);
        for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f);
    }
    for (int c = 0; c < 4; c++) { const float x = in[threadIdx.x + 256 + c * 32]; w[c] = make_float2(x, x); }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = __ffma2_rn(a[r], w[c], acc[r][c]);
    }
    float s = 0;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) s += acc[r][c].x + acc[r][c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace b200

using namespace b200;

// TFLOP/s of the chosen FFMA form on the current device: 256-thread CTAs, 1024 threads per SM, CUDA events on the library's
// stream, best of 3 timed launches after a warm-up.
extern "C" double ppo_b200_measure_fp32_peak(int form) {
    const int threads = 256, blocks = num_sms() * 4, iters = 4000;
    float* out = dmalloc<float>((size_t)blocks * threads);
    float* in = dmalloc<float>(4096);
    CUDA_CHECK(cudaMemsetAsync(in, 0, 4096 * sizeof(float), stream()));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0;
    const double flops = form == 0 ? 2.0 * 16 * (iters * 8.0) * blocks * threads : 2.0 * 128 * iters * (double)blocks * threads;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_CHECK(cudaEventRecord(e0, stream()));
        if (form == 0) ubench_ffma_const<<<blocks, threads, 0, stream()>>>(out, 1.0001f, 0.5f, iters * 8);
        else if (form == 1) ubench_ffma_reg<<<blocks, threads, 0, stream()>>>(out, in, iters);
        else ubench_ffma2_reg<<<blocks, threads, 0, stream()>>>(out, in, iters);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaEventRecord(e1, stream()));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    CUDA_CHECK(cudaEventDestroy(e0));
    CUDA_CHECK(cudaEventDestroy(e1));
    CUDA_CHECK(cudaFree(out));
    CUDA_CHECK(cudaFree(in));
    return best;
}
