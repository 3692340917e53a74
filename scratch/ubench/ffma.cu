// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scratch/ubench/ffma scratch/ubench/ffma.cu ; run: ./scratch/ubench/ffma
// FFMA throughput by operand form (sm_100a).  All loops are pure FFMA streams (checked in SASS).
#include <cstdio>
#include <cuda_runtime.h>
// (1) two constant-bank operands
__global__ void k_const(float* out, float a, float b, int iters) {
    float acc[16];
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (2) three register operands, 8x4 outer-product tile with loop-invariant a[], w[] held in registers
__global__ void k_reg(float* out, const float* in, int iters) {
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
    float s = 0; for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// (3) same tile with packed FFMA2 (rows packed in pairs, weight broadcast)
__global__ void k_reg2(float* out, const float* in, int iters) {
    float2 acc[4][4], a[4], w[4];
    for (int r = 0; r < 4; r++) { a[r] = make_float2(in[threadIdx.x + r * 64], in[threadIdx.x + r * 64 + 32]); for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f); }
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
    float s = 0; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) s += acc[r][c].x + acc[r][c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> static void run(const char* name, F launch, double flops) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %.3f ms  %.1f TFLOP/s\n", name, ms, flops / ms / 1e9);
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    const int iters = 5000;
    for (int threads = 128; threads <= 512; threads *= 2) {
        const int blocks = 148 * (1024 / threads);     // 1024 threads per SM (regs allow it)
        printf("-- %d threads/CTA, %d CTAs\n", threads, blocks);
        run("FFMA  R,R,c,c (constant operands)", [&] { k_const<<<blocks, threads>>>(out, 1.0001f, 0.5f, iters * 8); }, 2.0 * 16 * iters * 8 * (double)blocks * threads);
        run("FFMA  R,R,R,R (8x4 register tile)", [&] { k_reg<<<blocks, threads>>>(out, in, iters); }, 2.0 * 128 * iters * (double)blocks * threads);
        run("FFMA2 R,R,R,R (8x4 tile, rows packed)", [&] { k_reg2<<<blocks, threads>>>(out, in, iters); }, 2.0 * 128 * iters * (double)blocks * threads);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
