// 3-register FFMA vs FFMA2 throughput (sm_100a): outer-product tile 8x4 per thread like the MLP kernels.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, const float* in, int iters) {
    float acc[8][4], a[8], w[4];
    for (int r = 0; r < 8; r++) { a[r] = in[threadIdx.x + r * 32]; for (int c = 0; c < 4; c++) acc[r][c] = 0.f; }
    for (int c = 0; c < 4; c++) w[c] = in[threadIdx.x + 256 + c * 32];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], w[c], acc[r][c]);
        a[it & 7] += 1e-6f; w[it & 3] -= 1e-6f;     // keep operands live / varying
    }
    float s = 0; for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, const float* in, int iters) {
    float2 acc[4][4], a[4]; float w[4];        // rows packed in pairs
    for (int r = 0; r < 4; r++) { a[r] = make_float2(in[threadIdx.x + r * 64], in[threadIdx.x + r * 64 + 32]); for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f); }
    for (int c = 0; c < 4; c++) w[c] = in[threadIdx.x + 256 + c * 32];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = __ffma2_rn(a[r], make_float2(w[c], w[c]), acc[r][c]);
        a[it & 3].x += 1e-6f; w[it & 3] -= 1e-6f;
    }
    float s = 0; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) s += acc[r][c].x + acc[r][c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads = 256; threads <= 1024; threads *= 2) {
        const int blocks = 148 * (2048 / threads);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0); k_ffma<<<blocks, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 32 * iters * (double)blocks * threads;
            printf("threads %4d FFMA 3-reg: %.3f ms  %.1f TFLOP/s   ", threads, ms, fl / ms / 1e9);
            cudaEventRecord(e0); k_ffma2<<<blocks, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
