// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/lds_patterns scripts/ubench/lds_patterns.cu
// Shared-memory wavefront cost of the load patterns an fp32 register-tile GEMM can use (B200):
// cycles per warp-level load instruction, all 16 warps of a 512-thread CTA loading back to back (LSU-bound loop).
#include <cstdio>
#include <cuda_runtime.h>

template <int W>   // W = words per load (1, 2, 4)
__device__ __forceinline__ void lds(float (&v)[4], unsigned addr) {
    if (W == 4) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
    else if (W == 2) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr)); v[2] = v[3] = 0.f; }
    else { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr)); v[1] = v[2] = v[3] = 0.f; }
}

// pattern: 0 = every lane distinct, 1 = 8 distinct chunks (lane & 7), 2 = 4 distinct chunks (lane >> 3), 3 = all lanes one address,
//          4 = 2 distinct (lane >> 4), 5 = 16 distinct (lane & 15)
template <int W>
__global__ void __launch_bounds__(512, 1) k_lds(float* out, int iters, int pattern, long long* cycles) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int slot = lane;
    if (pattern == 1) slot = lane & 7;
    if (pattern == 2) slot = lane >> 3;
    if (pattern == 3) slot = 0;
    if (pattern == 4) slot = lane >> 4;
    if (pattern == 5) slot = lane & 15;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(slot * W * 4) + (unsigned)(warp * 512);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            float v[4];
            lds<W>(v, base + ((u * 2048 + it * 16) & 16383));
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const char* names[] = {"32 distinct", "8 distinct (lane&7)", "4 distinct (lane>>3)", "uniform", "2 distinct (lane>>4)", "16 distinct (lane&15)"};
    for (int W : {4, 2, 1})
        for (int p = 0; p < 6; p++) {
            for (int rep = 0; rep < 2; rep++) {
                if (W == 4) k_lds<4><<<148, 512, 32768>>>(out, iters, p, cyc);
                if (W == 2) k_lds<2><<<148, 512, 32768>>>(out, iters, p, cyc);
                if (W == 1) k_lds<1><<<148, 512, 32768>>>(out, iters, p, cyc);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            // 16 warps x iters x 16 loads per CTA in h[0] cycles -> cycles of the SM's LSU per warp-level load
            printf("LDS.%d %-22s %.2f cycles per warp load\n", W * 32, names[p], (double)h[0] / ((double)iters * 16 * 16));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
