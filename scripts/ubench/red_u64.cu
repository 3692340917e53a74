// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/red_u64 scripts/ubench/red_u64.cu && /tmp/red_u64
// How fast can 148 CTAs add their 4483-element gradient contributions into ONE shared accumulator vector with 64-bit integer
// reductions (deterministic fixed-point sums), compared with writing private slabs?  Per "step": every CTA issues n elements.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(unsigned long long* acc, int n, int steps, int per_cta_copies) {
    for (int s = 0; s < steps; s++) {
        unsigned long long* a = acc + (size_t)(s % 3) * n;
        for (int c = 0; c < per_cta_copies; c++)
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const unsigned long long v = (unsigned long long)(e + s + blockIdx.x + c);
                asm volatile("red.global.add.u64 [%0], %1;" :: "l"(a + e), "l"(v) : "memory");
            }
        __syncthreads();
    }
}
__global__ void k_slab(float* slabs, int n, int steps, int per_cta_copies) {
    for (int s = 0; s < steps; s++) {
        for (int c = 0; c < per_cta_copies; c++) {
            float* a = slabs + ((size_t)blockIdx.x * per_cta_copies + c) * n;
            for (int e = threadIdx.x; e < n; e += blockDim.x) a[e] = (float)(e + s);
        }
        __syncthreads();
    }
}
int main() {
    const int n = 4483, steps = 2000;
    unsigned long long* acc; float* slabs;
    cudaMalloc(&acc, 3 * n * 8); cudaMemset(acc, 0, 3 * n * 8);
    cudaMalloc(&slabs, (size_t)148 * 2 * n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int copies = 1; copies <= 2; copies++) {
        for (int which = 0; which < 2; which++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) k_red<<<148, 512>>>(acc, n, steps, copies); else k_slab<<<148, 512>>>(slabs, n, steps, copies);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s copies/CTA %d: %.3f us per step (148 CTAs x %d elements)\n", which == 0 ? "red.add.u64 shared accumulators" : "private slab stores", copies, 1e3 * ms / steps, n * copies);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
