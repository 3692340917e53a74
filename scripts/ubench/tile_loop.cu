// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/tile_loop scripts/ubench/tile_loop.cu
// What fraction of the FFMA2 rate does an fp32 register-tile inner loop reach when its operands come from shared memory?
// One CTA per SM, NT threads, per k-step each thread loads R float4 of A rows and C float4 of W columns (LDS.128) and issues
// (2R x 4C) FFMA2 = (4R x 4C) FMAs.  Reported: cycles per k-step per SM sub-partition vs the FFMA2 issue floor
// (warps per sub-partition x 2R x 4C x 2 cycles).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 lds4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int R, int C, int NT, bool LOADS, bool MATH, bool PREFETCH>
__global__ void __launch_bounds__(NT, 1) k_tile(float* out, int iters, long long* cycles) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 16384; i += NT) sm[i] = 1.0f + (float)(i & 7) * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    // A: 8 distinct row chunks per warp (lane & 7), W: 4 distinct column chunks (lane >> 3); rows of 132 floats
    const unsigned abase = sbase + (unsigned)((lane & 7) * 16 + (warp & 1) * 128);
    const unsigned wbase = sbase + 32768u + (unsigned)((lane >> 3) * 16 + (warp >> 1) * 64);
    float2 acc[2 * R][4 * C];
#pragma unroll
    for (int r = 0; r < 2 * R; r++)
#pragma unroll
        for (int c = 0; c < 4 * C; c++) acc[r][c] = make_float2(0.f, 0.f);
    float4 a[R], w[C], an[R], wn[C];
#pragma unroll
    for (int r = 0; r < R; r++) a[r] = an[r] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
    for (int c = 0; c < C; c++) w[c] = wn[c] = make_float4(1.f, 1.f, 1.f, 1.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            if (LOADS) {
                if (PREFETCH) {
#pragma unroll
                    for (int r = 0; r < R; r++) { a[r] = an[r]; an[r] = lds4(abase + (unsigned)(k * 528 + r * 256)); }
#pragma unroll
                    for (int c = 0; c < C; c++) { w[c] = wn[c]; wn[c] = lds4(wbase + (unsigned)(k * 272 + c * 1024)); }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) a[r] = lds4(abase + (unsigned)(k * 528 + r * 256));
#pragma unroll
                    for (int c = 0; c < C; c++) w[c] = lds4(wbase + (unsigned)(k * 272 + c * 1024));
                }
            }
            if (MATH) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float2 p0 = make_float2(a[r].x, a[r].y), p1 = make_float2(a[r].z, a[r].w);
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const float wv[4] = {w[c].x, w[c].y, w[c].z, w[c].w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            acc[2 * r][4 * c + q] = __ffma2_rn(p0, make_float2(wv[q], wv[q]), acc[2 * r][4 * c + q]);
                            acc[2 * r + 1][4 * c + q] = __ffma2_rn(p1, make_float2(wv[q], wv[q]), acc[2 * r + 1][4 * c + q]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) acc[0][0].x += a[r].x + a[r].y + a[r].z + a[r].w;
#pragma unroll
                for (int c = 0; c < C; c++) acc[0][1].x += w[c].x + w[c].y + w[c].z + w[c].w;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 2 * R; r++)
#pragma unroll
        for (int c = 0; c < 4 * C; c++) s += acc[r][c].x + acc[r][c].y;
    out[blockIdx.x * NT + threadIdx.x] = s;
}

template <int R, int C, int NT, bool LOADS, bool MATH, bool PREFETCH>
void run(const char* name, float* out, long long* cyc) {
    const int iters = 400;
    for (int rep = 0; rep < 2; rep++) {
        k_tile<R, C, NT, LOADS, MATH, PREFETCH><<<148, NT, 65536>>>(out, iters, cyc);
        cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_k = (double)h[0] / ((double)iters * 32);
    const double floor_ = (NT / 128.0) * (2 * R) * (4 * C) * 2.0;    // warps per sub-partition x FFMA2 per k-step x 2 cycles
    printf("%-44s %7.1f cycles per k-step, FFMA2 floor %6.1f -> %5.1f %% of the FFMA2 rate\n", name, per_k, floor_, 100.0 * floor_ / per_k);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
#define CFG(R, C, NT, L, M, P) cudaFuncSetAttribute(k_tile<R, C, NT, L, M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)
    CFG(2, 1, 512, false, true, false); CFG(2, 1, 512, true, false, false); CFG(2, 1, 512, true, true, false); CFG(2, 1, 512, true, true, true);
    CFG(2, 2, 512, true, true, false); CFG(2, 2, 256, true, true, false); CFG(2, 2, 256, true, true, true); CFG(4, 1, 512, true, true, false);
    CFG(4, 1, 256, true, true, false); CFG(1, 2, 512, true, true, false); CFG(4, 2, 256, true, true, false); CFG(2, 1, 256, true, true, false);
    CFG(2, 1, 384, true, true, false); CFG(2, 2, 384, true, true, false); CFG(2, 2, 128, true, true, false); CFG(2, 2, 128, true, true, true);
    run<2, 1, 512, false, true, false>("8x4 tile, 512 thr, math only", out, cyc);
    run<2, 1, 512, true, false, false>("8x4 tile, 512 thr, loads only", out, cyc);
    run<2, 1, 512, true, true, false>("8x4 tile, 512 thr, loads + math", out, cyc);
    run<2, 1, 512, true, true, true>("8x4 tile, 512 thr, loads + math, prefetch", out, cyc);
    run<2, 1, 384, true, true, false>("8x4 tile, 384 thr, loads + math", out, cyc);
    run<2, 1, 256, true, true, false>("8x4 tile, 256 thr, loads + math", out, cyc);
    run<2, 2, 512, true, true, false>("8x8 tile, 512 thr, loads + math", out, cyc);
    run<2, 2, 384, true, true, false>("8x8 tile, 384 thr, loads + math", out, cyc);
    run<2, 2, 256, true, true, false>("8x8 tile, 256 thr, loads + math", out, cyc);
    run<2, 2, 256, true, true, true>("8x8 tile, 256 thr, loads + math, prefetch", out, cyc);
    run<2, 2, 128, true, true, false>("8x8 tile, 128 thr, loads + math", out, cyc);
    run<2, 2, 128, true, true, true>("8x8 tile, 128 thr, loads + math, prefetch", out, cyc);
    run<4, 1, 512, true, true, false>("16x4 tile, 512 thr, loads + math", out, cyc);
    run<4, 1, 256, true, true, false>("16x4 tile, 256 thr, loads + math", out, cyc);
    run<1, 2, 512, true, true, false>("4x8 tile, 512 thr, loads + math", out, cyc);
    run<4, 2, 256, true, true, false>("16x8 tile, 256 thr, loads + math", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
