// gemm.cu — fp32 dense layer kernels, generic over layer shapes.
//
// Replaces mat_mul_cuda / mat_mul_backwards_cuda (reference src/mat_mul.cu:132-217: add_bias kernel +
// three cublasSgemm calls) and the separate ReLU / ReLU' / bias-gradient kernels
// (src/activation_function.cu:17-43, src/neural_network.cu:108-118) with three fused SIMT kernels:
//
//   forward        y  = act(x . W^T + b)                     bias + activation in the epilogue
//   backward-input gx = (g . W) * act'(x_in)                 previous layer's activation mask fused
//   backward-param dW = g^T . x  (reduction over the batch)  split-K over the batch: every split
//                  db = column sums of g                      writes its own slab; slabs are summed in a
//                                                             fixed order later (fused into Adam), so
//                                                             gradients are run-to-run deterministic.
//
// fp32 FFMA throughout: the north-star tolerance for these nets is 1e-5, which rules out
// single-pass TF32/bf16 tensor-core math here (the wide-MLP tensor path is separate).
// One templated kernel covers the three contractions: 64x64x16 tiles, 256 threads, 4x4 register
// micro-tiles, k-major shared tiles (conflict-free 128-bit LDS), register-staged double buffering.
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kBM = 64, kBN = 64, kBK = 16, kPad = 4;

enum GemmMode { kFwd = 0, kBwdInput = 1, kBwdParam = 2 };

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int M, N, K;
    int lda, ldb, ldc;
    const float* bias;   // kFwd
    const float* xin;    // kBwdInput: post-activation input of the layer, [M][N]
    int act;
    int k_per_split;     // kBwdParam
    size_t c_split_stride;
};

// A(i,k): kFwd/kBwdInput -> A[i*lda+k] ; kBwdParam -> A[k*lda+i]
// B(k,j): kFwd -> B[j*ldb+k] ; kBwdInput/kBwdParam -> B[k*ldb+j]
template <int MODE>
__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs p) {
    __shared__ __align__(16) float As[2][kBK][kBM + kPad];
    __shared__ __align__(16) float Bs[2][kBK][kBN + kPad];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    int kbeg = 0, kend = p.K;
    if (MODE == kBwdParam) {
        kbeg = blockIdx.z * p.k_per_split;
        kend = min(p.K, kbeg + p.k_per_split);
    }
    constexpr bool A_KCONTIG = (MODE != kBwdParam);
    constexpr bool B_KCONTIG = (MODE == kFwd);

    float ra[4], rb[4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = tid + 256 * q;
            {
                const int kk = A_KCONTIG ? (e & 15) : (e >> 6);
                const int ii = A_KCONTIG ? (e >> 4) : (e & 63);
                const int gi = m0 + ii, gk = k0 + kk;
                const bool ok = gi < p.M && gk < kend;
                const size_t off = A_KCONTIG ? (size_t)gi * p.lda + gk : (size_t)gk * p.lda + gi;
                ra[q] = ok ? __ldg(p.A + off) : 0.f;
            }
            {
                const int kk = B_KCONTIG ? (e & 15) : (e >> 6);
                const int jj = B_KCONTIG ? (e >> 4) : (e & 63);
                const int gj = n0 + jj, gk = k0 + kk;
                const bool ok = gj < p.N && gk < kend;
                const size_t off = B_KCONTIG ? (size_t)gj * p.ldb + gk : (size_t)gk * p.ldb + gj;
                rb[q] = ok ? __ldg(p.B + off) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = tid + 256 * q;
            As[buf][A_KCONTIG ? (e & 15) : (e >> 6)][A_KCONTIG ? (e >> 4) : (e & 63)] = ra[q];
            Bs[buf][B_KCONTIG ? (e & 15) : (e >> 6)][B_KCONTIG ? (e >> 4) : (e & 63)] = rb[q];
        }
    };

    float2 acc2[4][2];                       // [row][column pair]: packed FFMA2 (bit-identical to scalar fma.rn per element)
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) acc2[r][c] = make_float2(0.f, 0.f);

    const int ntiles = (kend - kbeg + kBK - 1) / kBK;
    if (ntiles > 0) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_tiles(kbeg + (t + 1) * kBK);
#pragma unroll
        for (int k = 0; k < kBK; k++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float2 bp[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) acc2[r][c] = __ffma2_rn(make_float2(av[r], av[r]), bp[c], acc2[r][c]);
        }
        if (t + 1 < ntiles) store_tiles(buf ^ 1);
        __syncthreads();
    }
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++) { acc[r][0] = acc2[r][0].x; acc[r][1] = acc2[r][0].y; acc[r][2] = acc2[r][1].x; acc[r][3] = acc2[r][1].y; }

    float* C = p.C + (MODE == kBwdParam ? (size_t)blockIdx.z * p.c_split_stride : 0);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int gi = m0 + ty * 4 + r;
        if (gi >= p.M) continue;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int gj = n0 + tx * 4 + c;
            if (gj >= p.N) continue;
            float val = acc[r][c];
            if (MODE == kFwd) {
                val = act_apply(val + p.bias[gj], p.act);
            } else if (MODE == kBwdInput) {
                if (p.act != kActNone) val = act_grad(p.xin[(size_t)gi * p.N + gj], val, p.act);
            }
            C[(size_t)gi * p.ldc + gj] = val;
        }
    }
}

// ---- 128x128x16 tile, 8x8 outputs per thread, 128-bit global loads, packed FFMA2 -------------------------------
// The large-batch path of the fp32 layers (C3: 2x256 at minibatch 65536).  Same operand conventions and epilogues as
// sgemm_kernel; requires 16-byte aligned operands with leading dimensions that are multiples of 4 (host-checked).
// Per k each thread issues 4 LDS.128 (8 row values + 8 column values) for 32 FFMA2 = 64 FMAs: 4 FMAs per loaded
// float, the ratio at which the 128 B/clk shared-memory pipe no longer bounds the FFMA pipe (see fused_mlp.cu).
// Thread (tx, ty): columns {4tx..4tx+3} U {64+4tx..}, rows {4ty..4ty+3} U {64+4ty..}.
constexpr int kTM = 128, kTN = 128, kTK = 16, kTPad = 4;
__device__ __forceinline__ float2 ffma2g(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int MODE>
__global__ void __launch_bounds__(256, 2) sgemm128_kernel(const GemmArgs p) {
    __shared__ __align__(16) float As[2][kTK][kTM + kTPad];
    __shared__ __align__(16) float Bs[2][kTK][kTN + kTPad];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
    int kbeg = 0, kend = p.K;
    if (MODE == kBwdParam) {
        kbeg = blockIdx.z * p.k_per_split;
        kend = min(p.K, kbeg + p.k_per_split);
    }
    constexpr bool A_KCONTIG = (MODE != kBwdParam);
    constexpr bool B_KCONTIG = (MODE == kFwd);

    float4 ra[2], rb[2];
    // one operand tile = 512 float4 chunks, two per thread
    auto load_op = [&](const float* base, int ld, int o0, int olim, bool kcontig, int k0, float4 (&r)[2]) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int e = tid + 256 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kcontig) {                       // [o][k]: chunk = (o = e >> 2, k offset 4 * (e & 3))
                const int o = o0 + (e >> 2), gk = k0 + 4 * (e & 3);
                if (o < olim) {
                    const float* src = base + (size_t)o * ld + gk;
                    if (gk + 3 < kend) v = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        if (gk + 0 < kend) v.x = src[0];
                        if (gk + 1 < kend) v.y = src[1];
                        if (gk + 2 < kend) v.z = src[2];
                    }
                }
            } else {                             // [k][o]: chunk = (k = e >> 5, o offset 4 * (e & 31))
                const int gk = k0 + (e >> 5), o = o0 + 4 * (e & 31);
                if (gk < kend) {
                    const float* src = base + (size_t)gk * ld + o;
                    if (o + 3 < olim) v = __ldg(reinterpret_cast<const float4*>(src));
                    else {
                        if (o + 0 < olim) v.x = src[0];
                        if (o + 1 < olim) v.y = src[1];
                        if (o + 2 < olim) v.z = src[2];
                    }
                }
            }
            r[q] = v;
        }
    };
    auto store_op = [&](float (*S)[kTM + kTPad], bool kcontig, const float4 (&r)[2]) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int e = tid + 256 * q;
            if (kcontig) {
                const int o = e >> 2, kk = 4 * (e & 3);
                S[kk + 0][o] = r[q].x; S[kk + 1][o] = r[q].y; S[kk + 2][o] = r[q].z; S[kk + 3][o] = r[q].w;
            } else {
                *reinterpret_cast<float4*>(&S[e >> 5][4 * (e & 31)]) = r[q];
            }
        }
    };

    float2 acc[8][4];                            // [row][column pair]
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f);

    const int ntiles = (kend - kbeg + kTK - 1) / kTK;
    if (ntiles > 0) {
        load_op(p.A, p.lda, m0, p.M, A_KCONTIG, kbeg, ra);
        load_op(p.B, p.ldb, n0, p.N, B_KCONTIG, kbeg, rb);
        store_op(As[0], A_KCONTIG, ra);
        store_op(Bs[0], B_KCONTIG, rb);
    }
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            load_op(p.A, p.lda, m0, p.M, A_KCONTIG, kbeg + (t + 1) * kTK, ra);
            load_op(p.B, p.ldb, n0, p.N, B_KCONTIG, kbeg + (t + 1) * kTK, rb);
        }
#pragma unroll
        for (int k = 0; k < kTK; k++) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * ty]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + 4 * tx]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = ffma2g(make_float2(av[r], av[r]), bp[c], acc[r][c]);
        }
        if (t + 1 < ntiles) {
            store_op(As[buf ^ 1], A_KCONTIG, ra);
            store_op(Bs[buf ^ 1], B_KCONTIG, rb);
        }
        __syncthreads();
    }

    float* C = p.C + (MODE == kBwdParam ? (size_t)blockIdx.z * p.c_split_stride : 0);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int gi = m0 + (r < 4 ? 4 * ty + r : 64 + 4 * ty + (r - 4));
        if (gi >= p.M) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int gj = n0 + 64 * h + 4 * tx;
            float v[4] = {acc[r][2 * h].x, acc[r][2 * h].y, acc[r][2 * h + 1].x, acc[r][2 * h + 1].y};
            if (gj + 3 < p.N && (p.ldc & 3) == 0) {
                if (MODE == kFwd) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gj));
                    v[0] = act_apply(v[0] + b.x, p.act); v[1] = act_apply(v[1] + b.y, p.act);
                    v[2] = act_apply(v[2] + b.z, p.act); v[3] = act_apply(v[3] + b.w, p.act);
                } else if (MODE == kBwdInput && p.act != kActNone) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(p.xin + (size_t)gi * p.N + gj));
                    v[0] = act_grad(x.x, v[0], p.act); v[1] = act_grad(x.y, v[1], p.act);
                    v[2] = act_grad(x.z, v[2], p.act); v[3] = act_grad(x.w, v[3], p.act);
                }
                *reinterpret_cast<float4*>(C + (size_t)gi * p.ldc + gj) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (gj + c >= p.N) continue;
                    float val = v[c];
                    if (MODE == kFwd) val = act_apply(val + p.bias[gj + c], p.act);
                    else if (MODE == kBwdInput && p.act != kActNone) val = act_grad(p.xin[(size_t)gi * p.N + gj + c], val, p.act);
                    C[(size_t)gi * p.ldc + gj + c] = val;
                }
            }
        }
    }
}

static bool big_tile_ok(const GemmArgs& a, int splits = 1) {
    auto al16 = [](const void* q) { return (((uintptr_t)q) & 15) == 0; };
    // 128x128 tiles pay off only when they still fill the machine (at least one CTA per SM)
    if ((long long)div_up(a.M, kTM) * div_up(a.N, kTN) * splits < num_sms()) return false;
    return a.M >= 128 && a.N >= 128 && a.K >= 64 && (a.lda & 3) == 0 && (a.ldb & 3) == 0 && al16(a.A) && al16(a.B) &&
           (a.xin == nullptr || (al16(a.xin) && (a.N & 3) == 0)) && (a.bias == nullptr || al16(a.bias));
}

// Small-K dense layer: out[m][N] = epi(sum_{k < K} in[m][k] * Wk[k][N]), K <= 32 (first layer of a low-dimensional
// env: K = S; dX of a <= 8 wide head: K = l).  Output-bandwidth bound (the 64x64 tile kernel spends its time on a
// K padded to 16/32 and scalar stores).  CTA = 32 rows x 256 columns; thread = 8 rows x 4 columns; the K x 256 weight
// slice is staged in shared memory (k-major), the 32 x K input block too (broadcast reads), stores are 128-bit.
//   TRANS_W = true : Wk[k][j] = W[j * K + k]   (forward, W is [N][K] row-major)      epilogue: bias + activation
//   TRANS_W = false: Wk[k][j] = W[k * N + j]   (dX, W is [K = l][N = n] row-major)   epilogue: activation' of xin
template <bool TRANS_W>
__global__ void __launch_bounds__(256)
smallk_linear_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ W,
                     const float* __restrict__ bias, const float* __restrict__ xin, int m, int K, int N, int act) {
    __shared__ __align__(16) float Ws[32][256 + 4];
    __shared__ float Xs[32][33];
    const int tid = threadIdx.x;
    constexpr int kRowsPerCta = 256;                 // 8 blocks of 32 rows share one staged weight slice
    const int n0 = blockIdx.x * 256;
    for (int e = tid; e < K * 256; e += 256) {
        const int k = TRANS_W ? (e % K) : (e >> 8), j = TRANS_W ? (e / K) : (e & 255);
        const int gj = n0 + j;
        Ws[k][j] = gj < N ? (TRANS_W ? W[(size_t)gj * K + k] : W[(size_t)k * N + gj]) : 0.f;
    }
    const int tc = tid & 63, tr = tid >> 6;          // columns 4tc..4tc+3, rows tr*8..tr*8+7
    const int gj = n0 + 4 * tc;
    const bool vec = gj + 3 < N && (N & 3) == 0;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (TRANS_W && vec) b4 = __ldg(reinterpret_cast<const float4*>(bias + gj));
    for (int m0 = blockIdx.y * kRowsPerCta; m0 < min(m, (int)(blockIdx.y + 1) * kRowsPerCta); m0 += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * K; e += 256) {
            const int r = e / K, k = e - r * K;
            Xs[r][k] = (m0 + r < m) ? in[(size_t)(m0 + r) * K + k] : 0.f;
        }
        __syncthreads();
        float2 acc[8][2];
#pragma unroll
        for (int r = 0; r < 8; r++) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
        for (int k = 0; k < K; k++) {
            const float4 w = *reinterpret_cast<const float4*>(&Ws[k][4 * tc]);
            const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float x = Xs[tr * 8 + r][k];
                acc[r][0] = __ffma2_rn(make_float2(x, x), w01, acc[r][0]);
                acc[r][1] = __ffma2_rn(make_float2(x, x), w23, acc[r][1]);
            }
        }
        if (gj >= N) continue;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int gi = m0 + tr * 8 + r;
            if (gi >= m) continue;
            float v[4] = {acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y};
            if (vec) {
                if (TRANS_W) {
                    v[0] = act_apply(v[0] + b4.x, act); v[1] = act_apply(v[1] + b4.y, act);
                    v[2] = act_apply(v[2] + b4.z, act); v[3] = act_apply(v[3] + b4.w, act);
                } else if (act != kActNone) {
                    const float4 h = __ldg(reinterpret_cast<const float4*>(xin + (size_t)gi * N + gj));
                    v[0] = act_grad(h.x, v[0], act); v[1] = act_grad(h.y, v[1], act);
                    v[2] = act_grad(h.z, v[2], act); v[3] = act_grad(h.w, v[3], act);
                }
                *reinterpret_cast<float4*>(out + (size_t)gi * N + gj) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (gj + c >= N) continue;
                    float val = v[c];
                    if (TRANS_W) val = act_apply(val + bias[gj + c], act);
                    else if (act != kActNone) val = act_grad(xin[(size_t)gi * N + gj + c], val, act);
                    out[(size_t)gi * N + gj + c] = val;
                }
            }
        }
    }
}

// Skinny forward for l <= 8 outputs (value head l=1, action heads): one warp per row, lanes split k.
__global__ void __launch_bounds__(256)
linear_forward_skinny_kernel(float* __restrict__ y, const float* __restrict__ x, const float* __restrict__ W,
                             const float* __restrict__ b, int m, int n, int l, int act) {
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < m; row += warps_total) {
        const float* xr = x + (size_t)row * n;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float xv = xr[k];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < l) acc[j] = fmaf(xv, __ldg(W + (size_t)j * n + k), acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < l) {
                const float s = warp_sum(acc[j]);
                if (lane == 0) y[(size_t)row * l + j] = act_apply(s + b[j], act);
            }
    }
}

// db slabs: gb_part[s][col] = sum over rows of split s of g[row][col].
// Thread = (float4 column group, row lane): 8 column groups x 32 row lanes, four independent 128-bit loads in
// flight per thread; fixed-order combine of the row lanes.
__global__ void __launch_bounds__(256)
colsum_kernel(float* __restrict__ gb_part, size_t stride, const float* __restrict__ g, int m, int l, int rows_per_split) {
    __shared__ float4 red[32][9];
    const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
    const int col = blockIdx.x * 32 + 4 * cx;
    const int r0 = blockIdx.y * rows_per_split, r1 = min(m, r0 + rows_per_split);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool vec = (l & 3) == 0 && col + 3 < l && (((uintptr_t)g) & 15) == 0;
    if (vec) {
        int r = r0 + ry;
        for (; r + 96 < r1; r += 128) {
            float4 t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = __ldg(reinterpret_cast<const float4*>(g + (size_t)(r + 32 * u) * l + col));
#pragma unroll
            for (int u = 0; u < 4; u++) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
        }
        for (; r < r1; r += 32) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * l + col));
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
    } else {
        for (int r = r0 + ry; r < r1; r += 32) {
            if (col + 0 < l) s.x += g[(size_t)r * l + col + 0];
            if (col + 1 < l) s.y += g[(size_t)r * l + col + 1];
            if (col + 2 < l) s.z += g[(size_t)r * l + col + 2];
            if (col + 3 < l) s.w += g[(size_t)r * l + col + 3];
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0) {
        float4 t = red[0][cx];
#pragma unroll 8
        for (int k = 1; k < 32; k++) { const float4 q = red[k][cx]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
        float* out = gb_part + (size_t)blockIdx.y * stride + col;
        if (col + 0 < l) out[0] = t.x;
        if (col + 1 < l) out[1] = t.y;
        if (col + 2 < l) out[2] = t.z;
        if (col + 3 < l) out[3] = t.w;
    }
}

// Skinny dW (+db): one side of the layer is <= 32 wide (first layer of a low-dimensional env: n = S; value / action
// heads: l <= 8).  gW[j][k] = sum_r g[r][j] * x[r][k].  The WIDE side is spread over threads (4 consecutive columns
// per thread, 256 per CTA, 128-bit coalesced loads), the SMALL side is staged in shared memory and broadcast; each
// broadcast LDS.128 feeds 16 FFMAs (4 columns x 4 small values) - with one column per thread the kernel was bound by
// shared-memory wavefronts at 9% of the FFMA rate.  Rows are split over 4 row lanes per CTA, `splits` slabs and
// gridDim.z extra row splits whose partials land in scratch and are folded in order by skinny_dw_fold_kernel.
//   WIDE_IS_L = true : wide = g columns (l), small = x columns (n <= 32); also emits db (column sums of g)
//   WIDE_IS_L = false: wide = x columns (n), small = g columns (l <= 32)
template <int SP, bool WIDE_IS_L>     // SP = padded small width (multiple of 4)
__global__ void __launch_bounds__(256)
skinny_dw_kernel(float* __restrict__ part, const float* __restrict__ g, const float* __restrict__ x, int m, int n, int l,
                 int rows_per_split, float* __restrict__ gW_part, float* __restrict__ gb_part, size_t stride) {
    constexpr int CH = 32;                                   // rows staged per chunk
    __shared__ __align__(16) float sm[CH][SP];
    __shared__ float red[64][4 * SP + 4];
    const int cx = threadIdx.x & 63, q = threadIdx.x >> 6;   // q = row lane 0..3
    const int wide_n = WIDE_IS_L ? l : n, small_n = WIDE_IS_L ? n : l;
    const float* wide = WIDE_IS_L ? g : x;
    const float* small = WIDE_IS_L ? x : g;
    const int col = blockIdx.x * 256 + 4 * cx;
    const bool vec = (wide_n & 3) == 0 && col + 3 < wide_n && (((uintptr_t)wide) & 15) == 0;
    const int R = gridDim.z;
    const int sub = (rows_per_split + R - 1) / R;
    const int s0 = blockIdx.y * rows_per_split;
    const int r0 = s0 + blockIdx.z * sub, r1 = min(min(m, s0 + rows_per_split), r0 + sub);
    float acc[4][SP], bsum[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        bsum[c] = 0.f;
#pragma unroll
        for (int k = 0; k < SP; k++) acc[c][k] = 0.f;
    }
    for (int c0 = r0; c0 < r1; c0 += CH) {
        const int rows = min(CH, r1 - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < CH * SP; e += 256) {
            const int r = e / SP, k = e - r * SP;
            sm[r][k] = (r < rows && k < small_n) ? small[(size_t)(c0 + r) * small_n + k] : 0.f;
        }
        __syncthreads();
        float4 wv[CH / 4];
#pragma unroll
        for (int i = 0; i < CH / 4; i++) {
            const int r = 4 * i + q;
            wv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < rows) {
                const float* wp = wide + (size_t)(c0 + r) * wide_n + col;
                if (vec) wv[i] = __ldg(reinterpret_cast<const float4*>(wp));
                else {
                    if (col + 0 < wide_n) wv[i].x = wp[0];
                    if (col + 1 < wide_n) wv[i].y = wp[1];
                    if (col + 2 < wide_n) wv[i].z = wp[2];
                    if (col + 3 < wide_n) wv[i].w = wp[3];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < CH / 4; i++) {
            const int r = 4 * i + q;
            const float w4[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
#pragma unroll
            for (int c = 0; c < 4; c++) bsum[c] += w4[c];
#pragma unroll
            for (int k4 = 0; k4 < SP / 4; k4++) {
                const float4 sv = *reinterpret_cast<const float4*>(&sm[r][4 * k4]);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    acc[c][4 * k4 + 0] = fmaf(w4[c], sv.x, acc[c][4 * k4 + 0]);
                    acc[c][4 * k4 + 1] = fmaf(w4[c], sv.y, acc[c][4 * k4 + 1]);
                    acc[c][4 * k4 + 2] = fmaf(w4[c], sv.z, acc[c][4 * k4 + 2]);
                    acc[c][4 * k4 + 3] = fmaf(w4[c], sv.w, acc[c][4 * k4 + 3]);
                }
            }
        }
    }
    // row lanes 1..3 hand their partials to lane 0 one after the other (fixed order, one staging buffer)
    for (int src = 1; src < 4; src++) {
        __syncthreads();
        if (q == src) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
#pragma unroll
                for (int k = 0; k < SP; k++) red[cx][c * SP + k] = acc[c][k];
                red[cx][4 * SP + c] = bsum[c];
            }
        }
        __syncthreads();
        if (q == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
#pragma unroll
                for (int k = 0; k < SP; k++) acc[c][k] += red[cx][c * SP + k];
                bsum[c] += red[cx][4 * SP + c];
            }
        }
    }
    if (q == 0 && R == 1) {
        // no extra row split: straight into the gradient slab
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (col + c >= wide_n) continue;
#pragma unroll
            for (int k = 0; k < SP; k++)
                if (k < small_n) {
                    if (WIDE_IS_L) gW_part[(size_t)blockIdx.y * stride + (size_t)(col + c) * n + k] = acc[c][k];
                    else gW_part[(size_t)blockIdx.y * stride + (size_t)k * n + col + c] = acc[c][k];
                }
            if (WIDE_IS_L) gb_part[(size_t)blockIdx.y * stride + col + c] = bsum[c];
        }
    } else if (q == 0) {
        // partial layout: [slab y][z][wide_n][SP + 1]  (last entry of a row = column sum of the wide array)
        float* out = part + ((size_t)(blockIdx.y * R + blockIdx.z) * wide_n) * (SP + 1);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (col + c >= wide_n) continue;
            float* o = out + (size_t)(col + c) * (SP + 1);
#pragma unroll
            for (int k = 0; k < SP; k++) o[k] = acc[c][k];
            o[SP] = bsum[c];
        }
    }
}

// Fold the R row-split partials of every slab in order and scatter into the gradient slab layout.
__global__ void __launch_bounds__(256)
skinny_dw_fold_kernel(float* __restrict__ gW_part, float* __restrict__ gb_part, size_t stride, const float* __restrict__ part,
                      int R, int wide_n, int small_n, int sp1, int n, int wide_is_l) {
    const int slab = blockIdx.y;
    const int total = wide_n * sp1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int col = e / sp1, k = e - col * sp1;
        float t = 0.f;
        for (int z = 0; z < R; z++) t += part[((size_t)(slab * R + z) * wide_n + col) * sp1 + k];
        // wide_is_l: 1 = first layer (wide = g columns; the last entry is this layer's db), 0 = head (wide = input columns),
        // 2 = head whose last entry holds the column sums of its dX = the db of the layer below (gb_part points there)
        if (k < small_n) {
            if (wide_is_l == 1) gW_part[(size_t)slab * stride + (size_t)col * n + k] = t;    // gW[j = col][k]
            else gW_part[(size_t)slab * stride + (size_t)k * n + col] = t;                   // gW[j = k][col]
        } else if (k == sp1 - 1 && wide_is_l != 0) {
            gb_part[(size_t)slab * stride + col] = t;
        }
    }
}

__global__ void __launch_bounds__(256) act_kernel(float* __restrict__ x, long long count, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) x[i] = act_apply(x[i], act);
}
__global__ void __launch_bounds__(256) act_grad_kernel(const float* __restrict__ y, float* __restrict__ g, long long count, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) g[i] = act_grad(y[i], g[i], act);
}

// ---- host wrappers ----------------------------------------------------------------------------------
int choose_splits(int m, size_t param_count) {
    long long s = div_up(m, 256);
    s = std::max<long long>(1, std::min<long long>(s, 64));
    const long long cap = std::max<long long>(1, (long long)((256ull << 20) / (std::max<size_t>(param_count, 1) * sizeof(float))));
    return (int)std::min(s, cap);
}

static int g_matmul_precision = -1;
int matmul_precision() {
    if (g_matmul_precision < 0) { const char* e = getenv("PPO_B200_TF32"); g_matmul_precision = (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : (e && e[0] == '3') ? 3 : 0; }
    return g_matmul_precision;
}
static bool use_tc(int m, int n, int l, const void* a, const void* b, int lda, int ldb) {
    return matmul_precision() == 1 && m >= 128 && n >= 64 && l >= 64 && tc_shape_ok(a, b, lda, ldb);
}

void linear_forward(float* y, const float* x, const float* W, const float* b, int m, int n, int l, int act) {
    if (m <= 0) return;
    if (use_tc(m, n, l, x, W, n, n) && (l % 4) == 0) { tc_linear_forward(y, x, W, b, m, n, l, act); return; }
    if (l <= 8 && narrow_head_forward(y, x, W, b, m, n, l, act)) return;
    if (n <= 32 && l >= 64 && narrow_first_forward(y, nullptr, x, W, b, m, n, l, act)) return;
    if (l <= 8) {
        const int blocks = std::min(div_up(m, 8), num_sms() * 8);
        B200_LAUNCH(linear_forward_skinny_kernel, blocks, 256, 0, y, x, W, b, m, n, l, act);
        return;
    }
    if (n <= 32 && l >= 64 && m >= 256) {
        dim3 gk(div_up(l, 256), div_up(m, 256), 1);
        B200_LAUNCH(smallk_linear_kernel<true>, gk, 256, 0, y, x, W, b, (const float*)nullptr, m, n, l, act);
        return;
    }
    GemmArgs a{};
    a.A = x; a.B = W; a.C = y; a.M = m; a.N = l; a.K = n; a.lda = n; a.ldb = n; a.ldc = l;
    a.bias = b; a.act = act;
    if (big_tile_ok(a)) {
        dim3 gridb(div_up(l, kTN), div_up(m, kTM), 1);
        B200_LAUNCH(sgemm128_kernel<kFwd>, gridb, 256, 0, a);
        return;
    }
    dim3 grid(div_up(l, kBN), div_up(m, kBM), 1);
    B200_LAUNCH(sgemm_kernel<kFwd>, grid, 256, 0, a);
}

void linear_backward_input(float* gx, const float* g, const float* W, const float* xin, int m, int n, int l, int act_prev) {
    if (m <= 0) return;
    if (use_tc(m, n, l, g, W, l, n)) { tc_linear_backward_input(gx, g, W, xin, m, n, l, act_prev); return; }
    if (l <= 32 && n >= 64 && m >= 256) {
        dim3 gk(div_up(n, 256), div_up(m, 256), 1);
        B200_LAUNCH(smallk_linear_kernel<false>, gk, 256, 0, gx, g, W, (const float*)nullptr, xin, m, l, n, act_prev);
        return;
    }
    GemmArgs a{};
    a.A = g; a.B = W; a.C = gx; a.M = m; a.N = n; a.K = l; a.lda = l; a.ldb = n; a.ldc = n;
    a.xin = xin; a.act = act_prev;
    if (big_tile_ok(a)) {
        dim3 gridb(div_up(n, kTN), div_up(m, kTM), 1);
        B200_LAUNCH(sgemm128_kernel<kBwdInput>, gridb, 256, 0, a);
        return;
    }
    dim3 grid(div_up(n, kBN), div_up(m, kBM), 1);
    B200_LAUNCH(sgemm_kernel<kBwdInput>, grid, 256, 0, a);
}

// db slabs of an l <= 8 wide head: one CTA per slab, thread = row lane (256 of them, four rows in flight each), every thread sums
// all l columns of its rows; fixed-order combine through shared memory.  (colsum_kernel keeps 3/4 of its threads idle at l = 6
// and walks 200 dependent iterations: 58 us for a 1.5 MB array.)
__global__ void __launch_bounds__(256)
colsum_narrow_kernel(float* __restrict__ gb_part, size_t stride, const float* __restrict__ g, int m, int l, int rows_per_split) {
    __shared__ float red[256][9];
    const int r0 = blockIdx.x * rows_per_split, r1 = min(m, r0 + rows_per_split);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = 0.f;
    int r = r0 + threadIdx.x;
    for (; r + 3 * 256 < r1; r += 4 * 256) {
        float t[4][8];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) t[u][j] = j < l ? __ldg(g + (size_t)(r + 256 * u) * l + j) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) s[j] += t[u][j];
    }
    for (; r < r1; r += 256)
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < l) s[j] += __ldg(g + (size_t)r * l + j);
#pragma unroll
    for (int j = 0; j < 8; j++) red[threadIdx.x][j] = s[j];
    __syncthreads();
    if (threadIdx.x < l) {
        float t = 0.f;
        for (int k = 0; k < 256; k++) t += red[k][threadIdx.x];
        gb_part[(size_t)blockIdx.x * stride + threadIdx.x] = t;
    }
}

// bias gradients of one layer as `splits` slabs: column sums of g over the rows of each split
void launch_colsum(float* gb_part, size_t stride, int splits, const float* g, int m, int l) {
    int rows = div_up(m, splits);
    rows = div_up(rows, 32) * 32;
    if (l <= 8) {
        B200_LAUNCH(colsum_narrow_kernel, splits, 256, 0, gb_part, stride, g, m, l, rows);
        return;
    }
    dim3 grid2(div_up(l, 32), splits, 1);
    B200_LAUNCH(colsum_kernel, grid2, 256, 0, gb_part, stride, g, m, l, rows);
}

void linear_backward_params(float* gW_part, float* gb_part, size_t stride, int splits, const float* g,
                            const float* x, int m, int n, int l, bool db_done) {
    if (m <= 0) return;
    int rows = div_up(m, splits);
    rows = div_up(rows, 32) * 32;          // multiple of both tile depths (16 FFMA, 32 TF32)
    if (use_tc(m, n, l, g, x, l, n) && ((stride * 4) % 16) == 0 && ((uintptr_t)gW_part & 15) == 0) {
        tc_linear_backward_weights(gW_part, stride, splits, g, x, m, n, l);
        if (!db_done) {       // the fused head kernel above this layer may already have written the db slabs (narrow.cu)
            dim3 grid2(div_up(l, 32), splits, 1);
            B200_LAUNCH(colsum_kernel, grid2, 256, 0, gb_part, stride, g, m, l, rows);
        }
        return;
    }
    const bool narrow_in = m >= 1024 && n <= 32 && l >= 64;      // first layer of a low-dimensional env: wide side = l (+ db)
    const bool narrow_out = m >= 1024 && l <= 8 && n >= 64;      // value / action heads: wide side = n
    // streaming kernels of narrow.cu (one cp.async-pipelined pass over the wide array); the kernels below stay for shapes they
    // do not take (wide side not a multiple of 4, unaligned bases) and for PPO_B200_NARROW=0 A/B runs
    if (narrow_in && narrow_first_layer_backward(gW_part, gb_part, stride, splits, g, x, m, n, l)) return;
    if (narrow_out && narrow_head_backward(gW_part, stride, splits, nullptr, nullptr, nullptr, g, x, nullptr, m, n, l, kActNone)) {
        launch_colsum(gb_part, stride, splits, g, m, l);
        return;
    }
    if (narrow_in || narrow_out) {
        const int wide_n = narrow_in ? l : n, small_n = narrow_in ? n : l;
        const int sp = narrow_in ? (n <= 8 ? 8 : n <= 16 ? 16 : n <= 24 ? 24 : 32) : 8;
        const int gx = div_up(wide_n, 256);
        const int R = m <= 16384 ? 1 : std::max(1, std::min(16, div_up(2 * num_sms(), gx * splits)));     // ~2 CTAs per SM
        // own slot: callers (mat_mul_backwards_cuda) keep their split-K slabs in kScratchPartials across this call
        float* part = static_cast<float*>(scratch(kScratchSkinny, (size_t)splits * R * wide_n * (sp + 1) * sizeof(float)));
        dim3 gs(gx, splits, R);
        if (narrow_in) {
            if (sp == 8) B200_LAUNCH((skinny_dw_kernel<8, true>), gs, 256, 0, part, g, x, m, n, l, rows, gW_part, gb_part, stride);
            else if (sp == 16) B200_LAUNCH((skinny_dw_kernel<16, true>), gs, 256, 0, part, g, x, m, n, l, rows, gW_part, gb_part, stride);
            else if (sp == 24) B200_LAUNCH((skinny_dw_kernel<24, true>), gs, 256, 0, part, g, x, m, n, l, rows, gW_part, gb_part, stride);
            else B200_LAUNCH((skinny_dw_kernel<32, true>), gs, 256, 0, part, g, x, m, n, l, rows, gW_part, gb_part, stride);
        } else {
            B200_LAUNCH((skinny_dw_kernel<8, false>), gs, 256, 0, part, g, x, m, n, l, rows, gW_part, gb_part, stride);
        }
        dim3 gf(std::max(1, std::min(32, div_up(wide_n * (sp + 1), 256))), splits, 1);
        if (R > 1) B200_LAUNCH(skinny_dw_fold_kernel, gf, 256, 0, gW_part, gb_part, stride, part, R, wide_n, small_n, sp + 1, n, narrow_in ? 1 : 0);
        if (narrow_out) {            // db of a <= 8 wide head: tiny column sum
            dim3 grid2(div_up(l, 32), splits, 1);
            B200_LAUNCH(colsum_kernel, grid2, 256, 0, gb_part, stride, g, m, l, rows);
        }
        return;
    }
    GemmArgs a{};
    a.A = g; a.B = x; a.C = gW_part; a.M = l; a.N = n; a.K = m; a.lda = l; a.ldb = n; a.ldc = n;
    a.k_per_split = rows; a.c_split_stride = stride;
    if (big_tile_ok(a, splits) && ((stride & 3) == 0) && ((((uintptr_t)gW_part) & 15) == 0)) {
        dim3 gridb(div_up(n, kTN), div_up(l, kTM), splits);
        B200_LAUNCH(sgemm128_kernel<kBwdParam>, gridb, 256, 0, a);
    } else {
        dim3 grid(div_up(n, kBN), div_up(l, kBM), splits);
        B200_LAUNCH(sgemm_kernel<kBwdParam>, grid, 256, 0, a);
    }
    dim3 grid2(div_up(l, 32), splits, 1);
    B200_LAUNCH(colsum_kernel, grid2, 256, 0, gb_part, stride, g, m, l, rows);
}

void skinny_fold(float* gW_part, float* gb_part, size_t stride, const float* part, int splits, int R, int wide_n, int small_n, int sp1, int n,
                 int wide_is_l) {
    dim3 gf(std::max(1, std::min(32, div_up(wide_n * sp1, 256))), splits, 1);
    B200_LAUNCH(skinny_dw_fold_kernel, gf, 256, 0, gW_part, gb_part, stride, part, R, wide_n, small_n, sp1, n, wide_is_l);
}

void activation_inplace(float* x, long long count, int act) {
    if (act == kActNone || count <= 0) return;
    const int blocks = (int)std::min<long long>(div_up(count, 256), (long long)num_sms() * 8);
    B200_LAUNCH(act_kernel, blocks, 256, 0, x, count, act);
}
void activation_grad_inplace(const float* y, float* grad, long long count, int act) {
    if (act == kActNone || count <= 0) return;
    const int blocks = (int)std::min<long long>(div_up(count, 256), (long long)num_sms() * 8);
    B200_LAUNCH(act_grad_kernel, blocks, 256, 0, y, grad, count, act);
}

}  // namespace b200

using namespace b200;

extern "C" {

void ppo_b200_set_matmul_precision(int mode) { g_matmul_precision = (mode >= 0 && mode <= 3) ? mode : 0; }

// include/mat_mul.h:19-20 (device pointers; handle ignored)
void mat_mul_cuda(cublasHandle_t handle, float* out, float* x, float* weight, float* bias, int m, int n, int l) {
    (void)handle;
    linear_forward(out, x, weight, bias, m, n, l, kActNone);
}

void mat_mul_backwards_cuda(cublasHandle_t handle, float* grad_x, float* grad_weight, float* grad_in,
                            float* x, float* weight, int m, int n, int l) {
    (void)handle;
    linear_backward_input(grad_x, grad_in, weight, nullptr, m, n, l, kActNone);
    const int splits = choose_splits(m, (size_t)n * l + l);
    const size_t stride = (size_t)n * l + l;
    float* part = static_cast<float*>(scratch(kScratchPartials, (size_t)splits * stride * sizeof(float)));
    linear_backward_params(part, part + (size_t)n * l, stride, splits, grad_in, x, m, n, l);
    reduce_partials(grad_weight, part, splits, stride, n * l);
}

// include/mat_mul.h:16-17 (host pointers, staged; still computed by the kernels above)
void mat_mul(float* out, float* x, float* weight, float* bias, int m, int n, int l) {
    HostStage st;
    st.add(out, (size_t)m * l * 4, false, true);
    st.add(x, (size_t)m * n * 4, true, false);
    st.add(weight, (size_t)l * n * 4, true, false);
    st.add(bias, (size_t)l * 4, true, false);
    st.upload();
    linear_forward(st.dev<float>(0), st.dev<float>(1), st.dev<float>(2), st.dev<float>(3), m, n, l, kActNone);
    st.download();
}

void mat_mul_backwards(float* grad_x, float* grad_weight, float* grad_in, float* x, float* weight, int m, int n, int l) {
    HostStage st;
    st.add(grad_x, (size_t)m * n * 4, false, true);
    st.add(grad_weight, (size_t)l * n * 4, false, true);
    st.add(grad_in, (size_t)m * l * 4, true, false);
    st.add(x, (size_t)m * n * 4, true, false);
    st.add(weight, (size_t)l * n * 4, true, false);
    st.upload();
    mat_mul_backwards_cuda(nullptr, st.dev<float>(0), st.dev<float>(1), st.dev<float>(2), st.dev<float>(3), st.dev<float>(4), m, n, l);
    st.download();
}

// include/activation_function.h:15-22
void ReLU_cuda(float* x, int m, int n) { activation_inplace(x, (long long)m * n, kActRelu); }
void ReLU_derivative_cuda(float* x, float* grad, int m, int n) { activation_grad_inplace(x, grad, (long long)m * n, kActRelu); }
void Tanh_cuda(float* x, int m, int n) { activation_inplace(x, (long long)m * n, kActTanh); }
void Tanh_derivative_cuda(float* x, float* grad, int m, int n) { activation_grad_inplace(x, grad, (long long)m * n, kActTanh); }

static void staged_act(float* x, float* grad, int m, int n, int act, bool deriv) {
    HostStage st;
    const size_t bytes = (size_t)m * n * 4;
    st.add(x, bytes, true, !deriv);
    if (deriv) st.add(grad, bytes, true, true);
    st.upload();
    if (deriv) activation_grad_inplace(st.dev<float>(0), st.dev<float>(1), (long long)m * n, act);
    else activation_inplace(st.dev<float>(0), (long long)m * n, act);
    st.download();
}
void ReLU(float* x, int m, int n) { staged_act(x, nullptr, m, n, kActRelu, false); }
void ReLU_derivative(float* x, float* grad, int m, int n) { staged_act(x, grad, m, n, kActRelu, true); }
static void Tanh_host(float* x, int m, int n) { staged_act(x, nullptr, m, n, kActTanh, false); }
static void Tanh_derivative_host(float* x, float* grad, int m, int n) { staged_act(x, grad, m, n, kActTanh, true); }

ActivationFunction* build_activation_function(char* name) {
    ActivationFunction* f = (ActivationFunction*)malloc(sizeof(ActivationFunction));
    const int a = act_code(name);
    f->activation = a == kActRelu ? &ReLU : a == kActTanh ? &Tanh_host : nullptr;
    f->activation_derivative = a == kActRelu ? &ReLU_derivative : a == kActTanh ? &Tanh_derivative_host : nullptr;
    return f;
}

ActivationFunction* build_activation_function_cuda(char* name) {
    ActivationFunction* f = (ActivationFunction*)malloc(sizeof(ActivationFunction));
    const int a = act_code(name);
    f->activation = a == kActRelu ? &ReLU_cuda : a == kActTanh ? &Tanh_cuda : nullptr;
    f->activation_derivative = a == kActRelu ? &ReLU_derivative_cuda : a == kActTanh ? &Tanh_derivative_cuda : nullptr;
    return f;
}

}  // extern "C"
