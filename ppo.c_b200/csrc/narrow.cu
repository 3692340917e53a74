// narrow.cu — streaming kernels for the NARROW layers of a wide net (the reference runs them as
// cublasSgemm like every other layer, src/mat_mul.cu:149-208).
//
// A 17 -> H first layer and an H -> 6 / H -> 1 head have one side of <= 32 columns: their contractions are a few hundred
// MFLOP but each one walks a [minibatch][H] activation or gradient array (67 MB at H = 256, 268 MB at H = 1024), so they
// are HBM-bound and every extra pass over such an array costs as much as a tensor-core GEMM of the wide layers.  The
// kernels here stream the wide array ONCE per stage, with the narrow companion broadcast from shared memory:
//
//   narrow_bwd_kernel<SP, false>  first layer backward:  dW0[j][k] = sum_r g[r][j] x[r][k],  db0[j] = sum_r g[r][j]
//   narrow_bwd_kernel<8,  true>   head backward, fused:  dW[j][k]  = sum_r g[r][j] h[r][k]            (j < l <= 8)
//                                                        gx[r][k]  = (sum_j g[r][j] W[j][k]) act'(h[r][k])   (+ its 3xTF32 lo part)
//                                 one pass over h instead of three (dW kernel, dX kernel, split pass)
//
// Thread = 4 consecutive wide columns x every narrow column (packed FFMA2 on pairs of narrow columns); CTA = 256 wide
// columns x 4 row lanes; 32-row chunks of both arrays arrive by cp.async into two shared-memory buffers (the next chunk is
// in flight while the current one is consumed; a thread reads back exactly the wide words it copied).  Row ranges follow the
// library's deterministic split-K scheme: `splits` slabs x R row splits per slab, partials [slab][z][wide column][SP + 1]
// folded in fixed order by skinny_dw_fold_kernel (gemm.cu).
#include "common.cuh"
#include "internal.h"

namespace b200 {

struct NarrowArgs {
    const float* wide;        // [m][wide_n]   first layer: g (gradient wrt the layer output);  head: h (layer input, post-activation)
    const float* small;       // [m][small_n]  first layer: x (layer input);                    head: g (gradient wrt the head output)
    int m, wide_n, small_n, rows_per_split;
    float* part;              // [slab][R][wide_n][SP + 1]
    const float* W;           // head: [small_n][wide_n]
    float* gx;                // head: [m][wide_n]
    float* gx_lo;             // head: optional lo companion of gx (3xTF32 mode)
    int act_prev;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nrw_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nrw_cp4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nrw_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void nrw_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kNrwCH = 32;        // rows per chunk
constexpr int kNrwCols = 256;     // wide columns per CTA

template <int SP, bool HEAD>
__global__ void __launch_bounds__(256, 2) narrow_bwd_kernel(const NarrowArgs p) {
    extern __shared__ __align__(16) float nsm[];
    float* wt = nsm;                                        // [2][CH][256]
    float* st = nsm + 2 * kNrwCH * kNrwCols;                // [2][CH][SP]
    float* red = wt;                                        // [64][4 * SP + 4]  row-lane hand-over (the tiles are dead by then)
    const int tid = threadIdx.x, cx = tid & 63, q = tid >> 6;
    const int col = blockIdx.x * kNrwCols + 4 * cx;
    const bool col_ok = col < p.wide_n;                     // wide_n is a multiple of 4 (host check)
    const int R = gridDim.z;
    const int sub = (p.rows_per_split + R - 1) / R;
    const int s0 = blockIdx.y * p.rows_per_split;
    const int r0 = s0 + blockIdx.z * sub, r1 = min(min(p.m, s0 + p.rows_per_split), r0 + sub);
    const int n_chunks = r1 > r0 ? (r1 - r0 + kNrwCH - 1) / kNrwCH : 0;

    // zero the narrow tiles once: columns [small_n, SP) are never written again
    for (int e = tid; e < 2 * kNrwCH * SP; e += 256) st[e] = 0.f;
    __syncthreads();

    auto issue = [&](int c) {
        const int b = c & 1, c0 = r0 + c * kNrwCH, rows = min(kNrwCH, r1 - c0);
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < kNrwCH / 4; i++) {
                const int r = 4 * i + q;
                if (r < rows) nrw_cp16(wt + (b * kNrwCH + r) * kNrwCols + 4 * cx, p.wide + (size_t)(c0 + r) * p.wide_n + col);
            }
        }
        for (int e = tid; e < rows * p.small_n; e += 256) {
            const int r = e / p.small_n, k = e - r * p.small_n;
            nrw_cp4(st + (b * kNrwCH + r) * SP + k, p.small + (size_t)(c0 + r) * p.small_n + k);
        }
        nrw_commit();
    };

    float2 acc[4][SP / 2];
    float bsum[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        bsum[c] = 0.f;
#pragma unroll
        for (int k = 0; k < SP / 2; k++) acc[c][k] = make_float2(0.f, 0.f);
    }
    // head: this thread's 4 columns of W for every narrow column j, as (col, col + 1) pairs
    float2 w01[HEAD ? SP : 1], w23[HEAD ? SP : 1];
    if (HEAD) {
#pragma unroll
        for (int j = 0; j < SP; j++) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.W && j < p.small_n && col_ok) w = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)j * p.wide_n + col));
            w01[j] = make_float2(w.x, w.y);
            w23[j] = make_float2(w.z, w.w);
        }
    }

    if (n_chunks > 0) issue(0);
    for (int c = 0; c < n_chunks; c++) {
        nrw_wait<0>();
        __syncthreads();              // chunk c is visible to everybody; everybody is done reading chunk c - 1
        if (c + 1 < n_chunks) issue(c + 1);
        const int b = c & 1, c0 = r0 + c * kNrwCH, rows = min(kNrwCH, r1 - c0);
        if (col_ok) {
#pragma unroll 2
            for (int i = 0; i < kNrwCH / 4; i++) {
                const int r = 4 * i + q;
                if (r >= rows) break;
                const float4 wv = *reinterpret_cast<const float4*>(wt + (b * kNrwCH + r) * kNrwCols + 4 * cx);
                const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
                const float* srow = st + (b * kNrwCH + r) * SP;
                float2 t01 = make_float2(0.f, 0.f), t23 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k4 = 0; k4 < SP / 4; k4++) {
                    const float4 sv = *reinterpret_cast<const float4*>(srow + 4 * k4);
                    const float2 s01 = make_float2(sv.x, sv.y), s23 = make_float2(sv.z, sv.w);
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) {
                        acc[cc][2 * k4] = __ffma2_rn(make_float2(w4[cc], w4[cc]), s01, acc[cc][2 * k4]);
                        acc[cc][2 * k4 + 1] = __ffma2_rn(make_float2(w4[cc], w4[cc]), s23, acc[cc][2 * k4 + 1]);
                    }
                    if (HEAD) {
                        const float s4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            t01 = __ffma2_rn(make_float2(s4[u], s4[u]), w01[HEAD ? 4 * k4 + u : 0], t01);
                            t23 = __ffma2_rn(make_float2(s4[u], s4[u]), w23[HEAD ? 4 * k4 + u : 0], t23);
                        }
                    }
                }
                if (HEAD && p.gx) {
                    float4 o;
                    o.x = act_grad(w4[0], t01.x, p.act_prev); o.y = act_grad(w4[1], t01.y, p.act_prev);
                    o.z = act_grad(w4[2], t23.x, p.act_prev); o.w = act_grad(w4[3], t23.y, p.act_prev);
                    const size_t off = (size_t)(c0 + r) * p.wide_n + col;
                    *reinterpret_cast<float4*>(p.gx + off) = o;
                    bsum[0] += o.x; bsum[1] += o.y; bsum[2] += o.z; bsum[3] += o.w;      // column sums of dX = db of the layer below
                    if (p.gx_lo) {
                        float4 lo;
                        lo.x = o.x - __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u);
                        lo.y = o.y - __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                        lo.z = o.z - __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u);
                        lo.w = o.w - __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4*>(p.gx_lo + off) = lo;
                    }
                } else if (!HEAD) {
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) bsum[cc] += w4[cc];
                }
            }
        }
    }
    // row lanes 1..3 hand their partials to lane 0 one after the other (fixed order, one staging buffer)
    constexpr int RS = 4 * SP + 4;
    for (int src = 1; src < 4; src++) {
        __syncthreads();
        if (q == src) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
#pragma unroll
                for (int k = 0; k < SP / 2; k++) { red[cx * RS + c * SP + 2 * k] = acc[c][k].x; red[cx * RS + c * SP + 2 * k + 1] = acc[c][k].y; }
                red[cx * RS + 4 * SP + c] = bsum[c];
            }
        }
        __syncthreads();
        if (q == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
#pragma unroll
                for (int k = 0; k < SP / 2; k++) { acc[c][k].x += red[cx * RS + c * SP + 2 * k]; acc[c][k].y += red[cx * RS + c * SP + 2 * k + 1]; }
                bsum[c] += red[cx * RS + 4 * SP + c];
            }
        }
    }
    if (q == 0 && col_ok) {
        // partial layout: [slab y][z][wide_n][SP + 1]  (last entry of a row = column sum of the wide array)
        float* out = p.part + ((size_t)(blockIdx.y * R + blockIdx.z) * p.wide_n) * (SP + 1);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float* o = out + (size_t)(col + c) * (SP + 1);
#pragma unroll
            for (int k = 0; k < SP / 2; k++) { o[2 * k] = acc[c][k].x; o[2 * k + 1] = acc[c][k].y; }
            o[SP] = bsum[c];
        }
    }
}

template <int SP, bool HEAD>
static void launch_narrow_bwd(const NarrowArgs& a, dim3 grid) {
    static_assert(64 * (4 * SP + 4) <= 2 * kNrwCH * kNrwCols, "the hand-over buffer aliases the wide tiles");
    const size_t smem = (size_t)(2 * kNrwCH * kNrwCols + 2 * kNrwCH * SP) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(narrow_bwd_kernel<SP, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    B200_LAUNCH((narrow_bwd_kernel<SP, HEAD>), grid, 256, smem, a);
}

static bool narrow_enabled() {      // PPO_B200_NARROW=0 keeps the round-1 kernels (A/B runs)
    static int on = -1;
    if (on < 0) { const char* e = getenv("PPO_B200_NARROW"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

static int narrow_row_splits(int m, int gx, int splits) {
    if (m <= 16384) return 1;
    return std::max(1, std::min(16, (2 * num_sms()) / std::max(1, gx * splits)));     // one wave of 2 CTAs per SM
}

void skinny_fold(float* gW_part, float* gb_part, size_t stride, const float* part, int splits, int R, int wide_n, int small_n, int sp1, int n,
                 int wide_is_l);

// First layer of a low-dimensional env (n <= 32 inputs, l >= 64 units): dW / db slabs from ONE pass over g.
bool narrow_first_layer_backward(float* gW_part, float* gb_part, size_t stride, int splits, const float* g, const float* x, int m, int n, int l) {
    if (!narrow_enabled() || m < 1024 || n > 32 || l < 64 || (l & 3) || ((uintptr_t)g & 15)) return false;
    int rows = div_up(m, splits);
    rows = div_up(rows, 32) * 32;
    const int sp = n <= 8 ? 8 : n <= 16 ? 16 : n <= 20 ? 20 : n <= 24 ? 24 : 32;
    const int gx = div_up(l, kNrwCols);
    const int R = narrow_row_splits(m, gx, splits);
    NarrowArgs a{};
    a.wide = g; a.small = x; a.m = m; a.wide_n = l; a.small_n = n; a.rows_per_split = rows;
    a.part = static_cast<float*>(scratch(kScratchSkinny, (size_t)splits * R * l * (sp + 1) * sizeof(float)));
    const dim3 grid(gx, splits, R);
    if (sp == 8) launch_narrow_bwd<8, false>(a, grid);
    else if (sp == 16) launch_narrow_bwd<16, false>(a, grid);
    else if (sp == 20) launch_narrow_bwd<20, false>(a, grid);
    else if (sp == 24) launch_narrow_bwd<24, false>(a, grid);
    else launch_narrow_bwd<32, false>(a, grid);
    skinny_fold(gW_part, gb_part, stride, a.part, splits, R, l, n, sp + 1, n, 1);
    return true;
}

// Head (l <= 8 outputs, n >= 64 inputs): dW slabs and gx = (g W) act'(h) (+ optional lo companion) from ONE pass over h.
// db (column sums of the m x l array g) stays with the caller; gb_below != null receives the column sums of gx (the db slabs of
// the layer below, same slab stride).
bool narrow_head_backward(float* gW_part, size_t stride, int splits, float* gx, float* gx_lo, float* gb_below, const float* g, const float* h,
                          const float* W, int m, int n, int l, int act_prev) {
    if (!narrow_enabled() || m < 1024 || l > 8 || n < 64 || (n & 3) || ((uintptr_t)h & 15) || (gx && (((uintptr_t)W & 15) || ((uintptr_t)gx & 15))) ||
        (gx_lo && ((uintptr_t)gx_lo & 15)))
        return false;
    int rows = div_up(m, splits);
    rows = div_up(rows, 32) * 32;
    const int gxb = div_up(n, kNrwCols);
    const int R = narrow_row_splits(m, gxb, splits);
    NarrowArgs a{};
    a.wide = h; a.small = g; a.m = m; a.wide_n = n; a.small_n = l; a.rows_per_split = rows;
    a.part = static_cast<float*>(scratch(kScratchSkinny, (size_t)splits * R * n * 9 * sizeof(float)));
    a.W = W; a.gx = gx; a.gx_lo = gx_lo; a.act_prev = act_prev;
    launch_narrow_bwd<8, true>(a, dim3(gxb, splits, R));
    skinny_fold(gW_part, gb_below, stride, a.part, splits, R, n, l, 9, n, (gx && gb_below) ? 2 : 0);
    return true;
}

}  // namespace b200

namespace b200 {

// ---- forward of an l <= 8 wide head: y[r][j] = act(sum_k h[r][k] W[j][k] + b[j]) -------------------------------------
// One pass over h: a warp takes four rows per iteration (lanes stride k in float4 steps, up to 8 independent 128-bit loads
// in flight per lane), W sits in shared memory (one LDS.128 per (j, k4) shared by the four rows), butterfly sums at the end.
constexpr int kNhfRows = 4;
__global__ void __launch_bounds__(256, 3) narrow_head_forward_kernel(float* __restrict__ y, const float* __restrict__ h, const float* __restrict__ W,
                                                                  const float* __restrict__ b, int m, int n, int l, int act) {
    extern __shared__ __align__(16) float hsm[];          // [l][n]
    for (int e = threadIdx.x; e < l * n; e += 256) hsm[e] = __ldg(W + e);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    const int n4 = n >> 2;
    for (int r0 = warp * kNhfRows; r0 < m; r0 += n_warps * kNhfRows) {
        float acc[kNhfRows][8];
#pragma unroll
        for (int u = 0; u < kNhfRows; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[u][j] = 0.f;
        for (int k4 = lane; k4 < n4; k4 += 64) {
            float4 xv[kNhfRows][2];
#pragma unroll
            for (int u = 0; u < kNhfRows; u++)
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int kk = k4 + 32 * t;
                    xv[u][t] = (r0 + u < m && kk < n4) ? ld_stream4(h + (size_t)(r0 + u) * n + 4 * kk) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int kk = k4 + 32 * t;
                if (kk >= n4) break;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (j >= l) break;
                    const float4 w = *reinterpret_cast<const float4*>(hsm + j * n + 4 * kk);
#pragma unroll
                    for (int u = 0; u < kNhfRows; u++)
                        acc[u][j] = fmaf(xv[u][t].w, w.w, fmaf(xv[u][t].z, w.z, fmaf(xv[u][t].y, w.y, fmaf(xv[u][t].x, w.x, acc[u][j]))));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kNhfRows; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j >= l) break;
                const float s = warp_sum(acc[u][j]);
                if (lane == 0 && r0 + u < m) y[(size_t)(r0 + u) * l + j] = act_apply(s + __ldg(b + j), act);
            }
    }
}

bool narrow_head_forward(float* y, const float* h, const float* W, const float* b, int m, int n, int l, int act) {
    if (!narrow_enabled() || m < 1024 || l > 8 || n < 64 || (n & 3) || n > 4096 || ((uintptr_t)h & 15)) return false;
    const size_t smem = (size_t)l * n * sizeof(float);
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        CUDA_CHECK(cudaFuncSetAttribute(narrow_head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int blocks = std::min(div_up(m, 8 * kNhfRows), num_sms() * 3);        // one wave of the 3 CTAs an SM holds
    B200_LAUNCH(narrow_head_forward_kernel, blocks, 256, smem, y, h, W, b, m, n, l, act);
    return true;
}

// ---- forward of the first layer of a low-dimensional env: y[r][j] = act(sum_{k < n <= 32} x[r][k] W[j][k] + b[j]) ------
// Output-bandwidth bound (m x l floats, plus the 3xTF32 lo companion when asked for).  CTA tile = 64 rows x 256 columns,
// thread = 8 rows x 8 columns of packed FFMA2; the n x 256 weight slice (k-major) is staged once per CTA, the 64 x n input
// block (one contiguous range of x) goes through registers into a k-major tile while the previous block is computed.
constexpr int kNffRows = 64;
__global__ void __launch_bounds__(256, 2)
narrow_first_forward_kernel(float* __restrict__ y, float* __restrict__ y_lo, const float* __restrict__ x, const float* __restrict__ W,
                            const float* __restrict__ b, int m, int n, int l, int act) {
    constexpr int WLD = kNrwCols + 8, XLD = kNffRows + 4;
    extern __shared__ __align__(16) float fsm[];
    float (*Ws)[WLD] = reinterpret_cast<float (*)[WLD]>(fsm);                       // [n][WLD]
    float (*Xs0)[XLD] = reinterpret_cast<float (*)[XLD]>(fsm + n * WLD);            // [2][n][XLD]
    const int tid = threadIdx.x, tc = tid & 31, tr = tid >> 5;
    const int n0 = blockIdx.x * kNrwCols;
    for (int e = tid; e < n * kNrwCols; e += 256) {
        const int j = e / n, k = e - j * n;
        Ws[k][j] = (n0 + j < l) ? __ldg(W + (size_t)(n0 + j) * n + k) : 0.f;
    }
    // thread columns: {4tc .. 4tc+3} and {128 + 4tc ..}: the 32 lanes of a store instruction then write 512 contiguous bytes (whole
    // sectors); two adjacent float4 per lane would leave every sector half-written per instruction
    const int gj0 = n0 + 4 * tc, gj1 = gj0 + 128;
    const bool ok0 = gj0 < l, ok1 = gj1 < l;        // l is a multiple of 8 (host check)
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (ok0) b0 = __ldg(reinterpret_cast<const float4*>(b + gj0));
    if (ok1) b1 = __ldg(reinterpret_cast<const float4*>(b + gj1));
    const int n_tiles = (m + kNffRows - 1) / kNffRows;
    constexpr int kPer = (kNffRows * 32 + 255) / 256;          // input floats per thread per tile (n <= 32)
    float xr[kPer];
    auto fetch = [&](int tile) {
        const size_t base = (size_t)tile * kNffRows * n;
        const int cnt = min(kNffRows, m - tile * kNffRows) * n;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int e = tid + 256 * i;
            xr[i] = e < cnt ? __ldg(x + base + e) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const int e = tid + 256 * i;
            if (e < kNffRows * n) Xs0[buf * n + e % n][e / n] = xr[i];
        }
    };
    int tile = blockIdx.y;
    if (tile < n_tiles) { fetch(tile); stash(0); }
    __syncthreads();
    for (int it = 0; tile < n_tiles; tile += gridDim.y, it++) {
        const int buf = it & 1;
        const int next = tile + gridDim.y;
        if (next < n_tiles) fetch(next);
        float2 acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f);
        for (int k = 0; k < n; k++) {
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[k][4 * tc]), w1 = *reinterpret_cast<const float4*>(&Ws[k][128 + 4 * tc]);
            const float4 x0 = *reinterpret_cast<const float4*>(&Xs0[buf * n + k][8 * tr]), x1 = *reinterpret_cast<const float4*>(&Xs0[buf * n + k][8 * tr + 4]);
            const float2 wp[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
            const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = __ffma2_rn(make_float2(xs[r], xs[r]), wp[c], acc[r][c]);
        }
        if (ok0) {
            auto lo = [](float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); };
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int gi = tile * kNffRows + 8 * tr + r;
                if (gi >= m) break;
                float4 o0, o1;
                o0.x = act_apply(acc[r][0].x + b0.x, act); o0.y = act_apply(acc[r][0].y + b0.y, act);
                o0.z = act_apply(acc[r][1].x + b0.z, act); o0.w = act_apply(acc[r][1].y + b0.w, act);
                o1.x = act_apply(acc[r][2].x + b1.x, act); o1.y = act_apply(acc[r][2].y + b1.y, act);
                o1.z = act_apply(acc[r][3].x + b1.z, act); o1.w = act_apply(acc[r][3].y + b1.w, act);
                float* row = y + (size_t)gi * l;
                *reinterpret_cast<float4*>(row + gj0) = o0;
                if (ok1) *reinterpret_cast<float4*>(row + gj1) = o1;
                if (y_lo) {
                    float* rl = y_lo + (size_t)gi * l;
                    *reinterpret_cast<float4*>(rl + gj0) = make_float4(lo(o0.x), lo(o0.y), lo(o0.z), lo(o0.w));
                    if (ok1) *reinterpret_cast<float4*>(rl + gj1) = make_float4(lo(o1.x), lo(o1.y), lo(o1.z), lo(o1.w));
                }
            }
        }
        if (next < n_tiles) stash(buf ^ 1);        // the other buffer was last read one iteration ago (barrier below)
        __syncthreads();
    }
}

bool narrow_first_forward(float* y, float* y_lo, const float* x, const float* W, const float* b, int m, int n, int l, int act) {
    if (!narrow_enabled() || m < 1024 || n > 32 || l < 64 || (l & 7) || ((uintptr_t)y & 15) || ((uintptr_t)b & 15) || (y_lo && ((uintptr_t)y_lo & 15)))
        return false;
    const int gx = div_up(l, kNrwCols);
    const int tiles = div_up(m, kNffRows);
    const dim3 grid(gx, std::max(1, std::min(tiles, (2 * num_sms()) / gx)), 1);
    const size_t smem = (size_t)n * ((kNrwCols + 8) + 2 * (kNffRows + 4)) * sizeof(float);      // 51 KB at n = 32
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(narrow_first_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    B200_LAUNCH(narrow_first_forward_kernel, grid, 256, smem, y, y_lo, x, W, b, m, n, l, act);
    return true;
}

}  // namespace b200
