// fused_mlp.cu — the small-net fast path (SURVEY.md §8 rows a7+a9..a14 in two launches per minibatch).
//
// For the reference-width actor/critic nets (every layer width <= 128: 2x64, the reference's default 2x128) one
// minibatch update is
//   fused_tile64_kernel   gather rows by permutation index -> forward through ALL layers -> fused loss
//                         head (MSE, or Gaussian log-prob + PPO-clip surrogate) -> backward through
//                         all layers -> this CTA's partial gradient slab.  Weights are staged ONCE per
//                         CTA in shared memory, activations never leave shared memory, nothing but
//                         the slab is written to HBM.
//   fused_reduce_adam_kernel  fixed-order sum of the slabs (+ the cross-GPU sum over NVLink peer memory under data
//                         parallelism) + Adam on the flat parameter vector (+ the log_std vector for the policy) +
//                         loss accumulation; also refreshes the pre-transposed weight image the next
//                         fused_tile64_kernel will stage.
// versus ~14 launches through the layer-wise kernels of gemm.cu/policy.cu/adam.cu (which remain the
// generic path for wider nets).  The reference does this with ~25 launches, 1-3 blocking D2H reads
// and 1-2 cudaMallocs per minibatch (src/ppo.cu:495-532).  The two kernels are chained with programmatic
// dependent launch: the gather prologue of minibatch k+1 overlaps the Adam kernel of minibatch k.
//
// Shared-memory layouts (64 rows per CTA, row stride TMP = 68 floats so that TMP % 32 == 4):
//   activations / gradients  At[feature][TMP]   feature-major: a thread reads 4 consecutive ROWS with
//                            one conflict-free LDS.128
//   weights                  Wt[in][ldw]        k-major transpose of the reference's W[out][in]; the whole
//                            "image" [Wt_0|Wt_1|..|biases] is kept pre-transposed in global memory by the
//                            Adam kernel and lands in shared memory with ONE TMA bulk copy
//                            (cp.async.bulk + mbarrier complete_tx)
// All arithmetic is fp32 (packed FFMA2 = two fma.rn per instruction; tolerance 1e-5).
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr double kPiF = 3.14159265358979323846;

enum FusedMode { kFusedForward = 0, kFusedValue = 1, kFusedPolicy = 2 };

struct FusedArgs {
    FusedNet net;
    const float* image;        // pre-transposed weight image in global memory
    const int* idx;            // permutation (may be null: row = offset + r)
    int offset, limit, m, m_total, mode;
    const float* state;        // [*][S]
    const float* action;       // [*][A]
    const float* logprob;      // [*]
    const float* advantage;    // [*]
    const float* adv_target;   // [*]
    float* y_out;              // forward mode: [m][out]
    const float* log_std;
    float epsilon, ent_coeff;
    float* partials;           // [gridDim.x][slab]  slab = P + A + 1 (grads | grad_log_std | loss term)
    int slab;
    int smem_g_off;            // offset of the two gradient buffers
    int smem_red_off;          // 64 floats of reduction scratch, TM ints of source rows, 1 mbarrier
    unsigned long long* dbg;   // phase timestamps (PPO_B200_PHASE_DEBUG), [block][16]
};


// ---- TMA bulk copy + mbarrier (SASS: UBLKCP + SYNCS) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ float fused_log_prob(const float* mu, const float* log_std, const float* action, int A) {
    float logprob = (float)(-0.5 * A * (double)logf((float)(2 * kPiF)));   // src/policy.cu:67-74
    for (int j = 0; j < A; j++) {
        const float z = __fdiv_rn(__fsub_rn(action[j], mu[j]), expf(log_std[j]));
        logprob = (float)((double)logprob - ((double)log_std[j] + 0.5 * (double)__fmul_rn(z, z)));
    }
    return logprob;
}

// ===================================================================================================
// Tile kernel (every layer width <= 128: the reference's Pendulum nets, 2x64 and the default 2x128).
//
// 256 threads = two groups of four warps working on one 64-row tile, two CTAs per SM (16 warps per SM) for 64-wide
// nets, one for 128-wide nets (layers wider than 64 are processed in 64-column / 64x64 blocks).
//   * 8 rows x 4 columns per thread in forward and dX: 3 LDS.128 feed 32 FMAs (a 4x4 tile needs 2 per 16;
//     shared memory delivers 128 B/clk/SM, so floats-loaded-per-FMA is what bounds an fp32 tile kernel);
//   * the two groups split the WORK, not the tile: forward = split-K halves that meet through one
//     exchange buffer; backward = group 0 computes dW_l/db_l while group 1 computes dX_l (independent given
//     G_{l+1} and A_l), so the backward critical path is max(dW, dX) instead of their sum;
//   * dX walks the reduction 4 j's at a time on the SAME k-major weight image (Wt rows padded to
//     ldw = pad4(out)+4 floats so the four k-rows a warp touches sit in different banks);
//   * skinny layers (out <= 8: value head, action mean) and the bias gradients have their own thread
//     mappings instead of running a mostly-empty 64x64 tile.
// Rows of a thread: {4tr..4tr+3} U {32+4tr..32+4tr+3}, tr = lt & 7; column group tc = lt >> 3 (lt = tid & 127).
// ===================================================================================================
constexpr int kT64Threads = 256;
constexpr int kT64TM = 64;
constexpr int kT64TMP = 68;      // feature row stride in floats; 68 % 32 == 4 keeps feature-strided LDS.128 conflict-free

__device__ __forceinline__ void t64_stamp(const FusedArgs& p, int slot) {
    if (p.dbg && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.dbg[(size_t)blockIdx.x * 16 + slot] = t;
    }
}

// Packed fp32 FMA (FFMA2): two lanes of a register pair per instruction.  Measured on B200 (profiles/NOTES.md): an 8x4
// register tile of scalar 3-register FFMAs sustains 48 TFLOP/s, the same tile with FFMA2 65 TFLOP/s.  Each element is
// still one fma.rn, so results are bit-identical to the scalar form.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 bcast2(float x) { return make_float2(x, x); }

__device__ __forceinline__ void t64_load8(const float* base, float (&a)[8]) {
    const float4 a0 = *reinterpret_cast<const float4*>(base);
    const float4 a1 = *reinterpret_cast<const float4*>(base + 32);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
}
__device__ __forceinline__ void t64_store8(float* base, const float (&a)[8]) {
    *reinterpret_cast<float4*>(base) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(base + 32) = make_float4(a[4], a[5], a[6], a[7]);
}

// Yt[j][r] = act(sum_k Xt[k][r] * Wt[k][j] + b[j]),  n_out in (8, 64].
// n_in >= 16: split-K.  Both groups accumulate the full 8x4 tile of (tr, tc) over their half of k; then group 0
//   finalises the tile's first row quad and group 1 the second: each thread passes the quad it does not
//   finalise through `xch` ([64][TMP]) and adds (lower-k partial) + (upper-k partial) for its own quad, so
//   the activation epilogue is spread over all 256 threads.
// n_in < 16: no split; group g simply computes its own row quad (4x4 tile) over all of k.
// Contains one __syncthreads.
__device__ __forceinline__ void t64_forward(const float* __restrict__ Xt, const float* __restrict__ Wt, int ldw,
                                            const float* __restrict__ bias, float* __restrict__ Yt, float* __restrict__ xch,
                                            int n_in, int n_out, int act, int lt, int g, int cb) {
    // cb = first column of the 64-column block this call computes (layers wider than 64 are done block by block)
    const int tr = lt & 7, tc = lt >> 3;
    const bool live = cb + 4 * tc < pad4(n_out);
    const bool split = n_in >= 16;
    float mine[4][4];                       // [row of my quad][col]
    if (split) {
        const int kh = (n_in + 1) >> 1;
        const int k0 = g ? kh : 0, k1 = g ? n_in : kh;
        float2 acc[4][4];                   // [row pair][col]: rows (4tr, 4tr+1), (4tr+2, 4tr+3), (32+4tr, ..), (32+4tr+2, ..)
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f);
        if (live) {
            const float* xp = Xt + 4 * tr;
            const float* wp = Wt + cb + 4 * tc;
#pragma unroll 2
            for (int k = k0; k < k1; k++) {
                const float4 a0 = *reinterpret_cast<const float4*>(xp + k * kT64TMP);
                const float4 a1 = *reinterpret_cast<const float4*>(xp + k * kT64TMP + 32);
                const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
                const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
                const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[r][c] = ffma2(ap[r], bcast2(wv[c]), acc[r][c]);
            }
            // pass the other group's quad: group 0 sends rows 32+4tr.. (pairs 2,3), group 1 sends rows 4tr.. (pairs 0,1)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float* dst = xch + (4 * tc + c) * kT64TMP + 4 * tr + (g ? 0 : 32);
                *reinterpret_cast<float4*>(dst) = g ? make_float4(acc[0][c].x, acc[0][c].y, acc[1][c].x, acc[1][c].y)
                                                    : make_float4(acc[2][c].x, acc[2][c].y, acc[3][c].x, acc[3][c].y);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float2 lo = g ? acc[2][c] : acc[0][c], hi = g ? acc[3][c] : acc[1][c];
            mine[0][c] = lo.x; mine[1][c] = lo.y; mine[2][c] = hi.x; mine[3][c] = hi.y;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) mine[r][c] = 0.f;
        if (live) {
            const float* xp = Xt + 4 * tr + 32 * g;
            const float* wp = Wt + cb + 4 * tc;
            for (int k = 0; k < n_in; k++) {
                const float4 a = *reinterpret_cast<const float4*>(xp + k * kT64TMP);
                const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
                const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) mine[r][c] = fmaf(av[r], wv[c], mine[r][c]);
            }
        }
    }
    __syncthreads();
    if (live) {
        const float4 b = *reinterpret_cast<const float4*>(bias + cb + 4 * tc);     // bias block is zero-padded to pad4
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float o[4] = {mine[0][c], mine[1][c], mine[2][c], mine[3][c]};
            if (split) {
                const float4 q = *reinterpret_cast<const float4*>(xch + (4 * tc + c) * kT64TMP + 4 * tr + 32 * g);
                const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int r = 0; r < 4; r++) o[r] = g ? qv[r] + o[r] : o[r] + qv[r];     // lower-k partial + upper-k partial
            }
#pragma unroll
            for (int r = 0; r < 4; r++) o[r] = act_apply(o[r] + bv[c], act);
            *reinterpret_cast<float4*>(Yt + (cb + 4 * tc + c) * kT64TMP + 4 * tr + 32 * g) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Skinny forward, n_out <= 8: thread = (row, k-quarter); the quarters meet through `scratch` (>= 3*8*64 floats).
// Contains one __syncthreads: every thread of the CTA must call it.
__device__ __forceinline__ void t64_forward_skinny(const float* __restrict__ Xt, const float* __restrict__ Wt, int ldw,
                                                   const float* __restrict__ bias, float* __restrict__ Yt, float* __restrict__ scratch,
                                                   int n_in, int n_out, int act) {
    const int r = threadIdx.x & 63, h = threadIdx.x >> 6;      // h = 0..3
    const int kq = (n_in + 3) >> 2;
    const int k0 = min(n_in, h * kq), k1 = min(n_in, k0 + kq);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int k = k0; k < k1; k++) {
        const float x = Xt[k * kT64TMP + r];
        const float4 w0 = *reinterpret_cast<const float4*>(Wt + k * ldw);
        acc[0] = fmaf(x, w0.x, acc[0]); acc[1] = fmaf(x, w0.y, acc[1]); acc[2] = fmaf(x, w0.z, acc[2]); acc[3] = fmaf(x, w0.w, acc[3]);
        if (n_out > 4) {
            const float4 w1 = *reinterpret_cast<const float4*>(Wt + k * ldw + 4);
            acc[4] = fmaf(x, w1.x, acc[4]); acc[5] = fmaf(x, w1.y, acc[5]); acc[6] = fmaf(x, w1.z, acc[6]); acc[7] = fmaf(x, w1.w, acc[7]);
        }
    }
    if (h) {
#pragma unroll
        for (int j = 0; j < 8; j++) if (j < n_out) scratch[((h - 1) * 8 + j) * 64 + r] = acc[j];
    }
    __syncthreads();
    if (!h) {
        const int out_pad = pad4(n_out);
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < out_pad) {
                const float s = ((acc[j] + scratch[j * 64 + r]) + scratch[(8 + j) * 64 + r]) + scratch[(16 + j) * 64 + r];
                Yt[j * kT64TMP + r] = (j < n_out) ? act_apply(s + bias[j], act) : 0.f;
            }
    }
}

// Gout[k][r] = (sum_j Gt[j][r] * Wt[k][j]) * act'(Ht[k][r]);  k = tc + 16c (interleaved so the four Wt rows a
// warp reads per load are consecutive -> different banks).  Gt rows [n_out, pad4(n_out)) are zero.
// Rows [n_in, pad4(n_in)) of Gout are zero-filled (they are read as padding by the next dX).
__device__ __forceinline__ void t64_backward_input(const float* __restrict__ Gt, const float* __restrict__ Wt, int ldw,
                                                   const float* __restrict__ Ht, float* __restrict__ Gout, int n_in, int n_out,
                                                   int act_prev, int lt, int kb) {
    // kb = first of the 64 input features (output rows of Gout) this call produces
    const int tr = lt & 7, tc = lt >> 3;
    if (kb + tc >= pad4(n_in)) return;
    float2 acc[4][4];                       // [row pair][k column]
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = make_float2(0.f, 0.f);
    const float* wrow[4];
#pragma unroll
    for (int c = 0; c < 4; c++) wrow[c] = Wt + (size_t)min(kb + tc + 16 * c, n_in - 1) * ldw;
    const float* gp = Gt + 4 * tr;
    const int jpad = pad4(n_out);
#pragma unroll 1
    for (int j0 = 0; j0 < jpad; j0 += 4) {
        float wv[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float4 w = *reinterpret_cast<const float4*>(wrow[c] + j0);
            wv[c][0] = w.x; wv[c][1] = w.y; wv[c][2] = w.z; wv[c][3] = w.w;
        }
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const float4 g0 = *reinterpret_cast<const float4*>(gp + (j0 + jj) * kT64TMP);
            const float4 g1 = *reinterpret_cast<const float4*>(gp + (j0 + jj) * kT64TMP + 32);
            const float2 gq[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = ffma2(gq[r], bcast2(wv[c][jj]), acc[r][c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int k = kb + tc + 16 * c;
        if (k >= pad4(n_in)) continue;
        float h[8], o[8];
        t64_load8(Ht + k * kT64TMP + 4 * tr, h);
        const float av[8] = {acc[0][c].x, acc[0][c].y, acc[1][c].x, acc[1][c].y, acc[2][c].x, acc[2][c].y, acc[3][c].x, acc[3][c].y};
#pragma unroll
        for (int r = 0; r < 8; r++) o[r] = (k < n_in) ? act_grad(h[r], av[r], act_prev) : 0.f;
        t64_store8(Gout + k * kT64TMP + 4 * tr, o);
    }
}

// gW[j][k] = sum_r Gt[j][r] * Xt[k][r];  j = tj + 16a (tj = lt >> 3), k = tk + 8b (tk = lt & 7): the 4 G rows
// and 8 X rows a warp loads per instruction are consecutive features.
template <int JJ, int KK>
__device__ __forceinline__ void t64_backward_weights(const float* __restrict__ Gt, const float* __restrict__ Xt,
                                                     float* __restrict__ gW, int n_in, int n_out, int lt, int jb, int kb) {
    // (jb, kb) = first output row / input column of the 64x64 block of gW this call produces
    const int tk = lt & 7, tj = lt >> 3;
    if (jb + (tj & ~3) >= n_out) return;     // warp-uniform: this warp owns no valid output row
    float2 acc[JJ][KK];                     // (sum over even rows, sum over odd rows): both FFMA2 operands are natural pairs
#pragma unroll
    for (int a = 0; a < JJ; a++)
#pragma unroll
        for (int b = 0; b < KK; b++) acc[a][b] = make_float2(0.f, 0.f);
    int jrow[JJ], krow[KK];
#pragma unroll
    for (int a = 0; a < JJ; a++) jrow[a] = min(jb + tj + 16 * a, n_out - 1) * kT64TMP;
#pragma unroll
    for (int b = 0; b < KK; b++) krow[b] = min(kb + tk + 8 * b, n_in - 1) * kT64TMP;
#pragma unroll 1
    for (int r = 0; r < kT64TM; r += 4) {
        float4 g[JJ], x[KK];
#pragma unroll
        for (int a = 0; a < JJ; a++) g[a] = *reinterpret_cast<const float4*>(Gt + jrow[a] + r);
#pragma unroll
        for (int b = 0; b < KK; b++) x[b] = *reinterpret_cast<const float4*>(Xt + krow[b] + r);
#pragma unroll
        for (int a = 0; a < JJ; a++)
#pragma unroll
            for (int b = 0; b < KK; b++) {
                acc[a][b] = ffma2(make_float2(g[a].x, g[a].y), make_float2(x[b].x, x[b].y), acc[a][b]);
                acc[a][b] = ffma2(make_float2(g[a].z, g[a].w), make_float2(x[b].z, x[b].w), acc[a][b]);
            }
    }
#pragma unroll
    for (int a = 0; a < JJ; a++) {
        const int j = jb + tj + 16 * a;
        if (j >= n_out) continue;
#pragma unroll
        for (int b = 0; b < KK; b++) {
            const int k = kb + tk + 8 * b;
            if (k < n_in) gW[(size_t)j * n_in + k] = acc[a][b].x + acc[a][b].y;
        }
    }
}

// Skinny dW.  WIDE_IS_K = true : n_out <= 8 (value head / action mean): thread = (k, row half), acc[j]
//            WIDE_IS_K = false: n_in  <= 8 (first layer of a low-dimensional env): thread = (j, row half), acc[k]
// The two row halves meet through one shuffle; sums run in a fixed order.
template <bool WIDE_IS_K>
__device__ __forceinline__ void t64_backward_weights_skinny(const float* __restrict__ Gt, const float* __restrict__ Xt,
                                                            float* __restrict__ gW, int n_in, int n_out, int lt, int wb) {
    const int wide = wb + (lt >> 1), h = lt & 1;       // wb = first of the 64 wide-side features of this call
    const int n_wide = WIDE_IS_K ? n_in : n_out, n_small = WIDE_IS_K ? n_out : n_in;
    const float* wide_base = (WIDE_IS_K ? Xt : Gt) + min(wide, n_wide - 1) * kT64TMP + 32 * h;
    const float* small_base = (WIDE_IS_K ? Gt : Xt) + 32 * h;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; q++) acc[q] = 0.f;
#pragma unroll 2
    for (int i = 0; i < 8; i++) {
        const float4 w = *reinterpret_cast<const float4*>(wide_base + 4 * i);
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (q < n_small) {
                const float4 sv = *reinterpret_cast<const float4*>(small_base + q * kT64TMP + 4 * i);
                acc[q] = fmaf(w.x, sv.x, acc[q]); acc[q] = fmaf(w.y, sv.y, acc[q]);
                acc[q] = fmaf(w.z, sv.z, acc[q]); acc[q] = fmaf(w.w, sv.w, acc[q]);
            }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) acc[q] += __shfl_xor_sync(kFull, acc[q], 1);
    if (h == 0 && wide < n_wide) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (q < n_small) {
                if (WIDE_IS_K) gW[(size_t)q * n_in + wide] = acc[q];
                else gW[(size_t)wide * n_in + q] = acc[q];
            }
    }
}

__device__ __forceinline__ void t64_weights_dispatch(const float* Gt, const float* Xt, float* gW, int n_in, int n_out, int lt) {
    // few tile shapes only (instruction-cache footprint); layers wider than 64 go 64x64 block by block
    if (n_out <= 8) {
        for (int wb = 0; wb < n_in; wb += 64) t64_backward_weights_skinny<true>(Gt, Xt, gW, n_in, n_out, lt, wb);
    } else if (n_in <= 8) {
        for (int wb = 0; wb < n_out; wb += 64) t64_backward_weights_skinny<false>(Gt, Xt, gW, n_in, n_out, lt, wb);
    } else {
        for (int jb = 0; jb < n_out; jb += 64)
            for (int kb = 0; kb < n_in; kb += 64) {
                if (n_out - jb <= 16) t64_backward_weights<1, 8>(Gt, Xt, gW, n_in, n_out, lt, jb, kb);
                else t64_backward_weights<4, 8>(Gt, Xt, gW, n_in, n_out, lt, jb, kb);
            }
    }
}

// gb[j] = sum_r Gt[j][r]: thread = (j, row half), fixed-order sums
__device__ __forceinline__ void t64_bias_grad(const float* __restrict__ Gt, float* __restrict__ gb, int n_out, int lt) {
    for (int jb = 0; jb < n_out; jb += 64) {
        const int j = jb + (lt >> 1), h = lt & 1;
        float s = 0.f;
        if (j < n_out) {
            const float* gp = Gt + j * kT64TMP + 32 * h;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 g = *reinterpret_cast<const float4*>(gp + 4 * i);
                s += (g.x + g.y) + (g.z + g.w);
            }
        }
        s += __shfl_xor_sync(kFull, s, 1);
        if (h == 0 && j < n_out) gb[j] = s;
    }
}

__global__ void __launch_bounds__(kT64Threads, 2) fused_tile64_kernel(const FusedArgs p) {
    constexpr int TM = kT64TM, TMP = kT64TMP;
    extern __shared__ __align__(128) float smem[];
    const FusedNet& net = p.net;
    const int tid = threadIdx.x, lt = tid & 127, grp = tid >> 7;
    const int row0 = blockIdx.x * TM;
    const int S = net.sizes[0], OUT = net.sizes[net.L];
    float* img = smem;
    float* act0 = smem + net.img_floats;
    float* red = smem + p.smem_red_off;                  // 64 floats
    int* src_rows = reinterpret_cast<int*>(red + 64);    // TM ints
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 64 + TM);
    float* ebuf = smem + p.smem_g_off;                   // E0 | E1: forward exchange / out-of-place dX ping-pong
    float* scratch = ebuf;                               // skinny forward: 3*8*64 floats (aliases E0)

    t64_stamp(p, 0);
    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    float h_target = 0.f, h_adv = 0.f, h_lp_old = 0.f, h_act[8];
#pragma unroll
    for (int j = 0; j < 8; j++) h_act[j] = 0.f;
    int my_src = -1;
    if (tid < TM) {
        const int r = row0 + tid;
        if (r < p.m) my_src = p.idx ? p.idx[(p.offset + r) % p.limit] : p.offset + r;
        src_rows[tid] = my_src;
    }
    __syncthreads();
    {   // gather the input tile: Xt[k][r] = state[src][k]; padded feature rows are zero
        float* Xt = act0 + net.a_off[0];
        const int SP = pad4(S);
        for (int e = tid; e < TM * SP; e += kT64Threads) {
            const int r = e / SP, k = e - r * SP;
            const int src = src_rows[r];
            Xt[k * TMP + r] = (src >= 0 && k < S) ? p.state[(size_t)src * S + k] : 0.f;
        }
    }
    if (my_src >= 0) {
        if (p.mode == kFusedValue) {
            h_target = p.adv_target[my_src];
        } else if (p.mode == kFusedPolicy) {
            h_adv = p.advantage[my_src];
            h_lp_old = p.logprob[my_src];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < OUT) h_act[j] = p.action[(size_t)my_src * OUT + j];
        }
    }
    // Everything above reads only rollout data and the permutation.  The weight image is written by the
    // previous minibatch's Adam kernel: under programmatic dependent launch this is where we wait for it.
    t64_stamp(p, 1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    t64_stamp(p, 2);
    if (tid == 0) {
        mbar_expect_tx(mbar, (uint32_t)net.img_floats * 4u);
        tma_bulk_g2s(img, p.image, (uint32_t)net.img_floats * 4u, mbar);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    mbar_wait(mbar, 0);
    __syncthreads();
    t64_stamp(p, 3);
    // ---- forward
    for (int l = 0; l < net.L; l++) {
        const float* Xt = act0 + net.a_off[l];
        float* Yt = act0 + net.a_off[l + 1];
        if (net.sizes[l + 1] <= 8)
            t64_forward_skinny(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, scratch, net.sizes[l], net.sizes[l + 1], net.acts[l]);
        else
            for (int cb = 0; cb < pad4(net.sizes[l + 1]); cb += 64) {
                if (cb) __syncthreads();           // the previous block's finalisation has read the exchange buffer
                t64_forward(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, ebuf, net.sizes[l], net.sizes[l + 1], net.acts[l], lt, grp, cb);
            }
        __syncthreads();
        t64_stamp(p, 4 + l);
    }
    float* Yt = act0 + net.a_off[net.L];
    if (p.mode == kFusedForward) {
        for (int e = tid; e < TM * OUT; e += kT64Threads) {
            const int r = e / OUT, j = e - r * OUT;
            if (row0 + r < p.m) p.y_out[(size_t)(row0 + r) * OUT + j] = Yt[j * TMP + r];
        }
        return;
    }
    float* slab = p.partials + (size_t)blockIdx.x * p.slab;
    // ---- fused loss head: dLoss/dy written IN PLACE over the output tile (rows >= OUT of it stay zero)
    {
        float loss_term = 0.f;
        float gls[8];
#pragma unroll
        for (int j = 0; j < 8; j++) gls[j] = 0.f;
        if (tid < TM) {
            float gout[8];
#pragma unroll
            for (int j = 0; j < 8; j++) gout[j] = 0.f;
            float yv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) yv[j] = (j < OUT) ? Yt[j * TMP + tid] : 0.f;
            if (my_src >= 0) {
                if (p.mode == kFusedValue) {            // src/loss.cu:5-23
                    gout[0] = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(yv[0], h_target)), (float)p.m_total);
                    const float d = __fsub_rn(h_target, yv[0]);
                    loss_term = __fmul_rn(d, d);
                } else {                                // src/policy.cu:67-111 + src/ppo.cu:89-98
                    const float lp = fused_log_prob(yv, p.log_std, h_act, OUT);
                    const float ratio = expf(__fsub_rn(lp, h_lp_old));
                    const bool adv_pos = h_adv > 0.f;
                    const bool hi = ratio > 1.f + p.epsilon, lo = ratio < 1.f - p.epsilon;
                    const float sel = adv_pos ? (hi ? 1.f + p.epsilon : ratio) : (lo ? 1.f - p.epsilon : ratio);
                    loss_term = __fmul_rn(h_adv, sel);
                    const int keep = adv_pos ? !hi : !lo;
                    const float g = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), h_adv), ratio), (float)p.m_total);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j < OUT) {
                            const float e2 = expf(-2.f * p.log_std[j]);
                            const float diff = __fsub_rn(h_act[j], yv[j]);
                            gout[j] = __fmul_rn(__fmul_rn(diff, e2), g);
                            gls[j] = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), g);
                        }
                }
            }
            const int out_act = net.acts[net.L - 1];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < OUT) Yt[j * TMP + tid] = act_grad(yv[j], gout[j], out_act);
        }
        if (tid < 64) {                                    // warps 0..1 hold the data
            const int warp = tid >> 5, lane = tid & 31;
            const float v = warp_sum(loss_term);
            if (lane == 0) red[warp] = v;
            if (p.mode == kFusedPolicy) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < OUT) { const float s = warp_sum(gls[j]); if (lane == 0) red[8 + j * 2 + warp] = s; }
            }
        }
        __syncthreads();
        if (tid == 0) slab[net.P + OUT] = red[0] + red[1];
        if (p.mode == kFusedPolicy && tid < OUT) slab[net.P + tid] = red[8 + tid * 2] + red[8 + tid * 2 + 1];
    }
    t64_stamp(p, 9);
    // ---- backward: group 0 -> dW_l, db_l; group 1 -> dX_l into the ping-pong buffer (both read G_{l+1}, A_l)
    const float* G = Yt;
    for (int l = net.L - 1; l >= 0; l--) {
        const int n_in = net.sizes[l], n_out = net.sizes[l + 1];
        const float* Xt = act0 + net.a_off[l];
        float* Gout = ebuf + ((net.L - 1 - l) & 1) * (net.max_width_pad * TMP);
        if (grp == 0) {
            t64_weights_dispatch(G, Xt, slab + net.w_off[l], n_in, n_out, lt);
            t64_bias_grad(G, slab + net.b_off[l], n_out, lt);
        } else if (l > 0) {
            for (int kb = 0; kb < pad4(n_in); kb += 64)
                t64_backward_input(G, img + net.wt_off[l], net.ldw[l], Xt, Gout, n_in, n_out, net.acts[l - 1], lt, kb);
        }
        if (l > 0) __syncthreads();
        t64_stamp(p, 10 + l);
        G = Gout;
    }
}

// ---- slab reduction + Adam ----------------------------------------------------------------------
struct AdamSeg {             // one optimiser: parameters [begin, end) of the slab
    float *w, *g, *m, *v;
    float beta1, beta2, omb1, omb2, bc2, step_size;
};
struct ReduceAdamArgs {
    const float* partials;
    int nparts, slab, P, A;
    AdamSeg net, ls;         // flat net parameters; log_std (policy mode only: ls.w != null)
    float* loss_slot;        // accumulates: value: sum/m_total ; policy: -sum/m_total - ent_coeff*entropy
    int mode, m_total;
    float ent_coeff;
    const float* log_std;
    FusedNet layout;         // to refresh the transposed weight image
    float* image;
    int apply;               // 0: only write the reduced slab to `reduced` (data-parallel path through NCCL)
    float* reduced;
    PeerView peer;           // peer.ready: data-parallel exchange over NVLink peer memory inside this kernel
};

__device__ __forceinline__ float adam_apply(const AdamSeg& s, int i, float g, float m, float v, float w) {
    m = __fadd_rn(__fmul_rn(s.beta1, m), __fmul_rn(s.omb1, g));
    v = __fadd_rn(__fmul_rn(s.beta2, v), __fmul_rn(s.omb2, __fmul_rn(g, g)));
    const float denom = (float)((double)__fsqrt_rn(__fdiv_rn(v, s.bc2)) + 1e-8);
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(s.step_size, m), denom));
    s.m[i] = m; s.v[i] = v; s.w[i] = w; s.g[i] = g;
    return w;
}

// position of flat parameter e inside the transposed image
__host__ __device__ inline int image_index(const FusedNet& n, int e) {
    for (int l = 0; l < n.L; l++) {
        const int n_in = n.sizes[l], n_out = n.sizes[l + 1];
        if (e < n.b_off[l]) {
            const int q = e - n.w_off[l], j = q / n_in, k = q - j * n_in;
            return n.wt_off[l] + k * n.ldw[l] + j;
        }
        if (e < n.b_off[l] + n_out) return n.bs_off[l] + (e - n.b_off[l]);
    }
    return -1;
}

// One CTA per 32 consecutive slab entries; 16 warps split the slabs (each keeps up to 16 independent loads in
// flight, so 256 slabs cost one L2 round trip), fixed-order combine -> deterministic gradients.
constexpr int kRedWarps = 16;
__global__ void __launch_bounds__(32 * kRedWarps) fused_reduce_adam_kernel(const ReduceAdamArgs p) {
    __shared__ float red[kRedWarps][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    const int total = p.P + p.A + 1;
    // Optimiser state of this thread's element: written by the PREVIOUS Adam launch (complete long ago), so it can
    // be fetched before the dependency wait and its latency hides behind the tail of the update kernel.
    float pm = 0.f, pv = 0.f, pw = 0.f;
    const bool is_net = e < p.P, is_ls = !is_net && e < p.P + p.A && p.mode == kFusedPolicy;
    if (warp == 0 && p.apply) {
        if (is_net) { pm = p.net.m[e]; pv = p.net.v[e]; pw = p.net.w[e]; }
        else if (is_ls) { pm = p.ls.m[e - p.P]; pv = p.ls.v[e - p.P]; pw = p.ls.w[e - p.P]; }
        else if (e == p.P + p.A && p.mode == kFusedPolicy) {          // entropy of the policy that produced this loss
            pw = (float)(p.A * 0.5 * (1 + log(2 * kPiF)));            // src/policy.cu:171-178
            for (int j = 0; j < p.A; j++) pw += p.log_std[j];
        }
    }
    // programmatic dependent launch: the slabs are written by the update kernel that precedes us in the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float s = 0.f;
    if (e < total) {
        const int per = (p.nparts + kRedWarps - 1) / kRedWarps;
        const int b0 = warp * per, b1 = min(p.nparts, b0 + per);
        const float* src = p.partials + e;
        int b = b0;
        for (; b + 16 <= b1; b += 16) {     // 16 independent loads in flight, added in slab order
            float t[16];
#pragma unroll
            for (int u = 0; u < 16; u++) t[u] = src[(size_t)(b + u) * p.slab];
#pragma unroll
            for (int u = 0; u < 16; u++) s += t[u];
        }
        for (; b + 4 <= b1; b += 4) {
            float t[4];
#pragma unroll
            for (int u = 0; u < 4; u++) t[u] = src[(size_t)(b + u) * p.slab];
#pragma unroll
            for (int u = 0; u < 4; u++) s += t[u];
        }
        for (; b < b1; b++) s += src[(size_t)b * p.slab];
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp != 0) return;
    float g = 0.f;
    if (e < total) {
        g = red[0][lane];
#pragma unroll
        for (int w = 1; w < kRedWarps; w++) g += red[w][lane];
    }
    if (p.peer.ready && e < total) {
        // ---- gradient exchange over NVLink peer memory (dist.cu "peer arena"): push {tag, value}, poll, ordered sum
        const PeerView& pv = p.peer;
        const unsigned long long packed = ((unsigned long long)pv.epoch << 32) | (unsigned long long)__float_as_uint(g);
        for (int r = 0; r < pv.world; r++)
            if (r != pv.rank) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(pv.peer_recv[r] + e), "l"(packed) : "memory");
        const long long t0 = clock64();
        float sum = 0.f;
        for (int r = 0; r < pv.world; r++) {
            float x = g;
            if (r != pv.rank) {
                const unsigned long long* src = pv.my_recv + (size_t)r * kPeerCap + e;
                unsigned long long w;
                do {
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
                    if ((unsigned int)(w >> 32) != pv.epoch && clock64() - t0 > 8000000000ll) {
                        printf("ppo_b200: peer exchange timed out (rank %d waiting for rank %d, epoch %u)\n", pv.rank, r, pv.epoch);
                        __trap();
                    }
                } while ((unsigned int)(w >> 32) != pv.epoch);
                x = __uint_as_float((unsigned int)w);
            }
            sum += x;                                               // same numbers, same (rank) order on every GPU
        }
        g = sum;
    }
    if (e >= total) return;
    if (!p.apply) { p.reduced[e] = g; return; }
    if (is_net) {
        const float w = adam_apply(p.net, e, g, pm, pv, pw);
        const int ii = image_index(p.layout, e);
        if (ii >= 0) p.image[ii] = w;
    } else if (e < p.P + p.A) {
        if (is_ls) adam_apply(p.ls, e - p.P, g + (-p.ent_coeff), pm, pv, pw);   // src/ppo.cu:436-438
    } else {
        if (p.mode == kFusedValue) {
            *p.loss_slot += g / (float)p.m_total;
        } else {
            *p.loss_slot += -g / (float)p.m_total - p.ent_coeff * pw;
        }
    }
}

// (re)build the image from the flat parameters (after host uploads / non-fused updates).  The image was
// zero-filled at allocation, padding slots are never written, so only parameter slots are refreshed.
__global__ void __launch_bounds__(256) build_image_kernel(const float* __restrict__ params, float* __restrict__ image, FusedNet n) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n.P; e += gridDim.x * blockDim.x) {
        const int ii = image_index(n, e);
        if (ii >= 0) image[ii] = params[e];
    }
}

// ---- host side ------------------------------------------------------------------------------------
struct FusedPlan { bool ok; int tm, rt; size_t smem_bytes; FusedNet net; int g_off, red_off; int kind; };

// Plan for fused_tile64_kernel: every width <= 128 (64-column blocks), <= 8 outputs, everything in <= 220 KB of shared memory.
static FusedPlan make_plan64(NetDev* nd) {
    FusedPlan pl{};
    pl.ok = false;
    pl.kind = 1;
    const int L = nd->num_layers - 1;
    if (L < 1 || L > kFusedMaxLayers) return pl;
    FusedNet& n = pl.net;
    n.L = L;
    for (int l = 0; l <= L; l++) { n.sizes[l] = nd->sizes[l]; if (n.sizes[l] > 128 || n.sizes[l] < 1) return pl; }
    if (n.sizes[L] > 8) return pl;
    int off = 0, maxw = 4;
    for (int l = 0; l < L; l++) {
        n.acts[l] = nd->acts[l];
        n.w_off[l] = (int)nd->w_off[l];
        n.b_off[l] = (int)nd->b_off[l];
        n.wt_off[l] = off;
        n.ldw[l] = pad4(n.sizes[l + 1]) + 4;         // +4: consecutive Wt rows start 4 banks apart (dX reads)
        if (n.sizes[l + 1] <= 8) n.ldw[l] = 8;       // skinny layers read up to 8 columns per row
        off += n.sizes[l] * n.ldw[l];
    }
    for (int l = 0; l < L; l++) { n.bs_off[l] = off; off += std::max(8, pad4(n.sizes[l + 1])); }
    n.img_floats = (off + 31) & ~31;
    n.P = (int)nd->param_count;
    off = 0;
    for (int l = 0; l <= L; l++) {
        n.a_off[l] = off;
        off += std::max(pad4(n.sizes[l]), l == L ? 8 : 4) * kT64TMP;
        maxw = std::max(maxw, pad4(n.sizes[l]));
    }
    n.max_width_pad = maxw;
    off += n.img_floats;
    maxw = std::max(maxw, 64);  // the forward exchange uses 64 feature rows, the skinny forward 24
    n.max_width_pad = maxw;
    pl.g_off = off;            // E0 | E1 (maxw rows each): forward split-K exchange, out-of-place dX ping-pong, skinny scratch
    off += 2 * maxw * kT64TMP;
    pl.red_off = off;
    off += 64 + kT64TM + 4;
    pl.smem_bytes = (size_t)off * sizeof(float);
    pl.tm = kT64TM;
    pl.rt = 8;
    pl.ok = pl.smem_bytes <= 220 * 1024;
    return pl;
}

static FusedPlan choose_plan(NetDev* nd) { return make_plan64(nd); }

bool fused_supported(NeuralNetwork* nn) { return choose_plan(net_dev(nn)).ok; }

// Layout + up-to-date device pointer of the 64-wide weight image (shared with the rollout kernel).
static float* ensure_image(NetDev* nd, const FusedPlan& pl);
bool fused_image64(NeuralNetwork* nn, FusedNet* layout, const float** image) {
    NetDev* nd = net_dev(nn);
    const FusedPlan pl = choose_plan(nd);
    if (!pl.ok || pl.kind != 1) return false;
    *layout = pl.net;
    *image = ensure_image(nd, pl);
    return true;
}

static float* ensure_image(NetDev* nd, const FusedPlan& pl) {
    if (!nd->image || nd->image_floats != pl.net.img_floats) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (nd->image) CUDA_CHECK(cudaFree(nd->image));
        nd->image = dmalloc<float>(pl.net.img_floats);
        CUDA_CHECK(cudaMemsetAsync(nd->image, 0, (size_t)pl.net.img_floats * sizeof(float), stream()));
        nd->image_floats = pl.net.img_floats;
        nd->image_dirty = true;
    }
    if (nd->image_dirty) {
        B200_LAUNCH(build_image_kernel, std::max(1, std::min(64, div_up(pl.net.P, 256))), 256, 0, nd->params, nd->image, pl.net);
        nd->image_dirty = false;
    }
    return nd->image;
}

static unsigned long long* g_phase_dbg = nullptr;
static unsigned long long* phase_dbg() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PPO_B200_PHASE_DEBUG"); on = (e && e[0] == '1') ? 1 : 0; }
    if (on && !g_phase_dbg) { g_phase_dbg = dmalloc<unsigned long long>(16 * 65536); CUDA_CHECK(cudaMemset(g_phase_dbg, 0, 16 * 65536 * 8)); }
    return g_phase_dbg;
}

static void launch_fused(NetDev* nd, const FusedPlan& pl, FusedArgs& a, bool pdl = false) {
    a.dbg = (a.mode != kFusedForward) ? phase_dbg() : nullptr;
    a.net = pl.net;
    a.image = ensure_image(nd, pl);
    a.smem_g_off = pl.g_off;
    a.smem_red_off = pl.red_off;
    const int blocks = div_up(a.m, pl.tm);
    static size_t configured = 0;
    if (pl.smem_bytes > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(fused_tile64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
        configured = pl.smem_bytes;
    }
    B200_LAUNCH_PDL(fused_tile64_kernel, blocks, kT64Threads, pl.smem_bytes, pdl, a);
}

void fused_forward(NeuralNetwork* nn, const float* x, int m, float* y_out) {
    NetDev* nd = net_dev(nn);
    const FusedPlan pl = choose_plan(nd);
    if (!pl.ok) B200_FATAL("fused_forward on an unsupported net");
    FusedArgs a{};
    a.idx = nullptr; a.offset = 0; a.limit = m; a.m = m; a.m_total = m; a.mode = kFusedForward;
    a.state = x; a.y_out = y_out;
    launch_fused(nd, pl, a);
}

static AdamSeg make_seg(float* w, float* g, Adam* adam, float lr) {
    AdamSeg s{};
    s.w = w; s.g = g; s.m = adam->m; s.v = adam->v;
    const float bc1 = 1 - powf(adam->beta1, adam->time_step);      // src/adam.cu:56-59
    s.bc2 = 1 - powf(adam->beta2, adam->time_step);
    s.step_size = lr / bc1;
    s.beta1 = adam->beta1; s.beta2 = adam->beta2;
    s.omb1 = 1 - adam->beta1; s.omb2 = 1 - adam->beta2;
    return s;
}

// One fused minibatch update.  policy == nullptr: value net (MSE on adv_target); else the policy.
// reduced_out != nullptr (data-parallel path): no Adam; the reduced slab [grads | grad_log_std | loss
// term] is written there and the caller all-reduces and applies the optimiser.
// Returns false when the net is outside the fused kernel's limits (caller uses the layer-wise path).
static bool pdl_enabled() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("PPO_B200_PDL"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}

// `chained`: the kernel that precedes this call in the stream is the fused_reduce_adam_kernel of the previous
// minibatch of the same net (so the update kernel may be launched programmatically dependent on it).
bool fused_minibatch_update(NeuralNetwork* nn, GaussianPolicy* policy, Adam* adam_net, Adam* adam_ls, float lr,
                            const int* perm, int offset, int limit, int m, int m_total, const TrajectoryBuffer* b,
                            float epsilon, float ent_coeff, float* loss_slot, float* reduced_out, bool chained, bool dp_peer) {
    NetDev* nd = net_dev(nn);
    const FusedPlan pl = choose_plan(nd);
    if (!pl.ok) return false;
    const int A = policy ? policy->action_size : 1;
    const int slab = (int)nd->param_count + A + 1;
    const int blocks = div_up(m, pl.tm);
    const size_t need = (size_t)blocks * slab;
    if (need > nd->partials_cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (nd->partials) CUDA_CHECK(cudaFree(nd->partials));
        nd->partials = dmalloc<float>(need);
        nd->partials_cap = need;
    }
    FusedArgs a{};
    a.idx = perm; a.offset = offset; a.limit = limit; a.m = m; a.m_total = m_total;
    a.mode = policy ? kFusedPolicy : kFusedValue;
    a.state = b->d_state_p; a.action = b->d_action_p; a.logprob = b->d_logprob_p;
    a.advantage = b->d_advantage_p; a.adv_target = b->d_adv_target_p;
    a.log_std = policy ? policy->d_log_std : nullptr;
    a.epsilon = epsilon; a.ent_coeff = ent_coeff;
    a.partials = nd->partials; a.slab = slab;
    const bool image_clean = nd->image && !nd->image_dirty && nd->image_floats == pl.net.img_floats;
    const bool pdl = pdl_enabled() && pl.kind == 1 && !reduced_out;
    launch_fused(nd, pl, a, pdl && chained && image_clean);
    nd->last_splits = blocks;

    ReduceAdamArgs r{};
    r.partials = nd->partials; r.nparts = blocks; r.slab = slab; r.P = (int)nd->param_count; r.A = A;
    r.mode = a.mode; r.m_total = m_total; r.ent_coeff = ent_coeff; r.loss_slot = loss_slot;
    r.log_std = a.log_std;
    r.layout = pl.net; r.image = nd->image;
    r.apply = reduced_out ? 0 : 1;
    r.reduced = reduced_out;
    if (dp_peer) {
        r.peer = dist_peer_next((size_t)slab);
        if (!r.peer.ready) B200_FATAL("peer exchange requested but the peer arena is not available (slab %d floats)", slab);
    }
    if (!reduced_out) {
        adam_net->time_step += 1;
        r.net = make_seg(nd->params, nd->grads, adam_net, lr);
        if (policy) {
            adam_ls->time_step += 1;
            r.ls = make_seg(policy->d_log_std, policy->d_log_std_grad, adam_ls, lr);
        }
    } else {
        nd->image_dirty = true;     // the caller updates the parameters with the generic Adam kernel
    }
    B200_LAUNCH_PDL(fused_reduce_adam_kernel, div_up(slab, 32), 32 * kRedWarps, 0, pdl, r);
    return true;
}

}  // namespace b200

// debug: copy the phase timestamps of the last fused_tile64_kernel launch ([blocks][16] u64) to the host
extern "C" void ppo_b200_debug_phase_stamps(unsigned long long* out, int blocks) {
    if (!b200::g_phase_dbg) return;
    CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    CUDA_CHECK(cudaMemcpy(out, b200::g_phase_dbg, (size_t)blocks * 16 * 8, cudaMemcpyDeviceToHost));
}
