// fused_mlp.cu — the small-net fast path (SURVEY.md §8 rows a6/a7 + a9..a14, (e)).
//
// For the reference-width actor/critic nets (every layer width <= 128: 2x64, the reference's default 2x128) the update phase
// runs in PERSISTENT PHASE KERNELS: one cooperative launch executes ALL minibatches of all value epochs (or all policy epochs)
//   fused_phase_spec64_kernel<ACT>   shape-specialised for S <= 8 -> 64 -> 64 -> 1 (BASELINE.json configs[1]): 128-row tiles, 8 x 8
//                                    register tiles, compile-time loop bounds
//   fused_phase_kernel               every other net with widths <= 128 (64-row tile slots, 1 or 2 per CTA)
// and per minibatch step does  [A] gather rows by permutation index -> forward through ALL layers -> fused loss head (MSE, or
// Gaussian log-prob + PPO-clip surrogate) -> backward through all layers -> one gradient slab per CTA;  [B] grid barrier;
// [C] every CTA reduces its slice of the parameter vector over the slabs in fixed order (+ the cross-GPU sum over NVLink peer
// memory under data parallelism), applies Adam and refreshes its entries of the pre-transposed weight image;  [D] grid barrier;
// [E] TMA re-stage of the image.  Weights and activations live in shared memory; HBM sees the gathered rows and the slabs.
//
// The one-launch-per-minibatch pair the phase kernels grew out of is kept (GAE value forwards, PPO_B200_PERSISTENT=0 A/B runs,
// data-parallel runs without peer memory):
//   fused_tile64_kernel       [A] for one minibatch, 64-row tiles
//   fused_reduce_adam_kernel  [C] for one minibatch, chained with programmatic dependent launch
// versus ~14 launches through the layer-wise kernels of gemm.cu/policy.cu/adam.cu (which remain the generic path for wider
// nets).  The reference does this with ~25 launches, 1-3 blocking D2H reads and 1-2 cudaMallocs per minibatch
// (src/ppo.cu:495-532).
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

// Barrier over the 256 threads that own one 64-row tile.  The one-tile-per-CTA kernel passes id 0 (== __syncthreads for its
// 256-thread CTA); the persistent phase kernel runs two tiles side by side in one 512-thread CTA on ids 1 and 2.
__device__ __forceinline__ void tile_sync(int bar) { asm volatile("bar.sync %0, 256;" :: "r"(bar) : "memory"); }

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
// Contains one tile_sync.
__device__ __forceinline__ void t64_forward(const float* __restrict__ Xt, const float* __restrict__ Wt, int ldw,
                                            const float* __restrict__ bias, float* __restrict__ Yt, float* __restrict__ xch,
                                            int n_in, int n_out, int act, int lt, int g, int cb, int bar) {
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
    tile_sync(bar);
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
// Contains one tile_sync: every thread of the tile must call it.
__device__ __forceinline__ void t64_forward_skinny(const float* __restrict__ Xt, const float* __restrict__ Wt, int ldw,
                                                   const float* __restrict__ bias, float* __restrict__ Yt, float* __restrict__ scratch,
                                                   int n_in, int n_out, int act, int tid, int bar) {
    const int r = tid & 63, h = tid >> 6;      // h = 0..3
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
    tile_sync(bar);
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
                                                     float* __restrict__ gW, int n_in, int n_out, int lt, int jb, int kb, bool accum) {
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
            if (k < n_in) {
                float* dst = gW + (size_t)j * n_in + k;
                const float v = acc[a][b].x + acc[a][b].y;
                *dst = accum ? *dst + v : v;        // accum: this (CTA, tile slot) adds a further tile of the same minibatch
            }
        }
    }
}

// Skinny dW.  WIDE_IS_K = true : n_out <= 8 (value head / action mean): thread = (k, row half), acc[j]
//            WIDE_IS_K = false: n_in  <= 8 (first layer of a low-dimensional env): thread = (j, row half), acc[k]
// The two row halves meet through one shuffle; sums run in a fixed order.
template <bool WIDE_IS_K>
__device__ __forceinline__ void t64_backward_weights_skinny(const float* __restrict__ Gt, const float* __restrict__ Xt,
                                                            float* __restrict__ gW, int n_in, int n_out, int lt, int wb, bool accum) {
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
                float* dst = WIDE_IS_K ? gW + (size_t)q * n_in + wide : gW + (size_t)wide * n_in + q;
                *dst = accum ? *dst + acc[q] : acc[q];
            }
    }
}

__device__ __forceinline__ void t64_weights_dispatch(const float* Gt, const float* Xt, float* gW, int n_in, int n_out, int lt, bool accum) {
    // few tile shapes only (instruction-cache footprint); layers wider than 64 go 64x64 block by block
    if (n_out <= 8) {
        for (int wb = 0; wb < n_in; wb += 64) t64_backward_weights_skinny<true>(Gt, Xt, gW, n_in, n_out, lt, wb, accum);
    } else if (n_in <= 8) {
        for (int wb = 0; wb < n_out; wb += 64) t64_backward_weights_skinny<false>(Gt, Xt, gW, n_in, n_out, lt, wb, accum);
    } else {
        for (int jb = 0; jb < n_out; jb += 64)
            for (int kb = 0; kb < n_in; kb += 64) {
                if (n_out - jb <= 16) t64_backward_weights<1, 8>(Gt, Xt, gW, n_in, n_out, lt, jb, kb, accum);
                else t64_backward_weights<4, 8>(Gt, Xt, gW, n_in, n_out, lt, jb, kb, accum);
            }
    }
}

// gb[j] = sum_r Gt[j][r]: thread = (j, row half), fixed-order sums
__device__ __forceinline__ void t64_bias_grad(const float* __restrict__ Gt, float* __restrict__ gb, int n_out, int lt, bool accum) {
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
        if (h == 0 && j < n_out) gb[j] = accum ? gb[j] + s : s;
    }
}

// ---- fused loss head (src/loss.cu:5-23 | src/policy.cu:67-111 + src/ppo.cu:89-98): dLoss/dy written IN PLACE over the
// output tile Yt (rows >= OUT of it stay zero); the tile's loss term and log_std gradient go to the slab tail.
// Threads tid < 64 own one row each; contains one tile_sync.
struct HeadCtx { int mode, m_total, OUT, P, out_act; const float* log_std; float epsilon; };
__device__ __forceinline__ void t64_loss_head(const HeadCtx& c, float* __restrict__ Yt, float* __restrict__ red, float* __restrict__ slab,
                                              int tid, int bar, bool row_valid, float h_target, float h_adv, float h_lp_old,
                                              const float (&h_act)[8], bool accum) {
    constexpr int TMP = kT64TMP;
    const int OUT = c.OUT;
    float loss_term = 0.f;
    float gls[8];
#pragma unroll
    for (int j = 0; j < 8; j++) gls[j] = 0.f;
    if (tid < kT64TM) {
        float gout[8];
#pragma unroll
        for (int j = 0; j < 8; j++) gout[j] = 0.f;
        float yv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) yv[j] = (j < OUT) ? Yt[j * TMP + tid] : 0.f;
        if (row_valid) {
            if (c.mode == kFusedValue) {            // src/loss.cu:5-23
                gout[0] = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(yv[0], h_target)), (float)c.m_total);
                const float d = __fsub_rn(h_target, yv[0]);
                loss_term = __fmul_rn(d, d);
            } else {                                // src/policy.cu:67-111 + src/ppo.cu:89-98
                float lsv[8];                       // L2 loads: another SM's Adam may have just updated log_std (persistent kernel)
#pragma unroll
                for (int j = 0; j < 8; j++) lsv[j] = (j < OUT) ? __ldcg(c.log_std + j) : 0.f;
                const float lp = fused_log_prob(yv, lsv, h_act, OUT);
                const float ratio = expf(__fsub_rn(lp, h_lp_old));
                const bool adv_pos = h_adv > 0.f;
                const bool hi = ratio > 1.f + c.epsilon, lo = ratio < 1.f - c.epsilon;
                const float sel = adv_pos ? (hi ? 1.f + c.epsilon : ratio) : (lo ? 1.f - c.epsilon : ratio);
                loss_term = __fmul_rn(h_adv, sel);
                const int keep = adv_pos ? !hi : !lo;
                const float g = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), h_adv), ratio), (float)c.m_total);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (j < OUT) {
                        const float e2 = expf(-2.f * lsv[j]);
                        const float diff = __fsub_rn(h_act[j], yv[j]);
                        gout[j] = __fmul_rn(__fmul_rn(diff, e2), g);
                        gls[j] = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), g);
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (j < OUT) Yt[j * TMP + tid] = act_grad(yv[j], gout[j], c.out_act);
        const int warp = tid >> 5, lane = tid & 31;       // warps 0..1 hold the data
        const float v = warp_sum(loss_term);
        if (lane == 0) red[warp] = v;
        if (c.mode == kFusedPolicy) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < OUT) { const float s = warp_sum(gls[j]); if (lane == 0) red[8 + j * 2 + warp] = s; }
        }
    }
    tile_sync(bar);
    if (tid == 0) { const float v = red[0] + red[1]; slab[c.P + OUT] = accum ? slab[c.P + OUT] + v : v; }
    if (c.mode == kFusedPolicy && tid < OUT) {
        const float v = red[8 + tid * 2] + red[8 + tid * 2 + 1];
        slab[c.P + tid] = accum ? slab[c.P + tid] + v : v;
    }
}

// ---- single-output nets (value head; Pendulum's action mean): last layer + loss head + dX of the last layer in ONE pass.
// Four threads per row (lane = (row % 8) * 4 + q, q = interleaved quarter of k):
//   y_r = sum_k H[k][r] w[k] + b  (quarters meet through two shuffles)  ->  loss head on y_r (all four threads, redundantly)
//   ->  g_r = dLoss/dy_r through the output activation  ->  Yt[0][r] = g_r (dW / db of this layer read it afterwards)
//   ->  Gout[k][r] = g_r * w[k] * act'(H[k][r])  for the thread's own k's (the dX tile the next backward layer consumes).
// Replaces: skinny forward (smem exchange + barrier) -> head on 64 threads (+ barrier) -> a 64x64-tile dX with one live row.
// Contains one tile_sync.
__device__ __forceinline__ void t64_head_fused1(const HeadCtx& c, const float* __restrict__ H, const float* __restrict__ Wt, int ldw,
                                                const float* __restrict__ bias, float* __restrict__ Yt, float* __restrict__ Gout,
                                                float* __restrict__ red, float* __restrict__ slab, int n_in, int act_prev, int tid, int bar,
                                                bool row_valid, float h_target, float h_adv, float h_lp_old, float h_act0, bool accum) {
    constexpr int TMP = kT64TMP;
    const int r = tid >> 2, q = tid & 3;
    float part = 0.f;
    for (int k = q; k < n_in; k += 4) part = fmaf(H[k * TMP + r], Wt[k * ldw], part);
    part += __shfl_xor_sync(kFull, part, 1);
    part += __shfl_xor_sync(kFull, part, 2);
    const float y = act_apply(part + bias[0], c.out_act);
    float loss_term = 0.f, gout = 0.f, gls = 0.f;
    if (row_valid) {
        if (c.mode == kFusedValue) {            // src/loss.cu:5-23
            gout = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(y, h_target)), (float)c.m_total);
            const float d = __fsub_rn(h_target, y);
            loss_term = __fmul_rn(d, d);
        } else {                                // src/policy.cu:67-111 + src/ppo.cu:89-98
            const float ls = __ldcg(c.log_std);
            const float lp = fused_log_prob(&y, &ls, &h_act0, 1);
            const float ratio = expf(__fsub_rn(lp, h_lp_old));
            const bool adv_pos = h_adv > 0.f;
            const bool hi = ratio > 1.f + c.epsilon, lo = ratio < 1.f - c.epsilon;
            const float sel = adv_pos ? (hi ? 1.f + c.epsilon : ratio) : (lo ? 1.f - c.epsilon : ratio);
            loss_term = __fmul_rn(h_adv, sel);
            const int keep = adv_pos ? !hi : !lo;
            const float g = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), h_adv), ratio), (float)c.m_total);
            const float e2 = expf(-2.f * ls);
            const float diff = __fsub_rn(h_act0, y);
            gout = __fmul_rn(__fmul_rn(diff, e2), g);
            gls = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), g);
        }
    }
    const float g = act_grad(y, gout, c.out_act);
    if (q == 0) Yt[r] = g;
    for (int k = q; k < pad4(n_in); k += 4)
        Gout[k * TMP + r] = (k < n_in) ? act_grad(H[k * TMP + r], __fmul_rn(g, Wt[k * ldw]), act_prev) : 0.f;
    {   // tile sums of the loss term / log_std gradient: one contribution per row (q == 0), warps in fixed order
        const int warp = tid >> 5, lane = tid & 31;
        const float v = warp_sum(q == 0 ? loss_term : 0.f);
        if (lane == 0) red[warp] = v;
        if (c.mode == kFusedPolicy) { const float s2 = warp_sum(q == 0 ? gls : 0.f); if (lane == 0) red[8 + warp] = s2; }
    }
    tile_sync(bar);
    if (tid == 0) {
        float v = red[0];
#pragma unroll
        for (int w = 1; w < 8; w++) v += red[w];
        slab[c.P + 1] = accum ? slab[c.P + 1] + v : v;
    }
    if (c.mode == kFusedPolicy && tid == 32) {
        float v = red[8];
#pragma unroll
        for (int w = 1; w < 8; w++) v += red[8 + w];
        slab[c.P] = accum ? slab[c.P] + v : v;
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
            t64_forward_skinny(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, scratch, net.sizes[l], net.sizes[l + 1], net.acts[l], tid, 0);
        else
            for (int cb = 0; cb < pad4(net.sizes[l + 1]); cb += 64) {
                if (cb) __syncthreads();           // the previous block's finalisation has read the exchange buffer
                t64_forward(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, ebuf, net.sizes[l], net.sizes[l + 1], net.acts[l], lt, grp, cb, 0);
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
    {
        HeadCtx hc;
        hc.mode = p.mode; hc.m_total = p.m_total; hc.OUT = OUT; hc.P = net.P; hc.out_act = net.acts[net.L - 1];
        hc.log_std = p.log_std; hc.epsilon = p.epsilon;
        t64_loss_head(hc, Yt, red, slab, tid, 0, my_src >= 0, h_target, h_adv, h_lp_old, h_act, false);
    }
    t64_stamp(p, 9);
    // ---- backward: group 0 -> dW_l, db_l; group 1 -> dX_l into the ping-pong buffer (both read G_{l+1}, A_l)
    const float* G = Yt;
    for (int l = net.L - 1; l >= 0; l--) {
        const int n_in = net.sizes[l], n_out = net.sizes[l + 1];
        const float* Xt = act0 + net.a_off[l];
        float* Gout = ebuf + ((net.L - 1 - l) & 1) * (net.max_width_pad * TMP);
        if (grp == 0) {
            t64_weights_dispatch(G, Xt, slab + net.w_off[l], n_in, n_out, lt, false);
            t64_bias_grad(G, slab + net.b_off[l], n_out, lt, false);
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

// ===================================================================================================
// Persistent phase kernel: ALL minibatches of all epochs of one phase (value or policy) in ONE cooperative launch.
//
// The reference's update is strictly sequential SGD (minibatch k+1 needs the weights after Adam step k,
// src/ppo.cu:495-532), so the only parallelism is inside a minibatch.  One CTA per SM stays resident for the whole phase:
//   [A] every CTA runs its tiles of the minibatch (two 64-row tiles side by side: 2 x 256 threads on named barriers 1/2;
//       further tiles of the same minibatch are accumulated into the same slab) with the tile code above: gather ->
//       forward -> loss head -> backward, activations in shared memory, weights from the shared image;
//   [B] grid barrier; [C] CTA c reduces ITS 1/gridDim slice of the parameter vector over all slabs in fixed order
//       (L2 reads), exchanges it with the other GPUs over NVLink peer memory under data parallelism, runs Adam on the
//       slice (reference arithmetic, src/adam.cu:56-69) and refreshes those entries of the global weight image;
//   [D] grid barrier; [E] every CTA re-stages the image with one TMA bulk copy.
// The gather of the NEXT minibatch (permutation lookup, cp.async of the state rows, scalar loads) is issued before [B], so
// it is in flight during [B]-[E].  Versus round 1 (two launches per minibatch, slabs through a second kernel): no launch
// or programmatic-dependency latency per minibatch, half as many image loads, and the reduction + Adam of a minibatch
// costs two grid barriers and one L2 round trip instead of a kernel.
// Determinism: slab order, slice ownership and the rank order of the cross-GPU sum are fixed -> bitwise reproducible.
// ===================================================================================================
struct PhaseArgs {
    FusedNet net;
    float* image;                 // global weight image (TMA source; refreshed by [C])
    const int* perms;             // [n_epochs][limit] permutations (null: identity)
    int limit, mb, m_total, row0, batch_stride, num_batches, n_steps, mode;
    const float *state, *action, *logprob, *advantage, *adv_target;
    const float* log_std;         // device log_std (policy mode): read by the heads, updated through ls.w
    float epsilon, ent_coeff;
    float* partials;              // [nslabs][slab]
    int slab, P, A;
    int sub_floats, g_off, red_off;   // shared-memory floats per tile slot; offsets of E0|E1 and of the reduction scratch
    AdamSeg netseg, ls;           // step_size / bc2 come from coef[step]
    const float4* coef;           // [n_steps] {step_size_net, bc2_net, step_size_ls, bc2_ls} (host powf, src/adam.cu:56-59)
    float* loss_slot;
    unsigned int* barrier;        // grid barrier counter, zero at launch
    PeerView peer;                // data parallelism: parity-0 lanes + parity stride; exchange s uses epoch + s
    long long spin_limit;         // clock64 ticks a barrier / peer poll may wait before the kernel traps
    unsigned long long* dbg;
};

__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int r;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

// All threads of all CTAs call it the same number of times (cooperative launch: every CTA is resident).
__device__ __forceinline__ void phase_grid_barrier(unsigned int* counter, unsigned int target, long long spin_limit) {
    if (gridDim.x == 1) {            // one CTA (small minibatches): a block barrier is the grid barrier
        __threadfence();
        __syncthreads();
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // release the writes of this CTA (ordered by the bar.sync above)
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (ld_acquire_u32(counter) < target) {
            if (clock64() - t0 > spin_limit) {
                printf("ppo_b200: grid barrier timed out (block %d, target %u, counter %u)\n", blockIdx.x, target, *counter);
                __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- [C] of the persistent phase kernels: slice reduction + (cross-GPU sum) + Adam + image refresh.
// CTA c reduces ITS 1/gridDim slice of the parameter vector over all slabs in fixed order, exchanges it with the other GPUs over
// NVLink peer memory under data parallelism, runs Adam on the slice (src/adam.cu:56-69) and refreshes those entries of the global
// weight image (and of the staged one when the grid is a single CTA).  Contains one __syncthreads on the many-CTA path.
__device__ __forceinline__ void phase_reduce_adam(const PhaseArgs& p, int s, int nslabs, float* img, float (*redw)[33], float s_entropy_v) {
    const int total = p.P + p.A + 1;
    const int G_ = gridDim.x;
    const int chunk = 32 * ((total + 32 * G_ - 1) / (32 * G_));
    const int e0 = blockIdx.x * chunk, e1 = min(total, e0 + chunk);
    const int ngroups = e1 > e0 ? (e1 - e0 + 31) >> 5 : 0;
    const int W = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 cf = __ldg(p.coef + s);
    const bool single = gridDim.x == 1;
    // optimiser state of element e (only ever touched by the thread that owns e, in every step)
    auto load_state = [&](int e, float& pm, float& pv, float& pw) {
        pm = 0.f; pv = 0.f; pw = 0.f;
        if (e < p.P) { pm = p.netseg.m[e]; pv = p.netseg.v[e]; pw = p.netseg.w[e]; }
        else if (e < p.P + p.A) { if (p.mode == kFusedPolicy) { pm = p.ls.m[e - p.P]; pv = p.ls.v[e - p.P]; pw = p.ls.w[e - p.P]; } }
        else if (p.mode == kFusedPolicy) pw = s_entropy_v;
    };
    // local sum g of element e -> (cross-GPU sum) -> Adam / loss accumulation
    auto finalize = [&](int e, float g, float pm, float pv, float pw) {
        if (p.peer.ready) {
            // gradient exchange over NVLink peer memory (dist.cu "peer arena"): push {tag, value}, poll, ordered sum
            const PeerView& pvw = p.peer;
            const unsigned int epoch = pvw.epoch + (unsigned int)s;
            const size_t poff = (size_t)(epoch & 1u) * pvw.parity_stride;
            const unsigned long long packed = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(g);
            for (int r = 0; r < pvw.world; r++)
                if (r != pvw.rank) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(pvw.peer_recv[r] + poff + e), "l"(packed) : "memory");
            const long long tstart = clock64();
            float acc = 0.f;
            for (int r = 0; r < pvw.world; r++) {
                float x = g;
                if (r != pvw.rank) {
                    const unsigned long long* srcw = pvw.my_recv + poff + (size_t)r * kPeerCap + e;
                    unsigned long long w;
                    do {
                        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(srcw) : "memory");
                        if ((unsigned int)(w >> 32) != epoch && clock64() - tstart > p.spin_limit) {
                            printf("ppo_b200: peer exchange timed out (rank %d waiting for rank %d, exchange %u)\n", pvw.rank, r, epoch);
                            __trap();
                        }
                    } while ((unsigned int)(w >> 32) != epoch);
                    x = __uint_as_float((unsigned int)w);
                }
                acc += x;                      // same numbers, same (rank) order on every GPU
            }
            g = acc;
        }
        if (e < p.P) {
            AdamSeg sg = p.netseg;
            sg.step_size = cf.x; sg.bc2 = cf.y;
            const float w = adam_apply(sg, e, g, pm, pv, pw);
            const int ii = image_index(p.net, e);
            if (ii >= 0) {
                p.image[ii] = w;
                if (single) img[ii] = w;       // one CTA: the staged image is refreshed in place, no re-stage
            }
        } else if (e < p.P + p.A) {
            if (p.mode == kFusedPolicy) {
                AdamSeg sg = p.ls;
                sg.step_size = cf.z; sg.bc2 = cf.w;
                adam_apply(sg, e - p.P, g + (-p.ent_coeff), pm, pv, pw);     // src/ppo.cu:436-438
            }
        } else {
            if (p.mode == kFusedValue) *p.loss_slot += g / (float)p.m_total;
            else *p.loss_slot += -g / (float)p.m_total - p.ent_coeff * pw;
        }
    };
    if (ngroups >= W) {
        // few CTAs, many elements each (small minibatches): one thread per element, four elements in flight
        for (int base = e0 + (int)threadIdx.x; base < e1; base += 4 * (int)blockDim.x) {
            float g[4], pm[4], pv[4], pw[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = base + u * (int)blockDim.x;
                g[u] = 0.f;
                if (e < e1) load_state(e, pm[u], pv[u], pw[u]);
            }
            for (int b = 0; b < nslabs; b++) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int e = base + u * (int)blockDim.x;
                    if (e < e1) g[u] += __ldcg(p.partials + (size_t)b * p.slab + e);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = base + u * (int)blockDim.x;
                if (e < e1) finalize(e, g[u], pm[u], pv[u], pw[u]);
            }
        }
    } else {
        // many CTAs, <= W 32-element groups each: the warps split the slabs, fixed-order combine through shared memory
        int ways = 1;
        while (ways * 2 * max(ngroups, 1) <= W) ways *= 2;
        const int per = (nslabs + ways - 1) / ways;
        const int gi = warp / ways, part = warp % ways;
        const int e = e0 + gi * 32 + lane;
        const bool live = gi < ngroups && e < e1;
        const bool fin = live && part == 0;
        float pm = 0.f, pv = 0.f, pw = 0.f;
        if (fin) load_state(e, pm, pv, pw);        // overlaps the slab loads
        float sum = 0.f;
        if (live) {
            const int b0 = part * per, b1 = min(nslabs, b0 + per);
            const float* src = p.partials + e;
            // up to 24 independent L2 loads in flight per lane (296 slabs / 16 warps = 19: ONE round trip), summed in slab order
            for (int b = b0; b < b1; b += 24) {
                const int cnt = min(24, b1 - b);
                float t[24];
#pragma unroll
                for (int u = 0; u < 24; u++) t[u] = (u < cnt) ? __ldcg(src + (size_t)(b + u) * p.slab) : 0.f;
#pragma unroll
                for (int u = 0; u < 24; u++) sum += t[u];
            }
        }
        redw[warp][lane] = sum;
        __syncthreads();
        if (fin) {
            float g = redw[warp][lane];
            for (int q = 1; q < ways; q++) g += redw[warp + q][lane];
            finalize(e, g, pm, pv, pw);
        }
    }
}

constexpr int kPhaseMaxThreads = 512;
__global__ void __launch_bounds__(kPhaseMaxThreads, 1) fused_phase_kernel(const PhaseArgs p) {
    constexpr int TM = kT64TM, TMP = kT64TMP;
    extern __shared__ __align__(128) float smem[];
    __shared__ float redw[kPhaseMaxThreads / 32][33];
    __shared__ float s_entropy;       // entropy of the policy BEFORE this step's Adam (the loss of step s is reported with it)
    const FusedNet& net = p.net;
    const int nsub = blockDim.x >> 8;
    const int sub = threadIdx.x >> 8, tid = threadIdx.x & 255, lt = tid & 127, grp = tid >> 7;
    const int bar = 1 + sub;
    const int S = net.sizes[0], OUT = net.sizes[net.L], SP = pad4(S);
    float* img = smem;
    float* act0 = smem + net.img_floats + sub * p.sub_floats;
    float* ebuf = act0 + p.g_off;
    float* red = act0 + p.red_off;                        // 64 floats
    int* src_rows = reinterpret_cast<int*>(red + 64);     // TM ints
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + net.img_floats + nsub * p.sub_floats);
    float* scratch = ebuf;
    float* Xt0 = act0 + net.a_off[0];

    const int n_tiles = (p.mb + TM - 1) / TM;
    const int stride = gridDim.x * nsub;
    const int t0 = blockIdx.x * nsub + sub;
    const int rounds = t0 < n_tiles ? (n_tiles - t0 + stride - 1) / stride : 0;
    const int nslabs = min(n_tiles, stride);
    float* slab = p.partials + (size_t)t0 * p.slab;
    unsigned int bar_gen = 0;

    HeadCtx hc;
    hc.mode = p.mode; hc.m_total = p.m_total; hc.OUT = OUT; hc.P = net.P; hc.out_act = net.acts[net.L - 1];
    hc.log_std = p.log_std; hc.epsilon = p.epsilon;
    const bool head1 = OUT == 1 && net.L >= 2;            // single-output nets: fused last layer + head + dX (t64_head_fused1)
    const int hrow = head1 ? (tid >> 2) : tid;            // the row whose per-row scalars this thread holds
    const bool hholds = head1 || tid < TM;

    // source row of buffer for row r of tile t at step s (src/trajectory_buffer.cu:208-209)
    auto src_of = [&](int s, int t, int r) -> int {
        const int row = t * TM + r;
        if (row >= p.mb) return -1;
        const int e = s / p.num_batches, k = s - e * p.num_batches;
        const int off = (k * p.batch_stride + p.row0 + row) % p.limit;
        return p.perms ? __ldg(p.perms + (size_t)e * p.limit + off) : off;
    };
    float h_target = 0.f, h_adv = 0.f, h_lp_old = 0.f, h_act[8];
#pragma unroll
    for (int j = 0; j < 8; j++) h_act[j] = 0.f;
    int my_src = -1;
    // issue the gather of one tile: state rows by cp.async into Xt0 (feature-major), per-row scalars into registers
    auto issue_gather = [&]() {
        for (int e = tid; e < TM * SP; e += 256) {
            const int r = e / SP, k = e - r * SP;
            const int src = src_rows[r];
            float* dst = Xt0 + k * TMP + r;
            if (src >= 0 && k < S) cp_async4(dst, p.state + (size_t)src * S + k);
            else *dst = 0.f;
        }
        if (hholds) {
            my_src = src_rows[hrow];
            h_target = 0.f; h_adv = 0.f; h_lp_old = 0.f;
            if (my_src >= 0) {
                if (p.mode == kFusedValue) {
                    h_target = __ldg(p.adv_target + my_src);
                } else {
                    h_adv = __ldg(p.advantage + my_src);
                    h_lp_old = __ldg(p.logprob + my_src);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j < OUT) h_act[j] = __ldg(p.action + (size_t)my_src * OUT + j);
                }
            }
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    auto stamp = [&](int slot) {      // PPO_B200_PHASE_DEBUG: globaltimer of tile slot 0, last item wins
        if (p.dbg && sub == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            p.dbg[(size_t)blockIdx.x * 16 + slot] = t;
        }
    };
    int nxt_src = -1;
    if (rounds > 0) {
        if (tid < TM) src_rows[tid] = src_of(0, t0, tid);
        tile_sync(bar);
        issue_gather();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar, (uint32_t)net.img_floats * 4u);
        tma_bulk_g2s(img, p.image, (uint32_t)net.img_floats * 4u, mbar);
    }
    uint32_t img_parity = 0;
    mbar_wait(mbar, img_parity);
    img_parity ^= 1;

    for (int s = 0; s < p.n_steps; s++) {
        // ---- [A] this slot's tiles of minibatch s
        for (int rd = 0; rd < rounds; rd++) {
            const bool accum = rd > 0;
            cp_async_wait_all();
            tile_sync(bar);                                // the gathered tile is complete and visible
            int ns = s, nt = t0 + (rd + 1) * stride;       // the item after this one
            if (rd + 1 >= rounds) { ns = s + 1; nt = t0; }
            const bool have_next = ns < p.n_steps;
            if (have_next && tid < TM) nxt_src = src_of(ns, nt, tid);     // in flight during the tile
            stamp(3);
            const int n_fwd = head1 ? net.L - 1 : net.L;
            for (int l = 0; l < n_fwd; l++) {
                const float* Xt = act0 + net.a_off[l];
                float* Yt = act0 + net.a_off[l + 1];
                if (net.sizes[l + 1] <= 8)
                    t64_forward_skinny(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, scratch, net.sizes[l], net.sizes[l + 1], net.acts[l], tid, bar);
                else
                    for (int cb = 0; cb < pad4(net.sizes[l + 1]); cb += 64) {
                        if (cb) tile_sync(bar);
                        t64_forward(Xt, img + net.wt_off[l], net.ldw[l], img + net.bs_off[l], Yt, ebuf, net.sizes[l], net.sizes[l + 1], net.acts[l], lt, grp, cb, bar);
                    }
                tile_sync(bar);
                stamp(8 + l);
            }
            float* Yt = act0 + net.a_off[net.L];
            if (head1) {
                const int lh = net.L - 1;
                t64_head_fused1(hc, act0 + net.a_off[lh], img + net.wt_off[lh], net.ldw[lh], img + net.bs_off[lh], Yt, ebuf, red, slab,
                                net.sizes[lh], net.acts[lh - 1], tid, bar, my_src >= 0, h_target, h_adv, h_lp_old, h_act[0], accum);
            } else {
                t64_loss_head(hc, Yt, red, slab, tid, bar, my_src >= 0, h_target, h_adv, h_lp_old, h_act, accum);
            }
            stamp(11);
            const float* G = Yt;
            for (int l = net.L - 1; l >= 0; l--) {
                const int n_in = net.sizes[l], n_out = net.sizes[l + 1];
                const float* Xt = act0 + net.a_off[l];
                float* Gout = ebuf + ((net.L - 1 - l) & 1) * (net.max_width_pad * TMP);
                const bool dx_done = head1 && l == net.L - 1;      // the fused head already wrote this layer's dX into E0
                if (dx_done) {
                    // nothing left for the dX group at this layer: it takes the (small) dW / db of the last layer and then goes
                    // straight on to dX of the layer below, while the dW group starts on that layer's dW: balanced halves
                    if (grp == 1) {
                        t64_weights_dispatch(G, Xt, slab + net.w_off[l], n_in, n_out, lt, accum);
                        t64_bias_grad(G, slab + net.b_off[l], n_out, lt, accum);
                    }
                } else if (grp == 0) {
                    t64_weights_dispatch(G, Xt, slab + net.w_off[l], n_in, n_out, lt, accum);
                    t64_bias_grad(G, slab + net.b_off[l], n_out, lt, accum);
                } else if (l > 0) {
                    for (int kb = 0; kb < pad4(n_in); kb += 64)
                        t64_backward_input(G, img + net.wt_off[l], net.ldw[l], Xt, Gout, n_in, n_out, net.acts[l - 1], lt, kb);
                }
                if (!dx_done) tile_sync(bar);              // (also after l == 0: Xt0 / src_rows are about to be refilled)
                if (l < 3) stamp(12 + l);
                G = Gout;
            }
            if (have_next) {
                if (tid < TM) src_rows[tid] = nxt_src;
                tile_sync(bar);
                issue_gather();
            }
        }
        if (p.dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(size_t)blockIdx.x * 16 + 4] = t; }
        if (threadIdx.x == 33 && p.mode == kFusedPolicy) {    // log_std is stable here: last written in [C] of step s-1, before [D]
            float ent = (float)(p.A * 0.5 * (1 + log(2 * kPiF)));             // src/policy.cu:171-178
            for (int j = 0; j < p.A; j++) ent += __ldcg(p.log_std + j);
            s_entropy = ent;                                   // read by the owner of the loss element after barrier [B]
        }
        // ---- [B] every slab of minibatch s is written
        ++bar_gen;
        phase_grid_barrier(p.barrier, bar_gen * gridDim.x, p.spin_limit);
        if (p.dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(size_t)blockIdx.x * 16 + 5] = t; }
        // ---- [C] slice reduction + (cross-GPU sum) + Adam + image refresh
        phase_reduce_adam(p, s, nslabs, img, redw, s_entropy);
        if (p.dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(size_t)blockIdx.x * 16 + 6] = t; }
        if (s + 1 < p.n_steps) {
            // ---- [D] every slice of the new weights is in the global image; [E] re-stage it
            ++bar_gen;
            phase_grid_barrier(p.barrier, bar_gen * gridDim.x, p.spin_limit);
            if (gridDim.x > 1) {
                if (threadIdx.x == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");   // other SMs' generic-proxy stores -> this TMA (async proxy) read
                    mbar_expect_tx(mbar, (uint32_t)net.img_floats * 4u);
                    tma_bulk_g2s(img, p.image, (uint32_t)net.img_floats * 4u, mbar);
                }
                mbar_wait(mbar, img_parity);
                img_parity ^= 1;
            }
            if (p.dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); p.dbg[(size_t)blockIdx.x * 16 + 7] = t; }
        }
    }
}

// ===================================================================================================
// Shape-specialised persistent phase kernel for the Pendulum-class nets  S -> 64 -> 64 -> 1  (S <= 8, one hidden activation,
// single output: the value net and a one-dimensional action mean; BASELINE.json configs[1]).
//
// Same outer structure as fused_phase_kernel ([A] tiles, [B] barrier, [C] slice reduce + Adam, [D] barrier, [E] image re-stage)
// and the same global weight image / slab / optimiser layout, but [A] is written for ONE 128-row tile per CTA with every
// loop bound, stride and thread mapping a compile-time constant (the generic kernel spends 73 % of its issue slots on
// address arithmetic, predicates and branches of runtime-shaped loops; ncu, profiles/r02_*), and with 8 x 8 register tiles:
// measured on B200 (scripts/ubench/tile_loop.cu) an LDS.128 costs ~3.3 cycles of the SM's load/store unit inside such a loop,
// so the 8 x 4 tile (3 loads per 16 FFMA2) is LSU-bound at 76 % of the FFMA2 rate while 8 x 8 (4 per 32) reaches 85 %.  The
// 8 x 8 accumulators plus prefetched operands need ~200 registers, hence 256 threads (two warps per sub-partition) per CTA.
//   L0   (K = S)    256 threads x (4 rows x 8 units), weights warp-uniform
//   L1   (64 x 64)  8 rows x 8 units per thread, split-K halves on the two 128-thread groups; the epilogue also forms the
//                   partial dot products of the output layer (y = w2 . h2), so h2 is not re-read
//   head            128 threads (one per row): loss head on y, dLoss/dy -> gvec
//   dPre2           G2 = g w2 act'(h2), dW2, db1, db2: eight lanes per hidden unit
//   bwd l=1         warps 0-3: dX1 -> G1 (8 rows x 8 inputs per thread), then dW0 / db0 from G1 and the input tile, loss sums;
//                   warps 4-7: dW1 (8 x 8 outputs per thread over a 64-row half; the halves are added by all threads afterwards)
// One slab per CTA (the generic kernel writes one per 64-row tile slot).  Arithmetic per element is the generic kernel's
// (fma.rn chains in fixed order, packed two per FFMA2), so results agree with it to rounding of the changed summation splits.
// ===================================================================================================
constexpr int kS64TM = 128, kS64TMP = 132, kS64LDW = 68, kS64Threads = 256;

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void sts4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
template <int ID, int N> __device__ __forceinline__ void named_sync() { asm volatile("bar.sync %0, %1;" :: "n"(ID), "n"(N) : "memory"); }

// shared-memory floats after the weight image
constexpr int kS64X0 = 0;                                  // [8][TMP]   input tile (feature-major)
constexpr int kS64H1 = kS64X0 + 8 * kS64TMP;               // [64][TMP]  hidden 1
constexpr int kS64H2 = kS64H1 + 64 * kS64TMP;              // [64][TMP]  hidden 2
constexpr int kS64G2 = kS64H2 + 64 * kS64TMP;              // [64][TMP]  dLoss/d(pre-activation 2)
constexpr int kS64G1 = kS64G2 + 64 * kS64TMP;              // [64][TMP]  dLoss/d(pre-activation 1); forward: split-K exchange
constexpr int kS64CBLd = 72;                               // row stride of the dW1 partials: 8 * tj + tk hits 32 distinct banks
constexpr int kS64CB = kS64G1 + 64 * kS64TMP;              // [2][64][72] dW1 partials of the two 64-row halves; forward: y partials [4][128]
constexpr int kS64Red = kS64CB + 2 * 64 * kS64CBLd;        // [2][128]   per-row loss terms | per-row log_std gradient terms
constexpr int kS64Gv = kS64Red + 256;                      // [128]      dLoss/dy per row
constexpr int kS64Src = kS64Gv + 128;                      // [128] int  source rows of the tile
constexpr int kS64Bar = kS64Src + 128;                     // mbarrier (8 bytes)
constexpr int kS64Floats = kS64Bar + 4;

template <int ACT>
__global__ void __launch_bounds__(kS64Threads, 1) fused_phase_spec64_kernel(const PhaseArgs p) {
    constexpr int TM = kS64TM, TMP = kS64TMP, LDW = kS64LDW;
    extern __shared__ __align__(128) float smem[];
    __shared__ float redw[kS64Threads / 32][33];
    __shared__ float s_entropy;
    __shared__ float s_ls[4];         // per step: log_std, exp(log_std), exp(-2 log_std), -0.5 log(2 pi)   (policy mode)
    const FusedNet& net = p.net;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int S = net.sizes[0], SP = pad4(S);
    float* img = smem;
    float* base = smem + net.img_floats;
    float* X0 = base + kS64X0;
    float* H1 = base + kS64H1;
    float* H2 = base + kS64H2;
    float* G2 = base + kS64G2;
    float* G1 = base + kS64G1;
    float* CB = base + kS64CB;
    float* red = base + kS64Red;
    float* gvec = base + kS64Gv;
    int* src_rows = reinterpret_cast<int*>(base + kS64Src);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(base + kS64Bar);
    const float* W0 = img + net.wt_off[0];
    const float* W1 = img + net.wt_off[1];
    const float* W2 = img + net.wt_off[2];      // w2[k] = W2[k * 8]
    const float* B0 = img + net.bs_off[0];
    const float* B1 = img + net.bs_off[1];
    const float* B2 = img + net.bs_off[2];
    const int out_act = net.acts[2];

    const int n_tiles = (p.mb + TM - 1) / TM;
    const int stride = gridDim.x;
    const int t0 = blockIdx.x;
    const int rounds = t0 < n_tiles ? (n_tiles - t0 + stride - 1) / stride : 0;
    const int nslabs = min(n_tiles, stride);
    float* slab = p.partials + (size_t)t0 * p.slab;
    unsigned int bar_gen = 0;

    auto src_of = [&](int s, int tile, int r) -> int {      // src/trajectory_buffer.cu:208-209
        const int row = tile * TM + r;
        if (row >= p.mb) return -1;
        const int e = s / p.num_batches, k = s - e * p.num_batches;
        const int off = (k * p.batch_stride + p.row0 + row) % p.limit;
        return p.perms ? __ldg(p.perms + (size_t)e * p.limit + off) : off;
    };
    float h_target = 0.f, h_adv = 0.f, h_lp_old = 0.f, h_act0 = 0.f;     // per-row scalars of row t (threads < 128)
    int my_src = -1;
    auto issue_gather = [&]() {
        for (int e = t; e < TM * SP; e += kS64Threads) {
            const int r = e / SP, k = e - r * SP;
            const int src = src_rows[r];
            float* dst = X0 + k * TMP + r;
            if (src >= 0 && k < S) cp_async4(dst, p.state + (size_t)src * S + k);
            else *dst = 0.f;
        }
        if (t < TM) {
            my_src = src_rows[t];
            h_target = 0.f; h_adv = 0.f; h_lp_old = 0.f; h_act0 = 0.f;
            if (my_src >= 0) {
                if (p.mode == kFusedValue) {
                    h_target = __ldg(p.adv_target + my_src);
                } else {
                    h_adv = __ldg(p.advantage + my_src);
                    h_lp_old = __ldg(p.logprob + my_src);
                    h_act0 = __ldg(p.action + my_src);
                }
            }
        }
    };
    auto stamp = [&](int slot, int who = 0) {
        if (p.dbg && t == who) {
            unsigned long long tm;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
            p.dbg[(size_t)blockIdx.x * 16 + slot] = tm;
        }
    };

    if (t == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    int nxt_src = -1;
    if (rounds > 0) {
        if (t < TM) src_rows[t] = src_of(0, t0, t);
        __syncthreads();
        issue_gather();
    }
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(mbar, (uint32_t)net.img_floats * 4u);
        tma_bulk_g2s(img, p.image, (uint32_t)net.img_floats * 4u, mbar);
    }
    uint32_t img_parity = 0;
    mbar_wait(mbar, img_parity);
    img_parity ^= 1;

    for (int s = 0; s < p.n_steps; s++) {
        if (t == 64 && p.mode == kFusedPolicy) {
            // log_std is final for this step (last written in [C] of step s-1, before [D]): ONE L2 read per CTA instead of one per
            // head warp (592 simultaneous requests for one sector cost the heads ~2 us), and the row-independent terms of
            // src/policy.cu:67-74 / :91-111 evaluated once
            const float ls = __ldcg(p.log_std);
            s_ls[0] = ls;
            s_ls[1] = expf(ls);
            s_ls[2] = expf(-2.f * ls);
            s_ls[3] = (float)(-0.5 * (double)logf((float)(2 * kPiF)));
        }
        for (int rd = 0; rd < rounds; rd++) {
            const bool accum = rd > 0;
            auto put = [&](int idx, float v) { slab[idx] = accum ? slab[idx] + v : v; };
            cp_async_wait_all();
            __syncthreads();                                   // the gathered tile is complete and visible
            int ns = s, nt = t0 + (rd + 1) * stride;
            if (rd + 1 >= rounds) { ns = s + 1; nt = t0; }
            const bool have_next = ns < p.n_steps;
            if (have_next && t < TM) nxt_src = src_of(ns, nt, t);
            stamp(3);
            // ---- L0: H1[j][r] = act(sum_k X0[k][r] W0[k][j] + b0[j]);  rows 4*lane.., units 8*warp..
            {
                float2 acc[2][8];
#pragma unroll
                for (int c = 0; c < 8; c++) { acc[0][c] = make_float2(0.f, 0.f); acc[1][c] = make_float2(0.f, 0.f); }
                for (int k = 0; k < S; k++) {
                    const float4 x = lds4(X0 + k * TMP + 4 * lane);
                    const float4 w0 = lds4(W0 + k * LDW + 8 * warp), w1 = lds4(W0 + k * LDW + 8 * warp + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        acc[0][c] = ffma2(make_float2(x.x, x.y), bcast2(wv[c]), acc[0][c]);
                        acc[1][c] = ffma2(make_float2(x.z, x.w), bcast2(wv[c]), acc[1][c]);
                    }
                }
                const float4 b0 = lds4(B0 + 8 * warp), b1 = lds4(B0 + 8 * warp + 4);
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int c = 0; c < 8; c++)
                    sts4(H1 + (8 * warp + c) * TMP + 4 * lane, act_apply(acc[0][c].x + bv[c], ACT), act_apply(acc[0][c].y + bv[c], ACT),
                         act_apply(acc[1][c].x + bv[c], ACT), act_apply(acc[1][c].y + bv[c], ACT));
            }
            __syncthreads();
            stamp(8);
            // ---- L1: 8 rows x 8 units per thread; split-K halves: group g = warp >> 2 accumulates k in [32g, 32g + 32);
            // rows {4tr..} U {64+4tr..}, units {4tc..} U {32+4tc..}
            {
                const int g = t >> 7, lt = t & 127, tr = lt & 15, tc = lt >> 4;
                float2 acc[4][8];                   // row pairs (4tr, +1), (4tr+2, +3), (64+4tr, +1), (64+4tr+2, +3)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) acc[r][c] = make_float2(0.f, 0.f);
                const float* xp = H1 + 32 * g * TMP + 4 * tr;
                const float* wp = W1 + 32 * g * LDW + 4 * tc;
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const float4 c0 = lds4(xp + k * TMP), c1 = lds4(xp + k * TMP + 64);
                    const float4 w0 = lds4(wp + k * LDW), w1 = lds4(wp + k * LDW + 32);
                    const float2 ap[4] = {make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), make_float2(c1.z, c1.w)};
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[r][c] = ffma2(ap[r], bcast2(wv[c]), acc[r][c]);
                }
                // group 0 finalises rows 4tr.. (pairs 0,1), group 1 rows 64+4tr.. (pairs 2,3): pass the other quad through G1
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int col = (c < 4) ? 4 * tc + c : 28 + 4 * tc + c;
                    float* dst = G1 + col * TMP + 4 * tr + (g ? 0 : 64);
                    if (g) sts4(dst, acc[0][c].x, acc[0][c].y, acc[1][c].x, acc[1][c].y);
                    else sts4(dst, acc[2][c].x, acc[2][c].y, acc[3][c].x, acc[3][c].y);
                }
                __syncthreads();
                const float4 b0 = lds4(B1 + 4 * tc), b1 = lds4(B1 + 32 + 4 * tc);
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float yp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int col = (c < 4) ? 4 * tc + c : 28 + 4 * tc + c;
                    const float2 lo = g ? acc[2][c] : acc[0][c], hi = g ? acc[3][c] : acc[1][c];
                    const float4 q = lds4(G1 + col * TMP + 4 * tr + 64 * g);
                    float o[4] = {lo.x, lo.y, hi.x, hi.y};
                    const float qv[4] = {q.x, q.y, q.z, q.w};
                    const float w2 = W2[col * 8];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        o[r] = g ? qv[r] + o[r] : o[r] + qv[r];            // lower-k partial + upper-k partial
                        o[r] = act_apply(o[r] + bv[c], ACT);
                        yp[r] = fmaf(o[r], w2, yp[r]);
                    }
                    sts4(H2 + col * TMP + 4 * tr + 64 * g, o[0], o[1], o[2], o[3]);
                }
                // y partials: the two unit groups of a warp meet through one shuffle; the 4 warps of a group through CB[4][128]
#pragma unroll
                for (int r = 0; r < 4; r++) yp[r] += __shfl_xor_sync(kFull, yp[r], 16);
                if (lane < 16) sts4(CB + ((warp & 3) * TM) + 64 * g + 4 * tr, yp[0], yp[1], yp[2], yp[3]);
            }
            __syncthreads();
            stamp(9);
            // ---- head: one thread per row (src/loss.cu:5-23 | src/policy.cu:67-111 + src/ppo.cu:89-98)
            if (t < TM) {
                const float part = (CB[t] + CB[TM + t]) + (CB[2 * TM + t] + CB[3 * TM + t]);
                const float y = act_apply(part + B2[0], out_act);
                float loss_term = 0.f, gout = 0.f, gls = 0.f;
                if (my_src >= 0) {
                    if (p.mode == kFusedValue) {
                        gout = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(y, h_target)), (float)p.m_total);
                        const float d = __fsub_rn(h_target, y);
                        loss_term = __fmul_rn(d, d);
                    } else {
                        const float ls = s_ls[0], e2 = s_ls[2];
                        const float z = __fdiv_rn(__fsub_rn(h_act0, y), s_ls[1]);           // fused_log_prob with A = 1
                        const float lp = (float)((double)s_ls[3] - ((double)ls + 0.5 * (double)__fmul_rn(z, z)));
                        const float ratio = expf(__fsub_rn(lp, h_lp_old));
                        const bool adv_pos = h_adv > 0.f;
                        const bool hi = ratio > 1.f + p.epsilon, lo = ratio < 1.f - p.epsilon;
                        const float sel = adv_pos ? (hi ? 1.f + p.epsilon : ratio) : (lo ? 1.f - p.epsilon : ratio);
                        loss_term = __fmul_rn(h_adv, sel);
                        const int keep = adv_pos ? !hi : !lo;
                        const float gg = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), h_adv), ratio), (float)p.m_total);
                        const float diff = __fsub_rn(h_act0, y);
                        gout = __fmul_rn(__fmul_rn(diff, e2), gg);
                        gls = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), gg);
                    }
                }
                gvec[t] = act_grad(y, gout, out_act);
                red[t] = loss_term;                            // summed (fixed order) in the tail of the dX warps, off this chain
                red[128 + t] = gls;
            }
            __syncthreads();
            stamp(10);
            // ---- dPre2: G2[k][r] = g_r w2[k] act'(H2[k][r]); dW2[k] = sum_r g_r H2[k][r]; db1[k] = sum_r G2[k][r]; db2 = sum_r g_r.
            // Eight lanes per hidden unit (k = (t >> 3) and 32 + (t >> 3)); lane l8 takes the row quads l8, l8 + 8, l8 + 16, l8 + 24
            // (a quarter-warp reads 128 contiguous bytes); the eight partial sums of a unit meet through three shuffles.
            {
                const int l8 = t & 7;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int k = (t >> 3) + 32 * h;
                    const float w2 = W2[k * 8];
                    float pw = 0.f, pb = 0.f, pg = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int r0 = 4 * (l8 + 8 * i);
                        const float4 hv = lds4(H2 + k * TMP + r0);
                        const float4 g4 = lds4(gvec + r0);
                        const float d0 = act_grad(hv.x, __fmul_rn(g4.x, w2), ACT), d1 = act_grad(hv.y, __fmul_rn(g4.y, w2), ACT);
                        const float d2 = act_grad(hv.z, __fmul_rn(g4.z, w2), ACT), d3 = act_grad(hv.w, __fmul_rn(g4.w, w2), ACT);
                        sts4(G2 + k * TMP + r0, d0, d1, d2, d3);
                        pw = fmaf(g4.w, hv.w, fmaf(g4.z, hv.z, fmaf(g4.y, hv.y, fmaf(g4.x, hv.x, pw))));
                        pb += (d0 + d1) + (d2 + d3);
                        pg += (g4.x + g4.y) + (g4.z + g4.w);
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        pw += __shfl_xor_sync(kFull, pw, o);
                        pb += __shfl_xor_sync(kFull, pb, o);
                        pg += __shfl_xor_sync(kFull, pg, o);
                    }
                    if (l8 == 0) {
                        put(net.w_off[2] + k, pw);
                        put(net.b_off[1] + k, pb);
                        if (k == 0) put(net.b_off[2], pg);
                    }
                }
            }
            __syncthreads();
            stamp(11);
            // ---- backward of layer 1 (and layer 0 behind it)
            if (t < 128) {
                // dX1: G1[k][r] = (sum_j G2[j][r] W1[k][j]) act'(H1[k][r]);  k = tc + 8c, rows 4tr.. and 64+4tr..
                const int tr = t & 15, tc = t >> 4;
                {
                    float2 acc[4][8];
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 8; c++) acc[r][c] = make_float2(0.f, 0.f);
                    const float* wrow = W1 + tc * LDW;
                    const float* gp = G2 + 4 * tr;
#pragma unroll 4
                    for (int j0 = 0; j0 < 64; j0 += 2) {
                        float2 wv[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) wv[c] = lds2(wrow + 8 * c * LDW + j0);
                        {
                            const float4 g0 = lds4(gp + j0 * TMP), g1 = lds4(gp + j0 * TMP + 64);
                            const float2 gq[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
                            for (int r = 0; r < 4; r++)
#pragma unroll
                                for (int c = 0; c < 8; c++) acc[r][c] = ffma2(gq[r], bcast2(wv[c].x), acc[r][c]);
                        }
                        {
                            const float4 g0 = lds4(gp + (j0 + 1) * TMP), g1 = lds4(gp + (j0 + 1) * TMP + 64);
                            const float2 gq[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
#pragma unroll
                            for (int r = 0; r < 4; r++)
#pragma unroll
                                for (int c = 0; c < 8; c++) acc[r][c] = ffma2(gq[r], bcast2(wv[c].y), acc[r][c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int k = tc + 8 * c;
                        const float4 h0 = lds4(H1 + k * TMP + 4 * tr), h1 = lds4(H1 + k * TMP + 4 * tr + 64);
                        sts4(G1 + k * TMP + 4 * tr, act_grad(h0.x, acc[0][c].x, ACT), act_grad(h0.y, acc[0][c].y, ACT),
                             act_grad(h0.z, acc[1][c].x, ACT), act_grad(h0.w, acc[1][c].y, ACT));
                        sts4(G1 + k * TMP + 4 * tr + 64, act_grad(h1.x, acc[2][c].x, ACT), act_grad(h1.y, acc[2][c].y, ACT),
                             act_grad(h1.z, acc[3][c].x, ACT), act_grad(h1.w, acc[3][c].y, ACT));
                    }
                }
                stamp(14);
                named_sync<1, 128>();
                // dW0[j][k] = sum_r G1[j][r] X0[k][r], db0[j] = sum_r G1[j][r]: j = 16*warp + (lane & 15), row half = lane >> 4
                {
                    const int j = 16 * warp + (lane & 15), hh = lane >> 4;
                    float accw[8], accb = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; k++) accw[k] = 0.f;
                    const float* gq = G1 + j * TMP + 64 * hh;
                    const float* xq = X0 + 64 * hh;
#pragma unroll 4
                    for (int i = 0; i < 16; i++) {
                        const float4 gv = lds4(gq + 4 * i);
                        accb += (gv.x + gv.y) + (gv.z + gv.w);
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (k < S) {
                                const float4 xv = lds4(xq + k * TMP + 4 * i);
                                accw[k] = fmaf(gv.w, xv.w, fmaf(gv.z, xv.z, fmaf(gv.y, xv.y, fmaf(gv.x, xv.x, accw[k]))));
                            }
                    }
                    accb += __shfl_xor_sync(kFull, accb, 16);
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (k < S) accw[k] += __shfl_xor_sync(kFull, accw[k], 16);
                    if (hh == 0) {
                        put(net.b_off[0] + j, accb);
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (k < S) put(net.w_off[0] + j * S + k, accw[k]);
                    }
                }
                // tile sums of the loss terms (warp 0) and of the log_std gradient terms (warp 1): fixed order
                if (warp < 2 && (warp == 0 || p.mode == kFusedPolicy)) {
                    const float4 v = lds4(red + 128 * warp + 4 * lane);
                    const float sv = warp_sum((v.x + v.y) + (v.z + v.w));
                    if (lane == 0) put(warp == 0 ? net.P + 1 : net.P, sv);
                }
                stamp(15);
            } else {
                // dW1[j][k] = sum_r G2[j][r] H1[k][r]: j = tj + 8a, k = tk + 8b, rows [64*half, 64*half + 64); two rows per step
                const int lt = t - 128, half = lt >> 6, l6 = lt & 63, tk = l6 & 7, tj = l6 >> 3;
                float2 acc[8][8];                   // (sum over even rows, sum over odd rows)
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int b = 0; b < 8; b++) acc[a][b] = make_float2(0.f, 0.f);
                const float* gb = G2 + tj * TMP + 64 * half;
                const float* xb = H1 + tk * TMP + 64 * half;
#pragma unroll 2
                for (int r = 0; r < 64; r += 2) {
                    float2 g[8], x[8];
#pragma unroll
                    for (int a = 0; a < 8; a++) g[a] = lds2(gb + 8 * a * TMP + r);
#pragma unroll
                    for (int b = 0; b < 8; b++) x[b] = lds2(xb + 8 * b * TMP + r);
#pragma unroll
                    for (int a = 0; a < 8; a++)
#pragma unroll
                        for (int b = 0; b < 8; b++) acc[a][b] = ffma2(g[a], x[b], acc[a][b]);
                }
                stamp(1, 128);
                // both halves park their partial sums; all threads add and store them (coalesced) after the barrier
                float* cb = CB + half * 64 * kS64CBLd + tj * kS64CBLd + tk;
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int b = 0; b < 8; b++) cb[8 * a * kS64CBLd + 8 * b] = acc[a][b].x + acc[a][b].y;
            }
            __syncthreads();                                   // X0 / src_rows are about to be refilled; the dW1 partials are complete
            stamp(13);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int e = t + kS64Threads * i, j = e >> 6, k = e & 63;
                put(net.w_off[1] + e, CB[j * kS64CBLd + k] + CB[64 * kS64CBLd + j * kS64CBLd + k]);
            }
            if (have_next) {
                if (t < TM) src_rows[t] = nxt_src;
                __syncthreads();
                issue_gather();
            }
            stamp(12);
        }
        stamp(4);
        if (t == 33 && p.mode == kFusedPolicy) {               // log_std is stable here: last written in [C] of step s-1, before [D]
            float ent = (float)(p.A * 0.5 * (1 + log(2 * kPiF)));             // src/policy.cu:171-178
            for (int j = 0; j < p.A; j++) ent += __ldcg(p.log_std + j);
            s_entropy = ent;
        }
        // ---- [B] every slab of minibatch s is written
        ++bar_gen;
        phase_grid_barrier(p.barrier, bar_gen * gridDim.x, p.spin_limit);
        stamp(5);
        // ---- [C] slice reduction + (cross-GPU sum) + Adam + image refresh
        phase_reduce_adam(p, s, nslabs, img, redw, s_entropy);
        stamp(6);
        if (s + 1 < p.n_steps) {
            // ---- [D] every slice of the new weights is in the global image; [E] re-stage it
            ++bar_gen;
            phase_grid_barrier(p.barrier, bar_gen * gridDim.x, p.spin_limit);
            if (gridDim.x > 1) {
                if (t == 0) {
                    asm volatile("fence.proxy.async;" ::: "memory");
                    mbar_expect_tx(mbar, (uint32_t)net.img_floats * 4u);
                    tma_bulk_g2s(img, p.image, (uint32_t)net.img_floats * 4u, mbar);
                }
                mbar_wait(mbar, img_parity);
                img_parity ^= 1;
            }
            stamp(7);
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
    // no dependent-launch overlap while per-kernel event pairs are being recorded: overlapping pairs would not add up to the step
    const bool pdl = pdl_enabled() && !g_profiling && pl.kind == 1 && !reduced_out;
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

// ---- persistent phase launch ------------------------------------------------------------------------
static bool phase_enabled() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("PPO_B200_PERSISTENT"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}

struct PhaseGeom { bool ok; int nsub, grid; size_t smem; int sub_floats; };
static PhaseGeom phase_geometry(const FusedPlan& pl, int mb) {
    PhaseGeom g{};
    if (!pl.ok) return g;
    // one tile slot: [activations | E0 E1 | reduction scratch] = everything of the plan after the image
    const int sub_floats = ((int)(pl.smem_bytes / sizeof(float)) - pl.net.img_floats + 3) & ~3;
    const int n_tiles = div_up(mb, kT64TM);
    const size_t limit = 226 * 1024;
    auto bytes = [&](int nsub) { return (size_t)(pl.net.img_floats + nsub * sub_floats + 4) * sizeof(float); };
    int nsub = (n_tiles > num_sms() && bytes(2) <= limit) ? 2 : 1;
    if (bytes(nsub) > limit) return g;
    g.ok = true;
    g.nsub = nsub;
    g.sub_floats = sub_floats;
    g.smem = bytes(nsub);
    g.grid = std::max(1, std::min(num_sms(), div_up(n_tiles, nsub)));
    return g;
}

bool fused_phase_supported(NeuralNetwork* nn) {
    if (!phase_enabled()) return false;
    const FusedPlan pl = choose_plan(net_dev(nn));
    return pl.ok && phase_geometry(pl, 64).ok;
}

// The shape-specialised kernel applies to  S<=8 -> 64 -> 64 -> 1  nets with one hidden activation (tanh / relu) when the minibatch
// gives every SM at least one 128-row tile (smaller minibatches keep the generic kernel's 64-row tiles: more CTAs busy).
// PPO_B200_SPEC64=0 disables it (A/B runs, parity tests of the generic kernel).
static int spec64_kind(const FusedPlan& pl, int mb) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("PPO_B200_SPEC64"); enabled = (e && e[0] == '0') ? 0 : 1; }
    const FusedNet& n = pl.net;
    if (!enabled || !pl.ok || n.L != 3) return 0;
    if (n.sizes[0] > 8 || n.sizes[1] != 64 || n.sizes[2] != 64 || n.sizes[3] != 1) return 0;
    if (n.acts[0] != n.acts[1] || (n.acts[0] != kActTanh && n.acts[0] != kActRelu)) return 0;
    if (n.ldw[0] != kS64LDW || n.ldw[1] != kS64LDW || n.ldw[2] != 8) return 0;
    // the 128-row tiles pay off when they fill the machine; between one tile and a full wave of 64-row tiles the generic kernel
    // has more CTAs to spread over.  A minibatch of <= 64 rows (the reference's default, src/main.c:36) is one CTA either way,
    // and the specialised code is the faster single tile.
    if (mb > kT64TM && div_up(mb, kT64TM) <= num_sms()) return 0;
    if ((size_t)(n.img_floats + kS64Floats) * sizeof(float) > 220 * 1024) return 0;
    return n.acts[0];
}

static unsigned int* g_phase_barrier = nullptr;
static float4* g_phase_coef = nullptr;
static int g_phase_coef_cap = 0;

void fused_phase_update(NeuralNetwork* nn, GaussianPolicy* policy, Adam* adam_net, Adam* adam_ls, float lr, const int* perms,
                        int limit, int mb, int m_total, int row0, int batch_stride, int num_batches, int n_epochs,
                        const TrajectoryBuffer* b, float epsilon, float ent_coeff, float* loss_slot, bool dp_peer) {
    NetDev* nd = net_dev(nn);
    const FusedPlan pl = choose_plan(nd);
    const PhaseGeom geo = phase_geometry(pl, mb);
    if (!geo.ok) B200_FATAL("fused_phase_update on an unsupported net");
    const int n_steps = n_epochs * num_batches;
    if (n_steps <= 0 || mb <= 0) return;
    const int A = policy ? policy->action_size : 1;
    const int slab = (int)nd->param_count + A + 1;
    const int spec = spec64_kind(pl, mb);             // 0: generic kernel; else the hidden activation of the specialised one
    const int spec_grid = std::max(1, std::min(num_sms(), div_up(mb, kS64TM)));
    const int nslabs = spec ? spec_grid : std::min(div_up(mb, kT64TM), geo.grid * geo.nsub);
    const size_t need = (size_t)nslabs * slab;
    if (need > nd->partials_cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (nd->partials) CUDA_CHECK(cudaFree(nd->partials));
        nd->partials = dmalloc<float>(need);
        nd->partials_cap = need;
    }
    if (!g_phase_barrier) g_phase_barrier = dmalloc<unsigned int>(32);
    if (n_steps > g_phase_coef_cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (g_phase_coef) CUDA_CHECK(cudaFree(g_phase_coef));
        g_phase_coef_cap = n_steps + 64;
        g_phase_coef = dmalloc<float4>(g_phase_coef_cap);
    }
    // per-step bias corrections, host powf exactly as src/adam.cu:56-59 evaluates them
    std::vector<float4> coef(n_steps);
    for (int s = 0; s < n_steps; s++) {
        const int tn = adam_net->time_step + s + 1;
        coef[s].x = lr / (1 - powf(adam_net->beta1, tn));
        coef[s].y = 1 - powf(adam_net->beta2, tn);
        coef[s].z = coef[s].w = 0.f;
        if (policy) {
            const int tl = adam_ls->time_step + s + 1;
            coef[s].z = lr / (1 - powf(adam_ls->beta1, tl));
            coef[s].w = 1 - powf(adam_ls->beta2, tl);
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(g_phase_coef, coef.data(), (size_t)n_steps * sizeof(float4), cudaMemcpyHostToDevice, stream()));
    CUDA_CHECK(cudaMemsetAsync(g_phase_barrier, 0, sizeof(unsigned int), stream()));

    PhaseArgs a{};
    a.net = pl.net;
    a.image = ensure_image(nd, pl);
    a.perms = perms; a.limit = limit; a.mb = mb; a.m_total = m_total; a.row0 = row0; a.batch_stride = batch_stride;
    a.num_batches = num_batches; a.n_steps = n_steps;
    a.mode = policy ? kFusedPolicy : kFusedValue;
    a.state = b->d_state_p; a.action = b->d_action_p; a.logprob = b->d_logprob_p;
    a.advantage = b->d_advantage_p; a.adv_target = b->d_adv_target_p;
    a.log_std = policy ? policy->d_log_std : nullptr;
    a.epsilon = epsilon; a.ent_coeff = ent_coeff;
    a.partials = nd->partials; a.slab = slab; a.P = (int)nd->param_count; a.A = A;
    a.sub_floats = geo.sub_floats;
    a.g_off = pl.g_off - pl.net.img_floats;
    a.red_off = pl.red_off - pl.net.img_floats;
    a.netseg = make_seg(nd->params, nd->grads, adam_net, lr);
    if (policy) a.ls = make_seg(policy->d_log_std, policy->d_log_std_grad, adam_ls, lr);
    a.coef = g_phase_coef;
    a.loss_slot = loss_slot;
    a.barrier = g_phase_barrier;
    a.spin_limit = dist_spin_limit();
    a.dbg = phase_dbg();
    if (dp_peer) {
        a.peer = dist_peer_reserve((size_t)slab, n_steps);
        if (!a.peer.ready) B200_FATAL("peer exchange requested but the peer arena is not available (slab %d floats)", slab);
    }
    if (spec) {
        const size_t smem = (size_t)(pl.net.img_floats + kS64Floats) * sizeof(float);
        static size_t configured64 = 0;
        if (smem > configured64) {
            CUDA_CHECK(cudaFuncSetAttribute(fused_phase_spec64_kernel<kActTanh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_CHECK(cudaFuncSetAttribute(fused_phase_spec64_kernel<kActRelu>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured64 = smem;
        }
        if (spec == kActTanh) B200_LAUNCH_COOP(fused_phase_spec64_kernel<kActTanh>, spec_grid, kS64Threads, smem, &a);
        else B200_LAUNCH_COOP(fused_phase_spec64_kernel<kActRelu>, spec_grid, kS64Threads, smem, &a);
    } else {
        static size_t configured = 0;
        if (geo.smem > configured) {
            CUDA_CHECK(cudaFuncSetAttribute(fused_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem));
            configured = geo.smem;
        }
        B200_LAUNCH_COOP(fused_phase_kernel, geo.grid, 256 * geo.nsub, geo.smem, &a);
    }
    adam_net->time_step += n_steps;
    if (policy) adam_ls->time_step += n_steps;
    nd->last_splits = nslabs;
}

}  // namespace b200

// debug: copy the phase timestamps of the last fused_tile64_kernel launch ([blocks][16] u64) to the host
extern "C" void ppo_b200_debug_phase_stamps(unsigned long long* out, int blocks) {
    if (!b200::g_phase_dbg) return;
    CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    CUDA_CHECK(cudaMemcpy(out, b200::g_phase_dbg, (size_t)blocks * 16 * 8, cudaMemcpyDeviceToHost));
}
