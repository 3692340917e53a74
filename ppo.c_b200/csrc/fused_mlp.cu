// fused_mlp.cu — the small-net fast path (SURVEY.md §8 rows a7+a9..a14 in two launches per minibatch).
//
// For the reference-width actor/critic nets (every layer width <= 128: 2x64, 2x128) one minibatch
// update is
//   fused_update_kernel   gather rows by permutation index -> forward through ALL layers -> fused loss
//                         head (MSE, or Gaussian log-prob + PPO-clip surrogate) -> backward through
//                         all layers -> this CTA's partial gradient slab.  Weights are staged ONCE per
//                         CTA in shared memory (transposed, k-major), activations never leave shared
//                         memory, nothing but the slab is written to HBM.
//   fused_reduce_adam_kernel  fixed-order sum of the slabs + Adam on the flat parameter vector (+ the
//                         log_std vector for the policy) + loss accumulation.
// versus ~14 launches through the layer-wise kernels of gemm.cu/policy.cu/adam.cu (which remain the
// generic path for wider nets).  The reference does this with ~25 launches, 1-3 blocking D2H reads
// and 1-2 cudaMallocs per minibatch (src/ppo.cu:495-532).
//
// Shared-memory layouts (TM = rows per CTA, TMP = TM + 4 so that TMP % 32 == 4):
//   activations / gradients  At[feature][TMP]   feature-major: a thread reads 4 consecutive ROWS with
//                            one conflict-free LDS.128
//   weights                  Wt[in][out_pad]    k-major transpose of the reference's W[out][in]
// Thread mappings (256 threads):
//   forward / dX : thread = (row lane tr, column group tc): rows {4tr..4tr+3, TM/2+4tr..+3} x 4 columns
//   dW           : thread = (tj, tk) in 16x16, owns j = tj+16*jj, k = tk+16*kk (interleaved so the 16
//                  lanes of a half-warp read 16 consecutive feature rows: stride TMP -> conflict-free)
// All arithmetic is fp32 FFMA (tolerance 1e-5).
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kFusedMaxLayers = 5;     // weight layers
constexpr int kFusedThreads = 256;
constexpr double kPiF = 3.14159265358979323846;

struct FusedNet {
    int L;                              // weight layers
    int sizes[kFusedMaxLayers + 1];
    int acts[kFusedMaxLayers];
    int w_off[kFusedMaxLayers], b_off[kFusedMaxLayers];   // offsets in the flat parameter vector
    int wt_off[kFusedMaxLayers];        // offsets of Wt[in][out_pad] in shared memory (floats)
    int a_off[kFusedMaxLayers + 1];     // offsets of At buffers in shared memory (floats)
    int bs_off[kFusedMaxLayers];        // offsets of the layer biases inside the shared bias region
    int P;
    int max_width_pad;
};

enum FusedMode { kFusedForward = 0, kFusedValue = 1, kFusedPolicy = 2 };

struct FusedArgs {
    FusedNet net;
    const float* params;
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
    int smem_b_off;            // biases
    int smem_red_off;
};

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }

template <int TM>
__device__ __forceinline__ void fused_forward_layer(const float* __restrict__ Xt, const float* __restrict__ Wt,
                                                    const float* __restrict__ bias, float* __restrict__ Yt,
                                                    int n_in, int n_out, int act) {
    constexpr int TMP = TM + 4, RL = TM / 8;
    const int tr = threadIdx.x % RL, tc = threadIdx.x / RL;
    const int out_pad = pad4(n_out);
    if (4 * tc >= out_pad) return;
    float acc[8][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float b = (4 * tc + c < n_out) ? bias[4 * tc + c] : 0.f;
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r][c] = b;
    }
    const float* xp = Xt + 4 * tr;
    const float* wp = Wt + 4 * tc;
#pragma unroll 4
    for (int k = 0; k < n_in; k++) {
        const float4 a0 = *reinterpret_cast<const float4*>(xp + k * TMP);
        const float4 a1 = *reinterpret_cast<const float4*>(xp + k * TMP + TM / 2);
        const float4 w = *reinterpret_cast<const float4*>(wp + k * out_pad);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = fmaf(a[r], wv[c], acc[r][c]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float* yp = Yt + (4 * tc + c) * TMP + 4 * tr;
        *reinterpret_cast<float4*>(yp) = make_float4(act_apply(acc[0][c], act), act_apply(acc[1][c], act),
                                                     act_apply(acc[2][c], act), act_apply(acc[3][c], act));
        *reinterpret_cast<float4*>(yp + TM / 2) = make_float4(act_apply(acc[4][c], act), act_apply(acc[5][c], act),
                                                              act_apply(acc[6][c], act), act_apply(acc[7][c], act));
    }
}

// GXt[k][r] = (sum_j Gt[j][r] * W[j][k]) * act'(Ht[k][r])   (Wt is [k][out_pad])
template <int TM>
__device__ __forceinline__ void fused_backward_input(const float* __restrict__ Gt, const float* __restrict__ Wt,
                                                     const float* __restrict__ Ht, float* __restrict__ GXt,
                                                     int n_in, int n_out, int act_prev) {
    constexpr int TMP = TM + 4, RL = TM / 8;
    const int tr = threadIdx.x % RL, tc = threadIdx.x / RL;
    const int out_pad = pad4(n_out);
    if (4 * tc >= pad4(n_in)) return;
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.f;
    const float* gp = Gt + 4 * tr;
    const float* wp = Wt + (4 * tc) * out_pad;
    bool kok[4];
#pragma unroll
    for (int c = 0; c < 4; c++) kok[c] = 4 * tc + c < n_in;
#pragma unroll 4
    for (int j = 0; j < n_out; j++) {
        const float4 g0 = *reinterpret_cast<const float4*>(gp + j * TMP);
        const float4 g1 = *reinterpret_cast<const float4*>(gp + j * TMP + TM / 2);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float wv[4];
#pragma unroll
        for (int c = 0; c < 4; c++) wv[c] = kok[c] ? wp[c * out_pad + j] : 0.f;
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = fmaf(g[r], wv[c], acc[r][c]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        if (!kok[c]) continue;
        const float* hp = Ht + (4 * tc + c) * TMP + 4 * tr;
        float* op = GXt + (4 * tc + c) * TMP + 4 * tr;
        const float4 h0 = *reinterpret_cast<const float4*>(hp), h1 = *reinterpret_cast<const float4*>(hp + TM / 2);
        *reinterpret_cast<float4*>(op) = make_float4(act_grad(h0.x, acc[0][c], act_prev), act_grad(h0.y, acc[1][c], act_prev),
                                                     act_grad(h0.z, acc[2][c], act_prev), act_grad(h0.w, acc[3][c], act_prev));
        *reinterpret_cast<float4*>(op + TM / 2) = make_float4(act_grad(h1.x, acc[4][c], act_prev), act_grad(h1.y, acc[5][c], act_prev),
                                                              act_grad(h1.z, acc[6][c], act_prev), act_grad(h1.w, acc[7][c], act_prev));
    }
}

// gW[j][k] = sum_r Gt[j][r] * Xt[k][r]  -> global slab (row-major [out][in], the reference layout)
template <int TM, int JJ, int KK>
__device__ __forceinline__ void fused_backward_weights(const float* __restrict__ Gt, const float* __restrict__ Xt,
                                                       float* __restrict__ gW, int n_in, int n_out) {
    constexpr int TMP = TM + 4;
    const int tk = threadIdx.x & 15, tj = threadIdx.x >> 4;
    float acc[JJ][KK];
#pragma unroll
    for (int a = 0; a < JJ; a++)
#pragma unroll
        for (int b = 0; b < KK; b++) acc[a][b] = 0.f;
    // clamp out-of-range rows to a valid one (results discarded) so loads stay in bounds
    int jrow[JJ], krow[KK];
#pragma unroll
    for (int a = 0; a < JJ; a++) jrow[a] = min(tj + 16 * a, n_out - 1) * TMP;
#pragma unroll
    for (int b = 0; b < KK; b++) krow[b] = min(tk + 16 * b, n_in - 1) * TMP;
#pragma unroll 2
    for (int r = 0; r < TM; r += 4) {
        float4 g[JJ], x[KK];
#pragma unroll
        for (int a = 0; a < JJ; a++) g[a] = *reinterpret_cast<const float4*>(Gt + jrow[a] + r);
#pragma unroll
        for (int b = 0; b < KK; b++) x[b] = *reinterpret_cast<const float4*>(Xt + krow[b] + r);
#pragma unroll
        for (int a = 0; a < JJ; a++)
#pragma unroll
            for (int b = 0; b < KK; b++) {
                acc[a][b] = fmaf(g[a].x, x[b].x, acc[a][b]);
                acc[a][b] = fmaf(g[a].y, x[b].y, acc[a][b]);
                acc[a][b] = fmaf(g[a].z, x[b].z, acc[a][b]);
                acc[a][b] = fmaf(g[a].w, x[b].w, acc[a][b]);
            }
    }
#pragma unroll
    for (int a = 0; a < JJ; a++) {
        const int j = tj + 16 * a;
        if (j >= n_out) continue;
#pragma unroll
        for (int b = 0; b < KK; b++) {
            const int k = tk + 16 * b;
            if (k < n_in) gW[(size_t)j * n_in + k] = acc[a][b];
        }
    }
}

template <int TM>
__device__ __forceinline__ void fused_weights_dispatch(const float* Gt, const float* Xt, float* gW, int n_in, int n_out) {
    const int jj = (n_out + 15) / 16, kk = (n_in + 15) / 16;
#define B200_DW(J, K) fused_backward_weights<TM, J, K>(Gt, Xt, gW, n_in, n_out)
    if (jj <= 1) { if (kk <= 1) B200_DW(1, 1); else if (kk <= 4) B200_DW(1, 4); else B200_DW(1, 8); }
    else if (jj <= 4) { if (kk <= 1) B200_DW(4, 1); else if (kk <= 2) B200_DW(4, 2); else if (kk <= 4) B200_DW(4, 4); else B200_DW(4, 8); }
    else { if (kk <= 1) B200_DW(8, 1); else if (kk <= 2) B200_DW(8, 2); else if (kk <= 4) B200_DW(8, 4); else B200_DW(8, 8); }
#undef B200_DW
}

// gb[j] = sum_r Gt[j][r] : one warp per feature row
template <int TM>
__device__ __forceinline__ void fused_bias_grad(const float* __restrict__ Gt, float* __restrict__ gb, int n_out) {
    constexpr int TMP = TM + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < n_out; j += kFusedThreads / 32) {
        float s = 0.f;
        for (int r = lane; r < TM; r += 32) s += Gt[j * TMP + r];
        s = warp_sum(s);
        if (lane == 0) gb[j] = s;
    }
}

__device__ __forceinline__ float fused_log_prob(const float* mu, const float* log_std, const float* action, int A) {
    float logprob = (float)(-0.5 * A * (double)logf((float)(2 * kPiF)));   // src/policy.cu:67-74
    for (int j = 0; j < A; j++) {
        const float z = __fdiv_rn(__fsub_rn(action[j], mu[j]), expf(log_std[j]));
        logprob = (float)((double)logprob - ((double)log_std[j] + 0.5 * (double)__fmul_rn(z, z)));
    }
    return logprob;
}

template <int TM>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_update_kernel(const FusedArgs p) {
    constexpr int TMP = TM + 4;
    extern __shared__ __align__(16) float smem[];
    const FusedNet& net = p.net;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * TM;
    const int S = net.sizes[0], OUT = net.sizes[net.L];
    float* bias_s = smem + p.smem_b_off;
    float* red = smem + p.smem_red_off;           // 64 floats
    int* src_rows = reinterpret_cast<int*>(red + 64);   // TM ints

    // ---- stage weights (transposed) + biases; resolve the source rows of this tile
    for (int l = 0; l < net.L; l++) {
        const int n_in = net.sizes[l], n_out = net.sizes[l + 1], out_pad = pad4(n_out);
        float* Wt = smem + net.wt_off[l];
        const float* W = p.params + net.w_off[l];
        for (int e = tid; e < n_in * out_pad; e += kFusedThreads) {
            const int j = e / n_in, k = e - j * n_in;    // coalesced along k in global memory
            Wt[k * out_pad + j] = (j < n_out) ? W[(size_t)j * n_in + k] : 0.f;
        }
        for (int j = tid; j < n_out; j += kFusedThreads) bias_s[net.bs_off[l] + j] = p.params[net.b_off[l] + j];
    }
    if (tid < TM) {
        const int r = row0 + tid;
        int src = -1;
        if (r < p.m) src = p.idx ? p.idx[(p.offset + r) % p.limit] : p.offset + r;
        src_rows[tid] = src;
    }
    __syncthreads();
    // ---- gather the input tile: Xt[k][r] = state[src][k]
    {
        float* Xt = smem + net.a_off[0];
        for (int e = tid; e < TM * S; e += kFusedThreads) {
            const int r = e / S, k = e - r * S;
            const int src = src_rows[r];
            Xt[k * TMP + r] = src >= 0 ? p.state[(size_t)src * S + k] : 0.f;
        }
    }
    __syncthreads();
    // ---- forward
    for (int l = 0; l < net.L; l++) {
        fused_forward_layer<TM>(smem + net.a_off[l], smem + net.wt_off[l], bias_s + net.bs_off[l],
                                smem + net.a_off[l + 1], net.sizes[l], net.sizes[l + 1], net.acts[l]);
        __syncthreads();
    }
    const float* Yt = smem + net.a_off[net.L];
    if (p.mode == kFusedForward) {
        for (int e = tid; e < TM * OUT; e += kFusedThreads) {
            const int r = e / OUT, j = e - r * OUT;
            if (row0 + r < p.m) p.y_out[(size_t)(row0 + r) * OUT + j] = Yt[j * TMP + r];
        }
        return;
    }
    float* slab = p.partials + (size_t)blockIdx.x * p.slab;
    float* Ga = smem + p.smem_g_off;
    float* Gb = Ga + net.max_width_pad * TMP;
    // ---- fused loss head: writes Ga[j][r] = dLoss/dy[r][j] (already through the output activation)
    {
        float loss_term = 0.f;
        float gls[8];
#pragma unroll
        for (int j = 0; j < 8; j++) gls[j] = 0.f;
        if (tid < TM) {
            const int src = src_rows[tid];
            float gout[8];
#pragma unroll
            for (int j = 0; j < 8; j++) gout[j] = 0.f;
            if (src >= 0) {
                if (p.mode == kFusedValue) {            // src/loss.cu:5-23
                    const float y = Yt[tid], t = p.adv_target[src];
                    gout[0] = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(y, t)), (float)p.m_total);
                    const float d = __fsub_rn(t, y);
                    loss_term = __fmul_rn(d, d);
                } else {                                // src/policy.cu:67-111 + src/ppo.cu:89-98
                    float mu[8], act[8];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j < OUT) { mu[j] = Yt[j * TMP + tid]; act[j] = p.action[(size_t)src * OUT + j]; }
                    const float lp = fused_log_prob(mu, p.log_std, act, OUT);
                    const float adv = p.advantage[src];
                    const float ratio = expf(__fsub_rn(lp, p.logprob[src]));
                    const bool adv_pos = adv > 0.f;
                    const bool hi = ratio > 1.f + p.epsilon, lo = ratio < 1.f - p.epsilon;
                    const float sel = adv_pos ? (hi ? 1.f + p.epsilon : ratio) : (lo ? 1.f - p.epsilon : ratio);
                    loss_term = __fmul_rn(adv, sel);
                    const int keep = adv_pos ? !hi : !lo;
                    const float g = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), adv), ratio), (float)p.m_total);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j < OUT) {
                            const float e2 = expf(-2.f * p.log_std[j]);
                            const float diff = __fsub_rn(act[j], mu[j]);
                            gout[j] = __fmul_rn(__fmul_rn(diff, e2), g);
                            gls[j] = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), g);
                        }
                }
            }
            const int out_act = net.acts[net.L - 1];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < OUT) Ga[j * TMP + tid] = act_grad(Yt[j * TMP + tid], gout[j], out_act);
        }
        // block reductions of the loss term and the log_std gradient (warps 0..TM/32-1 hold data)
        const int warp = tid >> 5, lane = tid & 31;
        float v = warp_sum(loss_term);
        if (lane == 0) red[warp] = v;
        if (p.mode == kFusedPolicy) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < OUT) { const float s = warp_sum(gls[j]); if (lane == 0) red[8 + j * 8 + warp] = s; }
        }
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < kFusedThreads / 32; w++) t += red[w];
            slab[net.P + OUT] = t;
        }
        if (p.mode == kFusedPolicy && tid < OUT) {
            float t = 0.f;
            for (int w = 0; w < kFusedThreads / 32; w++) t += red[8 + tid * 8 + w];
            slab[net.P + tid] = t;
        }
    }
    __syncthreads();
    // ---- backward
    float* G = Ga;
    float* Gn = Gb;
    for (int l = net.L - 1; l >= 0; l--) {
        const int n_in = net.sizes[l], n_out = net.sizes[l + 1];
        const float* Xt = smem + net.a_off[l];
        fused_weights_dispatch<TM>(G, Xt, slab + net.w_off[l], n_in, n_out);
        fused_bias_grad<TM>(G, slab + net.b_off[l], n_out);
        if (l > 0) {
            fused_backward_input<TM>(G, smem + net.wt_off[l], Xt, Gn, n_in, n_out, net.acts[l - 1]);
            __syncthreads();
            float* tmp = G; G = Gn; Gn = tmp;
        }
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
};

__device__ __forceinline__ void adam_apply(const AdamSeg& s, int i, float g) {
    float m = s.m[i], v = s.v[i], w = s.w[i];
    m = __fadd_rn(__fmul_rn(s.beta1, m), __fmul_rn(s.omb1, g));
    v = __fadd_rn(__fmul_rn(s.beta2, v), __fmul_rn(s.omb2, __fmul_rn(g, g)));
    const float denom = (float)((double)__fsqrt_rn(__fdiv_rn(v, s.bc2)) + 1e-8);
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(s.step_size, m), denom));
    s.m[i] = m; s.v[i] = v; s.w[i] = w; s.g[i] = g;
}

// One CTA per 32 consecutive slab entries; 8 warps split the slabs, fixed-order combine.
__global__ void __launch_bounds__(256) fused_reduce_adam_kernel(const ReduceAdamArgs p) {
    __shared__ float red[8][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    const int total = p.P + p.A + 1;
    float s = 0.f;
    if (e < total) {
        const int per = (p.nparts + 7) / 8;
        const int b0 = warp * per, b1 = min(p.nparts, b0 + per);
#pragma unroll 8
        for (int b = b0; b < b1; b++) s += p.partials[(size_t)b * p.slab + e];
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp != 0 || e >= total) return;
    float g = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; w++) g += red[w][lane];
    if (e < p.P) {
        adam_apply(p.net, e, g);
    } else if (e < p.P + p.A) {
        if (p.mode == kFusedPolicy) adam_apply(p.ls, e - p.P, g + (-p.ent_coeff));   // src/ppo.cu:436-438
    } else {
        if (p.mode == kFusedValue) {
            *p.loss_slot += g / (float)p.m_total;
        } else {
            float entropy = (float)(p.A * 0.5 * (1 + log(2 * kPiF)));
            for (int j = 0; j < p.A; j++) entropy += p.log_std[j];
            *p.loss_slot += -g / (float)p.m_total - p.ent_coeff * entropy;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
struct FusedPlan { bool ok; int tm; size_t smem_bytes; FusedNet net; int g_off, b_off, red_off; };

static FusedPlan make_plan(NetDev* nd, int tm) {
    FusedPlan pl{};
    pl.ok = false;
    const int L = nd->num_layers - 1;
    if (L < 1 || L > kFusedMaxLayers) return pl;
    const int tmp = tm + 4;
    const int cg = kFusedThreads / (tm / 8);        // column groups of 4
    FusedNet& n = pl.net;
    n.L = L;
    int off = 0, maxw = 4, boff = 0;
    for (int l = 0; l <= L; l++) n.sizes[l] = nd->sizes[l];
    for (int l = 0; l < L; l++) {
        n.acts[l] = nd->acts[l];
        n.w_off[l] = (int)nd->w_off[l];
        n.b_off[l] = (int)nd->b_off[l];
        if (pad4(n.sizes[l + 1]) > 4 * cg || n.sizes[l + 1] > 128) return pl;       // forward column coverage / dW tiles
        if (l > 0 && (pad4(n.sizes[l]) > 4 * cg || n.sizes[l] > 128)) return pl;   // dX column coverage
        if (n.sizes[l] > 128) return pl;
        n.wt_off[l] = off;
        off += n.sizes[l] * pad4(n.sizes[l + 1]);
        n.bs_off[l] = boff;
        boff += pad4(n.sizes[l + 1]);
    }
    if (n.sizes[L] > 8) return pl;                  // loss heads keep <= 8 outputs in registers
    n.P = (int)nd->param_count;
    for (int l = 0; l <= L; l++) {
        n.a_off[l] = off;
        off += pad4(n.sizes[l]) * tmp;
        if (l > 0) maxw = std::max(maxw, pad4(n.sizes[l]));
    }
    n.max_width_pad = maxw;
    pl.g_off = off;
    off += 2 * maxw * tmp;
    pl.b_off = off;
    off += boff;
    pl.red_off = off;
    off += 64 + tm;
    pl.smem_bytes = (size_t)off * sizeof(float);
    pl.tm = tm;
    pl.ok = pl.smem_bytes <= 220 * 1024;
    return pl;
}

static FusedPlan choose_plan(NetDev* nd) {
    FusedPlan p = make_plan(nd, 128);
    if (p.ok) return p;
    return make_plan(nd, 64);
}

bool fused_supported(NeuralNetwork* nn) { return choose_plan(net_dev(nn)).ok; }

static void launch_fused(const FusedPlan& pl, FusedArgs& a) {
    a.net = pl.net;
    a.smem_g_off = pl.g_off;
    a.smem_b_off = pl.b_off;
    a.smem_red_off = pl.red_off;
    const int blocks = div_up(a.m, pl.tm);
    static size_t configured[2] = {0, 0};
    if (pl.tm == 128) {
        if (pl.smem_bytes > configured[0]) {
            CUDA_CHECK(cudaFuncSetAttribute(fused_update_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
            configured[0] = pl.smem_bytes;
        }
        B200_LAUNCH(fused_update_kernel<128>, blocks, kFusedThreads, pl.smem_bytes, a);
    } else {
        if (pl.smem_bytes > configured[1]) {
            CUDA_CHECK(cudaFuncSetAttribute(fused_update_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
            configured[1] = pl.smem_bytes;
        }
        B200_LAUNCH(fused_update_kernel<64>, blocks, kFusedThreads, pl.smem_bytes, a);
    }
}

void fused_forward(NeuralNetwork* nn, const float* x, int m, float* y_out) {
    NetDev* nd = net_dev(nn);
    const FusedPlan pl = choose_plan(nd);
    if (!pl.ok) B200_FATAL("fused_forward on an unsupported net");
    FusedArgs a{};
    a.params = nd->params;
    a.idx = nullptr; a.offset = 0; a.limit = m; a.m = m; a.m_total = m; a.mode = kFusedForward;
    a.state = x; a.y_out = y_out;
    launch_fused(pl, a);
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
// Returns false when the net is outside the fused kernel's limits (caller uses the layer-wise path).
bool fused_minibatch_update(NeuralNetwork* nn, GaussianPolicy* policy, Adam* adam_net, Adam* adam_ls, float lr,
                            const int* perm, int offset, int limit, int m, int m_total, const TrajectoryBuffer* b,
                            float epsilon, float ent_coeff, float* loss_slot, bool apply_adam) {
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
    a.params = nd->params;
    a.idx = perm; a.offset = offset; a.limit = limit; a.m = m; a.m_total = m_total;
    a.mode = policy ? kFusedPolicy : kFusedValue;
    a.state = b->d_state_p; a.action = b->d_action_p; a.logprob = b->d_logprob_p;
    a.advantage = b->d_advantage_p; a.adv_target = b->d_adv_target_p;
    a.log_std = policy ? policy->d_log_std : nullptr;
    a.epsilon = epsilon; a.ent_coeff = ent_coeff;
    a.partials = nd->partials; a.slab = slab;
    launch_fused(pl, a);
    nd->last_splits = blocks;

    ReduceAdamArgs r{};
    r.partials = nd->partials; r.nparts = blocks; r.slab = slab; r.P = (int)nd->param_count; r.A = A;
    r.mode = a.mode; r.m_total = m_total; r.ent_coeff = ent_coeff; r.loss_slot = loss_slot;
    r.log_std = a.log_std;
    if (!apply_adam) return true;      // data-parallel callers reduce + all-reduce + update themselves
    adam_net->time_step += 1;
    r.net = make_seg(nd->params, nd->grads, adam_net, lr);
    if (policy) {
        adam_ls->time_step += 1;
        r.ls = make_seg(policy->d_log_std, policy->d_log_std_grad, adam_ls, lr);
    }
    B200_LAUNCH(fused_reduce_adam_kernel, div_up(slab, 32), 256, 0, r);
    return true;
}

}  // namespace b200
