// gae.cu — GAE / returns / advantage normalisation.
//
// Replaces compute_gae's scalar loops (reference src/ppo.cu:338-368) and the CUDA twin's K1-K4
// (src/ppo.cu:171-259, 286-316; only correct for horizon < 512).
//
//   delta_i = r_i + gamma * v'_i * !term_i - v_i                         (src/ppo.cu:340-342)
//   adv_i   = delta_i + (gamma*lambda * !(trunc_i || term_i)) * adv_{i+1} (src/ppo.cu:344-349)
//   target_i= v_i + adv_i                                                 (src/ppo.cu:351-353)
//
// The recurrence is the associative operator (c1,a1) o (c2,a2) = (c1*c2, a1 + c1*a2) on pairs
// c_i = gamma*lambda*(1-done_i), a_i = delta_i, applied from the END of the flat buffer.  One
// single-pass kernel does a segmented reverse scan over the whole flat buffer:
//   * a warp owns a chunk of 512 consecutive elements, held entirely in registers
//     (4 tiles x 32 lanes x float4; all global loads/stores are 128-bit, coalesced, L1-bypassing);
//   * inside a tile: 4-element serial recurrence per lane + 5-step warp-shuffle suffix scan;
//   * across chunks: decoupled look-back.  Chunks are claimed in DESCENDING order through an
//     atomic ticket, each publishes (C_total, A_first) then its inclusive A_first with
//     st.release; a successor chunk is therefore always resident or finished -> no deadlock.
//     A chunk whose last element is done (every env stream of a TxN buffer) never looks back.
// Algorithmic HBM traffic: 12 B (r,v,v') + 2 B (flags) read + 8 B written = 22 B/element; the
// optional normalisation is a second 8 B/element pass -> 30 B/element.
//
// Normalisation statistics: per-chunk (mean, M2, n) Welford triples, combined in a fixed order in
// float64 (merge formula of reference include/welford_var.h:58-66).  The reference's own float
// accumulators drift at large B; the float64 restatement is the arbiter.
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr int kGaeWarps = 8;
constexpr int kGaeTiles = 4;
constexpr int kGaeTile = 128;                         // 32 lanes x 4
constexpr int kGaeChunk = kGaeTile * kGaeTiles;       // 512 elements per warp
constexpr int kGaeBlockElems = kGaeChunk * kGaeWarps; // 4096 elements per CTA

struct GaeDesc {                 // each word = (epoch << 32) | float bits
    unsigned long long aggA;     // A_first with zero carry-in
    unsigned long long aggC;     // product of c over the chunk
    unsigned long long inc;      // A_first including everything behind the chunk
    unsigned long long pad;
};

__device__ __forceinline__ unsigned long long pack(unsigned epoch, float x) {
    return ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(x);
}

__global__ void __launch_bounds__(kGaeWarps * 32)
gae_scan_kernel(const float* __restrict__ reward, const float* __restrict__ v,
                const float* __restrict__ v_next, const unsigned char* __restrict__ terminated,
                const unsigned char* __restrict__ truncated, int n, float gamma, float gl,
                float* __restrict__ adv_out, float* __restrict__ target_out, GaeDesc* desc,
                int* ticket, unsigned epoch, float4* __restrict__ wstats, int vec_ok) {
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    const int cb = (int)gridDim.x - 1 - s_ticket;      // descending chunk order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (n + kGaeChunk - 1) / kGaeChunk;
    const int wc = cb * kGaeWarps + warp;
    if (wc >= nchunks) return;
    const long long base = (long long)wc * kGaeChunk;

    float dl[kGaeTiles][4], cc[kGaeTiles][4], vv[kGaeTiles][4];
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
        float r4[4], n4[4];
        unsigned char te[4], tr[4];
        if (vec_ok && i0 + 3 < n) {
            const float4 a = ld_stream4(reward + i0), b = ld_stream4(v + i0), c = ld_stream4(v_next + i0);
            const uint32_t ft = ld_stream_u32(terminated + i0), fr = ld_stream_u32(truncated + i0);
            r4[0] = a.x; r4[1] = a.y; r4[2] = a.z; r4[3] = a.w;
            vv[t][0] = b.x; vv[t][1] = b.y; vv[t][2] = b.z; vv[t][3] = b.w;
            n4[0] = c.x; n4[1] = c.y; n4[2] = c.z; n4[3] = c.w;
#pragma unroll
            for (int e = 0; e < 4; e++) { te[e] = (ft >> (8 * e)) & 0xff; tr[e] = (fr >> (8 * e)) & 0xff; }
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool ok = i0 + e < n;
                r4[e] = ok ? reward[i0 + e] : 0.f;
                vv[t][e] = ok ? v[i0 + e] : 0.f;
                n4[e] = ok ? v_next[i0 + e] : 0.f;
                te[e] = ok ? terminated[i0 + e] : 1;   // padding behaves as a terminated, zero-delta step
                tr[e] = ok ? truncated[i0 + e] : 1;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float nt = te[e] ? 0.f : 1.f;
            dl[t][e] = r4[e] + gamma * n4[e] * nt - vv[t][e];
            cc[t][e] = (te[e] || tr[e]) ? 0.f : gl;
        }
    }

    // ---- local scan (zero carry-in): loc = advantage, cend = product of c from element to chunk end
    float loc[kGaeTiles][4], cend[kGaeTiles][4];
    float X = 0.f, Cx = 1.f;   // head of the already-scanned suffix (tiles behind this one)
#pragma unroll
    for (int t = kGaeTiles - 1; t >= 0; t--) {
        float A = 0.f, C = 1.f;
#pragma unroll
        for (int e = 3; e >= 0; e--) { A = dl[t][e] + cc[t][e] * A; C = cc[t][e] * C; }
        float Ai = A, Ci = C;  // inclusive suffix over lanes lane..31
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float Ao = __shfl_down_sync(kFull, Ai, off), Co = __shfl_down_sync(kFull, Ci, off);
            if (lane + off < 32) { Ai = Ai + Ci * Ao; Ci = Ci * Co; }
        }
        const float An = __shfl_down_sync(kFull, Ai, 1), Cn = __shfl_down_sync(kFull, Ci, 1);
        float a = (lane == 31) ? X : An + Cn * X;
        float ce = (lane == 31) ? Cx : Cn * Cx;
#pragma unroll
        for (int e = 3; e >= 0; e--) {
            a = dl[t][e] + cc[t][e] * a;
            ce = cc[t][e] * ce;
            loc[t][e] = a;
            cend[t][e] = ce;
        }
        X = __shfl_sync(kFull, loc[t][0], 0);
        Cx = __shfl_sync(kFull, cend[t][0], 0);
    }

    // ---- publish the chunk aggregate, then resolve the carry-in by looking at later chunks
    if (lane == 0) {
        st_release_u64(&desc[wc].aggA, pack(epoch, X));
        st_release_u64(&desc[wc].aggC, pack(epoch, Cx));
    }
    const float c_last = __shfl_sync(kFull, cend[kGaeTiles - 1][3], 31);
    float carry = 0.f;
    if (c_last != 0.f && wc + 1 < nchunks) {   // warp-uniform
        if (lane == 0) {
            float mult = 1.f;
            int j = wc + 1;
            while (j < nchunks) {
                const unsigned long long inc = ld_acquire_u64(&desc[j].inc);
                if ((unsigned)(inc >> 32) == epoch) { carry += mult * __uint_as_float((unsigned)inc); break; }
                const unsigned long long a = ld_acquire_u64(&desc[j].aggA);
                const unsigned long long c = ld_acquire_u64(&desc[j].aggC);
                if ((unsigned)(a >> 32) == epoch && (unsigned)(c >> 32) == epoch) {
                    carry += mult * __uint_as_float((unsigned)a);
                    mult *= __uint_as_float((unsigned)c);
                    if (mult == 0.f) break;
                    j++;
                } else {
                    __nanosleep(40);
                }
            }
        }
        carry = __shfl_sync(kFull, carry, 0);
    }
    if (lane == 0) st_release_u64(&desc[wc].inc, pack(epoch, X + Cx * carry));

    // ---- final values, stores, Welford partial
    float sum = 0.f;
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
        float a4[4], t4[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            a4[e] = loc[t][e] + cend[t][e] * carry;
            t4[e] = vv[t][e] + a4[e];
            loc[t][e] = a4[e];
            if (i0 + e < n) { sum += a4[e]; cnt++; }
        }
        if (vec_ok && i0 + 3 < n) {
            st_stream4(adv_out + i0, make_float4(a4[0], a4[1], a4[2], a4[3]));
            st_stream4(target_out + i0, make_float4(t4[0], t4[1], t4[2], t4[3]));
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (i0 + e < n) { adv_out[i0 + e] = a4[e]; target_out[i0 + e] = t4[e]; }
        }
    }
    sum = warp_sum(sum);
    int total = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
    const float mean = sum / (float)total;
    float m2 = 0.f;
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (i0 + e < n) { const float d = loc[t][e] - mean; m2 += d * d; }
    }
    m2 = warp_sum(m2);
    if (lane == 0) wstats[wc] = make_float4(mean, m2, (float)total, 0.f);
}

// Fixed-order float64 combine of the chunk triples (welford_var.h:58-66 merge).  One CTA.
// out_d = {mean, M2, n} (float64, for the cross-rank merge), out_f = {mean, std} (float32).
__global__ void __launch_bounds__(1024)
gae_stats_kernel(const float4* __restrict__ wstats, int nchunks, double* out_d, float* out_f) {
    __shared__ double s_mean[1024], s_m2[1024], s_n[1024];
    const int tid = threadIdx.x;
    // contiguous slab per thread -> fixed combine order regardless of scheduling
    const int per = (nchunks + 1023) / 1024;
    double mean = 0.0, m2 = 0.0, cnt = 0.0;
    for (int i = tid * per; i < min(nchunks, (tid + 1) * per); i++) {
        const float4 w = wstats[i];
        const double nb = w.z;
        if (nb > 0.0) {
            const double delta = (double)w.x - mean, nn = cnt + nb;
            mean += delta * nb / nn;
            m2 += (double)w.y + delta * delta * cnt * nb / nn;
            cnt = nn;
        }
    }
    s_mean[tid] = mean; s_m2[tid] = m2; s_n[tid] = cnt;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (tid < s) {
            const double na = s_n[tid], nb = s_n[tid + s];
            if (nb > 0.0) {
                const double delta = s_mean[tid + s] - s_mean[tid], nn = na + nb;
                s_mean[tid] += delta * nb / nn;
                s_m2[tid] += s_m2[tid + s] + delta * delta * na * nb / nn;
                s_n[tid] = nn;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        out_d[0] = s_mean[0]; out_d[1] = s_m2[0]; out_d[2] = s_n[0];
        out_f[0] = (float)s_mean[0];
        out_f[1] = (float)sqrt(s_m2[0] / s_n[0]);   // population std, src/ppo.cu:362
    }
}

// Merge the per-rank {mean, M2, n} triples (rank order) into the global {mean, std}.
__global__ void gae_merge_ranks_kernel(const double* __restrict__ triples, int world, float* out_f) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double mean = 0.0, m2 = 0.0, cnt = 0.0;
    for (int r = 0; r < world; r++) {
        const double mb = triples[3 * r], m2b = triples[3 * r + 1], nb = triples[3 * r + 2];
        if (nb > 0.0) {
            const double delta = mb - mean, nn = cnt + nb;
            mean += delta * nb / nn;
            m2 += m2b + delta * delta * cnt * nb / nn;
            cnt = nn;
        }
    }
    out_f[0] = (float)mean;
    out_f[1] = (float)sqrt(m2 / cnt);
}

// adv <- (adv - mean) / (std + 1e-8)   (src/ppo.cu:366-368; the double add is the reference's)
__global__ void __launch_bounds__(256)
gae_normalize_kernel(float* __restrict__ adv, int n, const float* __restrict__ stats, int vec_ok) {
    const float mean = stats[0];
    const float denom = (float)((double)stats[1] + 1e-8);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec_ok ? n / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = ld_stream4(adv + 4 * i);
        a.x = (a.x - mean) / denom; a.y = (a.y - mean) / denom;
        a.z = (a.z - mean) / denom; a.w = (a.w - mean) / denom;
        st_stream4(adv + 4 * i, a);
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adv[i] = (adv[i] - mean) / denom;
}

// ---- host side ------------------------------------------------------------------------------------
static unsigned g_gae_epoch = 0;

GaeWork gae_scan(const float* reward, const float* v, const float* v_next, const bool* terminated,
                 const bool* truncated, int n, float gamma, float lambda, float* advantage,
                 float* adv_target) {
    GaeWork w{};
    if (n <= 0) return w;
    const int nchunks = div_up(n, kGaeChunk);
    const int nblocks = div_up(n, kGaeBlockElems);
    // layout: [ticket (256 B)] [stats_d 3 doubles + stats_f 2 floats (256 B)] [desc] [wstats]
    const size_t desc_bytes = (size_t)nchunks * sizeof(GaeDesc);
    const size_t bytes = 512 + desc_bytes + (size_t)nchunks * sizeof(float4);
    char* ws = static_cast<char*>(scratch(kScratchGae, bytes));
    int* ticket = reinterpret_cast<int*>(ws);
    w.stats_d = reinterpret_cast<double*>(ws + 256);
    w.stats_f = reinterpret_cast<float*>(ws + 256 + 64);
    GaeDesc* desc = reinterpret_cast<GaeDesc*>(ws + 512);
    float4* wstats = reinterpret_cast<float4*>(ws + 512 + desc_bytes);
    w.nchunks = nchunks;
    w.wstats = wstats;
    const unsigned epoch = ++g_gae_epoch;
    if (epoch == 0) B200_FATAL("GAE epoch counter wrapped");
    CUDA_CHECK(cudaMemsetAsync(ticket, 0, sizeof(int), stream()));
    const float gl = gamma * lambda;  // formed first in float, src/ppo.cu:346
    const uintptr_t al = (uintptr_t)reward | (uintptr_t)v | (uintptr_t)v_next | (uintptr_t)advantage | (uintptr_t)adv_target;
    const uintptr_t alb = (uintptr_t)terminated | (uintptr_t)truncated;
    const int vec_ok = ((al & 15) == 0 && (alb & 3) == 0) ? 1 : 0;
    B200_LAUNCH(gae_scan_kernel, nblocks, kGaeWarps * 32, 0, reward, v, v_next,
                reinterpret_cast<const unsigned char*>(terminated),
                reinterpret_cast<const unsigned char*>(truncated), n, gamma, gl, advantage,
                adv_target, desc, ticket, epoch, wstats, vec_ok);
    B200_LAUNCH(gae_stats_kernel, 1, 1024, 0, wstats, nchunks, w.stats_d, w.stats_f);
    return w;
}

void gae_merge_ranks(const double* triples_dev, int world, float* stats_f) {
    B200_LAUNCH(gae_merge_ranks_kernel, 1, 32, 0, triples_dev, world, stats_f);
}

void gae_normalize(float* advantage, int n, const float* stats_f) {
    if (n <= 0) return;
    const int vec_ok = (((uintptr_t)advantage) & 15) == 0;
    const int blocks = (int)std::min<long long>(div_up(div_up(n, 4), 256), (long long)num_sms() * 8);
    B200_LAUNCH(gae_normalize_kernel, blocks, 256, 0, advantage, n, stats_f, vec_ok);
}

}  // namespace b200

extern "C" void ppo_b200_gae(const float* reward, const float* v, const float* v_next,
                             const bool* terminated, const bool* truncated, int n, float gamma,
                             float lambda, float* advantage, float* adv_target, int normalize,
                             float* stats_out) {
    using namespace b200;
    GaeWork w = gae_scan(reward, v, v_next, terminated, truncated, n, gamma, lambda, advantage, adv_target);
    if (n <= 0) return;
    if (normalize) gae_normalize(advantage, n, w.stats_f);
    if (stats_out)
        CUDA_CHECK(cudaMemcpyAsync(stats_out, w.stats_f, 2 * sizeof(float), cudaMemcpyDeviceToDevice, stream()));
}
