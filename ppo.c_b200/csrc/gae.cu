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
//     atomic ticket; each publishes ONE self-contained 64-bit word {epoch, state, A_first} — first its
//     aggregate (state says whether the chunk contains a done, i.e. whether C_total is 0 or
//     (gamma*lambda)^512), later its inclusive value.  Because flag and payload share one naturally
//     atomic word no fence is needed (an ncu pass showed st.release membars costing 27% of the stall
//     samples); a successor chunk is always resident or finished -> no deadlock.
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
constexpr int kGaeTile = 128;                         // 32 lanes x 4

// One descriptor word per chunk: [63:34] epoch, [33:32] state, [31:0] float bits of A_first.
enum { kGaeAggDone = 1u, kGaeAggOpen = 2u, kGaeInclusive = 3u };
typedef unsigned long long GaeDesc;

__device__ __forceinline__ unsigned long long pack(unsigned epoch, unsigned state, float x) {
    return ((unsigned long long)epoch << 34) | ((unsigned long long)state << 32) | (unsigned long long)__float_as_uint(x);
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}

template <int kGaeTiles>
struct GaeRaw {                      // one chunk as it comes out of HBM: 14 registers per tile and lane
    float4 r[kGaeTiles], v[kGaeTiles], vn[kGaeTiles];
    uint32_t ft[kGaeTiles], fr[kGaeTiles];     // 4 terminated / truncated bytes each
};

template <int kGaeTiles, bool HINT>
__device__ __forceinline__ void gae_load_chunk(GaeRaw<kGaeTiles>& raw, long long base, int lane, int n, int vec_ok,
                                               const float* __restrict__ reward, const float* __restrict__ v,
                                               const float* __restrict__ v_next, const unsigned char* __restrict__ terminated,
                                               const unsigned char* __restrict__ truncated) {
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
        if (vec_ok && i0 + 3 < n) {
            if (HINT) {
                raw.r[t] = ld_stream4(reward + i0);
                raw.v[t] = ld_stream4(v + i0);
                raw.vn[t] = ld_stream4(v_next + i0);
                raw.ft[t] = ld_stream_u32(terminated + i0);
                raw.fr[t] = ld_stream_u32(truncated + i0);
            } else {
                raw.r[t] = __ldg(reinterpret_cast<const float4*>(reward + i0));
                raw.v[t] = __ldg(reinterpret_cast<const float4*>(v + i0));
                raw.vn[t] = __ldg(reinterpret_cast<const float4*>(v_next + i0));
                raw.ft[t] = __ldg(reinterpret_cast<const uint32_t*>(terminated + i0));
                raw.fr[t] = __ldg(reinterpret_cast<const uint32_t*>(truncated + i0));
            }
        } else {   // ragged tail / unaligned buffers: padding behaves as a terminated, zero-delta step
            float rr[4], vv[4], nn[4];
            uint32_t ft = 0, fr = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool ok = i0 + e < n;
                rr[e] = ok ? reward[i0 + e] : 0.f;
                vv[e] = ok ? v[i0 + e] : 0.f;
                nn[e] = ok ? v_next[i0 + e] : 0.f;
                ft |= (uint32_t)(ok ? (terminated[i0 + e] != 0) : 1) << (8 * e);
                fr |= (uint32_t)(ok ? (truncated[i0 + e] != 0) : 1) << (8 * e);
            }
            raw.r[t] = make_float4(rr[0], rr[1], rr[2], rr[3]);
            raw.v[t] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            raw.vn[t] = make_float4(nn[0], nn[1], nn[2], nn[3]);
            raw.ft[t] = ft;
            raw.fr[t] = fr;
        }
    }
}

// Scan one chunk held in registers: local suffix scan, look-back for the carry-in, outputs, Welford partial.
template <int kGaeTiles, bool HINT, bool FULL>      // FULL: the chunk lies entirely inside [0, n) and the arrays are 16-byte aligned
__device__ __forceinline__ void gae_process_chunk(const GaeRaw<kGaeTiles>& cur, int wc, int nchunks, int lane, int n, int vec_ok,
                                                  float gamma, float gl, float cfull, float* __restrict__ adv_out,
                                                  float* __restrict__ target_out, GaeDesc* desc, unsigned epoch,
                                                  float4* __restrict__ wstats) {
    constexpr int kGaeChunk = kGaeTile * kGaeTiles;
    const long long base = (long long)wc * kGaeChunk;
    float dl[kGaeTiles][4], cc[kGaeTiles][4];
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const float r4[4] = {cur.r[t].x, cur.r[t].y, cur.r[t].z, cur.r[t].w};
        const float v4[4] = {cur.v[t].x, cur.v[t].y, cur.v[t].z, cur.v[t].w};
        const float n4[4] = {cur.vn[t].x, cur.vn[t].y, cur.vn[t].z, cur.vn[t].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const bool te = (cur.ft[t] >> (8 * e)) & 0xff, tr = (cur.fr[t] >> (8 * e)) & 0xff;
            dl[t][e] = r4[e] + gamma * n4[e] * (te ? 0.f : 1.f) - v4[e];
            cc[t][e] = (te || tr) ? 0.f : gl;
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
    // ---- publish the chunk aggregate, then resolve the carry-in by looking at later chunks.
    // Cx is exactly 0 when the chunk holds a done step, else (gamma*lambda)^512 up to rounding.
    if (lane == 0) st_relaxed_u64(&desc[wc], pack(epoch, Cx == 0.f ? kGaeAggDone : kGaeAggOpen, X));
    const float c_last = __shfl_sync(kFull, cend[kGaeTiles - 1][3], 31);
    float carry = 0.f;
    if (c_last != 0.f && wc + 1 < nchunks) {   // warp-uniform
        if (lane == 0) {
            float mult = 1.f;
            int j = wc + 1;
            while (j < nchunks) {
                const unsigned long long w = ld_relaxed_u64(&desc[j]);
                if ((unsigned)(w >> 34) != epoch) { __nanosleep(20); continue; }
                const unsigned state = (unsigned)(w >> 32) & 3u;
                carry += mult * __uint_as_float((unsigned)w);
                if (state != kGaeAggOpen) break;            // inclusive value, or the chunk cuts the chain
                mult *= cfull;
                if (mult == 0.f) break;
                j++;
            }
        }
        carry = __shfl_sync(kFull, carry, 0);
    }
    if (lane == 0 && Cx != 0.f) st_relaxed_u64(&desc[wc], pack(epoch, kGaeInclusive, X + Cx * carry));

    // ---- final values, stores, Welford partial
    float sum = 0.f;
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
        const float v4[4] = {cur.v[t].x, cur.v[t].y, cur.v[t].z, cur.v[t].w};
        float a4[4], t4[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            a4[e] = loc[t][e] + cend[t][e] * carry;
            t4[e] = v4[e] + a4[e];
            loc[t][e] = a4[e];
            if (i0 + e < n) { sum += a4[e]; cnt++; }
        }
        if (FULL || (vec_ok && i0 + 3 < n)) {
            if (HINT) {
                st_stream4(adv_out + i0, make_float4(a4[0], a4[1], a4[2], a4[3]));
                st_stream4(target_out + i0, make_float4(t4[0], t4[1], t4[2], t4[3]));
            } else {
                *reinterpret_cast<float4*>(adv_out + i0) = make_float4(a4[0], a4[1], a4[2], a4[3]);
                *reinterpret_cast<float4*>(target_out + i0) = make_float4(t4[0], t4[1], t4[2], t4[3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (i0 + e < n) { adv_out[i0 + e] = a4[e]; target_out[i0 + e] = t4[e]; }
        }
    }
    sum = warp_sum(sum);
    int total = cnt;
    if (FULL) total = kGaeChunk;
    else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
    }
    const float mean = sum / (float)total;
    float m2 = 0.f;
#pragma unroll
    for (int t = 0; t < kGaeTiles; t++) {
        const long long i0 = base + t * kGaeTile + lane * 4;
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (FULL || i0 + e < n) { const float d = loc[t][e] - mean; m2 += d * d; }
    }
    m2 = warp_sum(m2);
    if (lane == 0) wstats[wc] = make_float4(mean, m2, (float)total, 0.f);
}

// Persistent kernel: global warp g handles chunks nchunks-1-(i*W+g), i = 0,1,...  (descending order, so a
// chunk's successor is handled in the same or an earlier iteration by a resident warp -> the look-back
// cannot deadlock as long as the whole grid is co-resident, which the host guarantees through the
// occupancy API).  While a chunk is being scanned the NEXT chunk's 20 loads per lane are already in
// flight (register double buffering), so every warp keeps HBM requests outstanding all the time.
template <int kGaeTiles, bool HINT, int MINB>
__global__ void __launch_bounds__(kGaeWarps * 32, MINB)
gae_scan_kernel(const float* __restrict__ reward, const float* __restrict__ v,
                const float* __restrict__ v_next, const unsigned char* __restrict__ terminated,
                const unsigned char* __restrict__ truncated, int n, float gamma, float gl,
                float* __restrict__ adv_out, float* __restrict__ target_out, GaeDesc* desc,
                unsigned epoch, float4* __restrict__ wstats, int vec_ok) {
    constexpr int kGaeChunk = kGaeTile * kGaeTiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (n + kGaeChunk - 1) / kGaeChunk;
    const int W = gridDim.x * kGaeWarps;
    const int g = blockIdx.x * kGaeWarps + warp;
    int wc = nchunks - 1 - g;
    if (wc < 0) return;
    float cfull = gl;
#pragma unroll
    for (int q = 1; q < kGaeChunk; q <<= 1) cfull *= cfull;     // gl^chunk: C_total of a chunk without a done step

    GaeRaw<kGaeTiles> cur;
    gae_load_chunk<kGaeTiles, HINT>(cur, (long long)wc * kGaeChunk, lane, n, vec_ok, reward, v, v_next, terminated, truncated);
    for (; wc >= 0; wc -= W) {
        GaeRaw<kGaeTiles> nxt;
        const bool has_next = wc - W >= 0;
        if (has_next) gae_load_chunk<kGaeTiles, HINT>(nxt, (long long)(wc - W) * kGaeChunk, lane, n, vec_ok, reward, v, v_next, terminated, truncated);
        gae_process_chunk<kGaeTiles, HINT, false>(cur, wc, nchunks, lane, n, vec_ok, gamma, gl, cfull, adv_out, target_out, desc, epoch, wstats);
        if (has_next) cur = nxt;
    }
}

// ---- TMA-staged variant (n a multiple of 512, 16-byte aligned arrays) ---------------------------------------------------
// Same schedule, but the inputs of the next kGaeStages chunks of every warp are in flight as 1-D bulk copies
// (cp.async.bulk global -> shared, one mbarrier per stage and warp: SASS UBLKCP + SYNCS) instead of one register-buffered
// chunk: 16 warps x 2 stages x 7 KB = 224 KB of requests outstanding per SM.  The register version is bound by bytes in
// flight (an ncu pass attributed 48 % of its stall samples to the first use of the prefetched chunk).
constexpr int kGaeStages = 2;
constexpr int kGaeTmaWarps = 16;              // 16 warps x 2 stages x 7 KB = 224 KB in flight per SM, 4 warps per scheduler
constexpr int kGaeStageBytes = 3 * 512 * 4 + 2 * 512;          // r, v, v' (2 KB each) + two flag arrays (512 B each)

__device__ __forceinline__ uint32_t gae_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kGaeTmaWarps * 32, 1)
gae_scan_tma_kernel(const float* __restrict__ reward, const float* __restrict__ v,
                    const float* __restrict__ v_next, const unsigned char* __restrict__ terminated,
                    const unsigned char* __restrict__ truncated, int n, float gamma, float gl,
                    float* __restrict__ adv_out, float* __restrict__ target_out, GaeDesc* desc,
                    unsigned epoch, float4* __restrict__ wstats) {
    constexpr int kTiles = 4, kChunk = kGaeTile * kTiles;      // 512 elements
    extern __shared__ __align__(128) unsigned char gsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* stage0 = gsm + (size_t)warp * kGaeStages * kGaeStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + (size_t)kGaeTmaWarps * kGaeStages * kGaeStageBytes) + warp * kGaeStages;
    const int nchunks = n / kChunk;
    const int W = gridDim.x * kGaeTmaWarps;
    const int g = blockIdx.x * kGaeTmaWarps + warp;
    const int wc0 = nchunks - 1 - g;
    if (wc0 < 0) return;
    float cfull = gl;
#pragma unroll
    for (int q = 1; q < kChunk; q <<= 1) cfull *= cfull;

    auto issue = [&](int wc, int s) {                          // lane 0 only
        unsigned char* dst = stage0 + (size_t)s * kGaeStageBytes;
        const uint32_t bar = gae_smem_u32(&bars[s]);
        const size_t e0 = (size_t)wc * kChunk;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)kGaeStageBytes) : "memory");
        const void* src[5] = {reward + e0, v + e0, v_next + e0, terminated + e0, truncated + e0};
        const uint32_t off[5] = {0u, 2048u, 4096u, 6144u, 6656u}, bytes[5] = {2048u, 2048u, 2048u, 512u, 512u};
#pragma unroll
        for (int q = 0; q < 5; q++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(gae_smem_u32(dst + off[q])), "l"(src[q]), "r"(bytes[q]), "r"(bar) : "memory");
    };
    if (lane == 0) {
        for (int s = 0; s < kGaeStages; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(gae_smem_u32(&bars[s])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kGaeStages; s++)
            if (wc0 - s * W >= 0) issue(wc0 - s * W, s);
    }
    __syncwarp();
    int it = 0;
    for (int wc = wc0; wc >= 0; wc -= W, it++) {
        const int s = it % kGaeStages;
        const uint32_t parity = (uint32_t)(it / kGaeStages) & 1u;
        {
            const uint32_t bar = gae_smem_u32(&bars[s]);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "GW_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra GD_%=;\n\t"
                "bra GW_%=;\n\t"
                "GD_%=:\n\t}"
                :: "r"(bar), "r"(parity) : "memory");
        }
        const unsigned char* sb = stage0 + (size_t)s * kGaeStageBytes;
        GaeRaw<kTiles> cur;
#pragma unroll
        for (int t = 0; t < kTiles; t++) {
            const int o = t * kGaeTile + lane * 4;
            cur.r[t] = *reinterpret_cast<const float4*>(sb + 4 * o);
            cur.v[t] = *reinterpret_cast<const float4*>(sb + 2048 + 4 * o);
            cur.vn[t] = *reinterpret_cast<const float4*>(sb + 4096 + 4 * o);
            cur.ft[t] = *reinterpret_cast<const uint32_t*>(sb + 6144 + o);
            cur.fr[t] = *reinterpret_cast<const uint32_t*>(sb + 6656 + o);
        }
        __syncwarp();                                           // every lane has its copy: the stage can be refilled
        if (lane == 0 && wc - kGaeStages * W >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(wc - kGaeStages * W, s);
        }
        gae_process_chunk<kTiles, true, true>(cur, wc, nchunks, lane, n, 1, gamma, gl, cfull, adv_out, target_out, desc, epoch, wstats);
    }
}

// Fixed-order float64 combine of the chunk triples (welford_var.h:58-66 merge), two levels:
// level 1: every CTA folds 4096 chunk triples into one double triple; level 2: one CTA folds those.
// out_d = {mean, M2, n} (float64, for the cross-rank merge), out_f = {mean, std} (float32).
constexpr int kStatsPerBlock = 4096;

__global__ void __launch_bounds__(256)
gae_stats_partial_kernel(const float4* __restrict__ wstats, int nchunks, double* __restrict__ part /* [blocks][3] */) {
    __shared__ double s_mean[256], s_m2[256], s_n[256];
    const int tid = threadIdx.x;
    const int base = blockIdx.x * kStatsPerBlock + tid * (kStatsPerBlock / 256);
    double mean = 0.0, m2 = 0.0, cnt = 0.0;
#pragma unroll 4
    for (int i = base; i < min(nchunks, base + kStatsPerBlock / 256); i++) {
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
    for (int s = 128; s > 0; s >>= 1) {
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
    if (tid == 0) { part[3 * blockIdx.x] = s_mean[0]; part[3 * blockIdx.x + 1] = s_m2[0]; part[3 * blockIdx.x + 2] = s_n[0]; }
}

__global__ void __launch_bounds__(1024)
gae_stats_kernel(const double* __restrict__ part, int nparts, double* out_d, float* out_f) {
    __shared__ double s_mean[1024], s_m2[1024], s_n[1024];
    const int tid = threadIdx.x;
    // contiguous slab per thread -> fixed combine order regardless of scheduling
    const int per = (nparts + 1023) / 1024;
    double mean = 0.0, m2 = 0.0, cnt = 0.0;
    for (int i = tid * per; i < min(nparts, (tid + 1) * per); i++) {
        const double mb = part[3 * i], m2b = part[3 * i + 1], nb = part[3 * i + 2];
        if (nb > 0.0) {
            const double delta = mb - mean, nn = cnt + nb;
            mean += delta * nb / nn;
            m2 += m2b + delta * delta * cnt * nb / nn;
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
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {          // four independent 128-bit loads in flight per thread
        float4 a[4];
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] = ld_stream4(adv + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u].x = (a[u].x - mean) / denom; a[u].y = (a[u].y - mean) / denom;
            a[u].z = (a[u].z - mean) / denom; a[u].w = (a[u].w - mean) / denom;
            st_stream4(adv + 4 * (i + u * stride), a[u]);
        }
    }
    for (; i < n4; i += stride) {
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
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("PPO_B200_GAE_VARIANT"); variant = e ? atoi(e) : 4; }   // 4 = <4 tiles, streaming hints, 1 CTA/SM>: fastest measured (profiles/)
    const int tiles = (variant == 1 || variant == 3) ? 2 : 4;
    const int chunk = kGaeTile * tiles;
    const int nchunks = div_up(n, chunk);
    const int nblocks = div_up(nchunks, kGaeWarps);
    // layout: [ticket (256 B)] [stats_d 3 doubles + stats_f 2 floats (256 B)] [desc] [wstats] [level-1 stats]
    const size_t desc_bytes = (((size_t)nchunks * sizeof(GaeDesc)) + 255) & ~size_t(255);
    const int nparts = div_up(nchunks, kStatsPerBlock);
    const size_t bytes = 512 + desc_bytes + (size_t)nchunks * sizeof(float4) + (size_t)nparts * 3 * sizeof(double);
    char* ws = static_cast<char*>(scratch(kScratchGae, bytes));
    w.stats_d = reinterpret_cast<double*>(ws + 256);
    w.stats_f = reinterpret_cast<float*>(ws + 256 + 64);
    GaeDesc* desc = reinterpret_cast<GaeDesc*>(ws + 512);
    float4* wstats = reinterpret_cast<float4*>(ws + 512 + desc_bytes);
    w.nchunks = nchunks;
    w.wstats = wstats;
    const unsigned epoch = ++g_gae_epoch;
    if ((epoch & 0x3FFFFFFFu) == 0) { g_gae_epoch = 1; }
    const float gl = gamma * lambda;  // formed first in float, src/ppo.cu:346
    const uintptr_t al = (uintptr_t)reward | (uintptr_t)v | (uintptr_t)v_next | (uintptr_t)advantage | (uintptr_t)adv_target;
    const uintptr_t alb = (uintptr_t)terminated | (uintptr_t)truncated;
    const int vec_ok = ((al & 15) == 0 && (alb & 3) == 0) ? 1 : 0;
    // persistent grid: never more CTAs than can be co-resident (the look-back relies on it)
#define B200_GAE_LAUNCH(TILES, HINT, MINB)                                                                         \
    do {                                                                                                           \
        int per_sm = 0;                                                                                            \
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gae_scan_kernel<TILES, HINT, MINB>, kGaeWarps * 32, 0)); \
        const int grid = std::min(nblocks, std::max(1, per_sm) * num_sms());                                       \
        B200_LAUNCH((gae_scan_kernel<TILES, HINT, MINB>), grid, kGaeWarps * 32, 0, reward, v, v_next,              \
                    reinterpret_cast<const unsigned char*>(terminated), reinterpret_cast<const unsigned char*>(truncated), \
                    n, gamma, gl, advantage, adv_target, desc, g_gae_epoch & 0x3FFFFFFFu, wstats, vec_ok);          \
    } while (0)
    static int tma_cfg = -1;
    const size_t tma_smem = (size_t)kGaeTmaWarps * kGaeStages * kGaeStageBytes + kGaeTmaWarps * kGaeStages * sizeof(uint64_t);
    if (tma_cfg < 0) {
        const char* e = getenv("PPO_B200_GAE_TMA");
        tma_cfg = (e && e[0] == '0') ? 0 : 1;
        if (tma_cfg) CUDA_CHECK(cudaFuncSetAttribute(gae_scan_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_smem));
    }
    if (tma_cfg && vec_ok && tiles == 4 && (n % 512) == 0 && n >= 512 * kGaeTmaWarps * 4) {
        const int grid = std::min(div_up(nchunks, kGaeTmaWarps), num_sms());            // 224 KB of shared memory: one CTA per SM, all co-resident
        B200_LAUNCH(gae_scan_tma_kernel, grid, kGaeTmaWarps * 32, tma_smem, reward, v, v_next,
                    reinterpret_cast<const unsigned char*>(terminated), reinterpret_cast<const unsigned char*>(truncated),
                    n, gamma, gl, advantage, adv_target, desc, g_gae_epoch & 0x3FFFFFFFu, wstats);
    } else if (variant == 0) B200_GAE_LAUNCH(4, true, 2);
    else if (variant == 1) B200_GAE_LAUNCH(2, true, 4);
    else if (variant == 2) B200_GAE_LAUNCH(4, false, 2);
    else if (variant == 3) B200_GAE_LAUNCH(2, false, 4);
    else B200_GAE_LAUNCH(4, true, 1);
#undef B200_GAE_LAUNCH
    double* part = reinterpret_cast<double*>(ws + 512 + desc_bytes + (size_t)nchunks * sizeof(float4));
    B200_LAUNCH(gae_stats_partial_kernel, nparts, 256, 0, wstats, nchunks, part);
    B200_LAUNCH(gae_stats_kernel, 1, 1024, 0, part, nparts, w.stats_d, w.stats_f);
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
