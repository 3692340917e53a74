// buffer.cu — trajectory buffer, permutation, minibatch gather.
//
// Same SoA layout and three pointer sets as reference src/trajectory_buffer.cu:41-94; the host set is
// pinned memory so the mirrors of src/trajectory_buffer.cu:227-273 run at PCIe speed on one stream.
//
// Gather (reference K5, src/trajectory_buffer.cu:168-185: one thread copies a whole row serially,
// uncoalesced): here one thread moves one float of the packed output row
// [state(S) | action(A) | logprob | advantage | adv_target], so every store is coalesced and the loads
// of one sample row are contiguous.  Payload is copied verbatim (bit-exact).
// Algorithmic traffic per sample: 4 B index + 2 * 4*(S + A + 3) B.
//
// Permutation: shuffle_buffer[_cuda] keep the reference's glibc rand() swap chain on the host
// (src/trajectory_buffer.cu:132-141; "integer work must be bit-exact") into pinned memory with an
// asynchronous upload.  ppo_b200_permutation is an additive device generator (keyed Feistel
// bijection with cycle walking) for runs that do not need the reference's index stream.
#include "common.cuh"
#include "internal.h"

namespace b200 {

// One warp per gathered row, lane = element of the packed row [state(S) | action(A) | logprob | advantage |
// adv_target]; U rows per warp iteration so that U independent index loads and then U independent payload loads
// are in flight per lane (the gather is latency-bound: bytes in flight per SM set its bandwidth).
constexpr int kGatherU = 8;
__global__ void __launch_bounds__(256)
gather_kernel(const int* __restrict__ idx, int offset, int limit, int batch_size, int S, int A,
              const float* __restrict__ state, const float* __restrict__ action,
              const float* __restrict__ logprob, const float* __restrict__ advantage,
              const float* __restrict__ adv_target, float* __restrict__ states, float* __restrict__ actions,
              float* __restrict__ logprobs, float* __restrict__ advantages, float* __restrict__ adv_targets) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int W = S + A + 3;
    for (int row0 = warp * kGatherU; row0 < batch_size; row0 += nwarps * kGatherU) {
        int src[kGatherU];
#pragma unroll
        for (int u = 0; u < kGatherU; u++) {
            const int r = row0 + u;
            src[u] = r < batch_size ? __ldg(idx + (offset + r) % limit) : -1;   // src/trajectory_buffer.cu:171-173
        }
        for (int c0 = 0; c0 < W; c0 += 32) {
            const int c = c0 + lane;
            const float* sp;
            float* dp;
            int stride, off;
            if (c < S) { sp = state; dp = states; stride = S; off = c; }
            else if (c < S + A) { sp = action; dp = actions; stride = A; off = c - S; }
            else if (c == S + A) { sp = logprob; dp = logprobs; stride = 1; off = 0; }
            else if (c == S + A + 1) { sp = advantage; dp = advantages; stride = 1; off = 0; }
            else { sp = adv_target; dp = adv_targets; stride = 1; off = 0; }
            float v[kGatherU];
#pragma unroll
            for (int u = 0; u < kGatherU; u++)
                if (c < W && src[u] >= 0) v[u] = sp[(size_t)src[u] * stride + off];
#pragma unroll
            for (int u = 0; u < kGatherU; u++)
                if (c < W && src[u] >= 0) dp[(size_t)(row0 + u) * stride + off] = v[u];
        }
    }
}

void launch_gather(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* state,
                   const float* action, const float* logprob, const float* advantage, const float* adv_target,
                   float* states, float* actions, float* logprobs, float* advantages, float* adv_targets) {
    if (batch_size <= 0) return;
    const int warps = div_up(batch_size, kGatherU);
    const int blocks = std::max(1, std::min(div_up(warps, 8), num_sms() * 8));
    B200_LAUNCH(gather_kernel, blocks, 256, 0, idx, offset, limit, batch_size, S, A, state, action, logprob,
                advantage, adv_target, states, actions, logprobs, advantages, adv_targets);
}

// ---- row-packed mirror for the layer-wise update path ------------------------------------------------------------------
// A gather from the SoA buffer touches five arrays per sample: a 68-byte state row straddles 3-4 32-byte DRAM sectors, a
// 24-byte action row 1-2, each of the three scalars pulls a whole sector for 4 bytes -- ~250 bytes fetched for 104 wanted
// (ncu: gather_kernel at 35 % of HBM peak with the memory pipe saturated).  The update phase gathers every row once per epoch,
// 14 epochs per iteration, from arrays that do not change after GAE, so ONE streaming pass builds a row-packed mirror
// [row][PW] = state | action | logprob | advantage | adv_target | pad, PW = the row rounded up to whole sectors (8 floats), and
// every epoch's gather then reads exactly PW * 4 contiguous, sector-aligned bytes per sample.
int packed_row_floats(int S, int A) { return ((S + A + 3) + 7) & ~7; }

__global__ void __launch_bounds__(256)
pack_rows_kernel(float* __restrict__ packed, long long rows, int S, int A, int PW, const float* __restrict__ state, const float* __restrict__ action,
                 const float* __restrict__ logprob, const float* __restrict__ advantage, const float* __restrict__ adv_target) {
    const long long total = rows * PW, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long r = e / PW;
        const int c = (int)(e - r * PW);
        float v = 0.f;
        if (c < S) v = state[r * S + c];
        else if (c < S + A) v = action[r * A + (c - S)];
        else if (c == S + A) v = logprob[r];
        else if (c == S + A + 1) v = advantage[r];
        else if (c == S + A + 2) v = adv_target[r];
        packed[e] = v;
    }
}

// A packed row is PW / 4 float4 words: that many lanes take one row, a warp 32 / (PW / 4) rows per load instruction and U
// independent loads per lane.  The rows land in a per-warp shared-memory tile; from there every output array is written as
// ONE dense run of consecutive floats (the RW output rows of a warp iteration are adjacent in each of the reference's five
// arrays), so each store instruction fills whole sectors -- scattering 4-byte pieces straight from the loaded float4 words
// cost four partial-sector writes per sector and held the kernel at 41 % of HBM peak.  Bit-exact copy.
constexpr int kPackedU = 4;
__global__ void __launch_bounds__(256)
gather_packed_kernel(const int* __restrict__ idx, int offset, int limit, int batch_size, int S, int A, int PW, const float* __restrict__ packed,
                     float* __restrict__ states, float* __restrict__ actions, float* __restrict__ logprobs, float* __restrict__ advantages,
                     float* __restrict__ adv_targets) {
    extern __shared__ __align__(16) float gsm[];   // [warps][RW][PW]
    const int lpr = PW >> 2;                       // lanes per row: 2, 4, 8, ... (PW is a multiple of 8)
    const int lane = threadIdx.x & 31;
    const int rpw = 32 / lpr;                      // rows per warp per load
    const int RW = rpw * kPackedU;                 // rows per warp iteration
    const int li = lane % lpr, lr = lane / lpr;
    float* tile = gsm + (size_t)(threadIdx.x >> 5) * RW * PW;
    const int s_dr = 32 / S, s_dc = 32 % S, a_dr = 32 / A, a_dc = 32 % A;      // 32 elements further = this many rows + columns
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int row0 = warp * RW; row0 < batch_size; row0 += n_warps * RW) {
        const int nrows = min(RW, batch_size - row0);
        int src[kPackedU];
#pragma unroll
        for (int u = 0; u < kPackedU; u++) {
            const int r = u * rpw + lr;
            src[u] = r < nrows ? __ldg(idx + (offset + row0 + r) % limit) : -1;      // src/trajectory_buffer.cu:171-173
        }
        float4 v[kPackedU];
#pragma unroll
        for (int u = 0; u < kPackedU; u++)
            if (src[u] >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(packed + (size_t)src[u] * PW) + li);
        __syncwarp();                              // the previous iteration's reads of the tile are done
#pragma unroll
        for (int u = 0; u < kPackedU; u++)
            if (src[u] >= 0) *reinterpret_cast<float4*>(tile + (u * rpw + lr) * PW + 4 * li) = v[u];
        __syncwarp();
        // (row, column) of element e advance without a division: + 32 columns, then carry whole rows
        float* so = states + (size_t)row0 * S;
        for (int e = lane, r = lane / S, c = lane - (lane / S) * S; e < nrows * S; e += 32) {
            so[e] = tile[r * PW + c];
            r += s_dr; c += s_dc;
            if (c >= S) { c -= S; r++; }
        }
        float* ao = actions + (size_t)row0 * A;
        for (int e = lane, r = lane / A, c = lane - (lane / A) * A; e < nrows * A; e += 32) {
            ao[e] = tile[r * PW + S + c];
            r += a_dr; c += a_dc;
            if (c >= A) { c -= A; r++; }
        }
        for (int r = lane; r < nrows; r += 32) {
            const float* t3 = tile + r * PW + S + A;
            logprobs[row0 + r] = t3[0];
            advantages[row0 + r] = t3[1];
            adv_targets[row0 + r] = t3[2];
        }
    }
}

void launch_pack_rows(float* packed, long long rows, int S, int A, const float* state, const float* action, const float* logprob,
                      const float* advantage, const float* adv_target) {
    if (rows <= 0) return;
    const int PW = packed_row_floats(S, A);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((rows * PW + 255) / 256, (long long)num_sms() * 16));
    B200_LAUNCH(pack_rows_kernel, blocks, 256, 0, packed, rows, S, A, PW, state, action, logprob, advantage, adv_target);
}

void launch_gather_packed(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* packed, float* states, float* actions,
                          float* logprobs, float* advantages, float* adv_targets) {
    if (batch_size <= 0) return;
    const int PW = packed_row_floats(S, A);
    if (PW > 128) B200_FATAL("packed gather: rows of %d floats are not supported (S + A + 3 <= 128)", PW);
    const int rows_per_warp = (32 / (PW >> 2)) * kPackedU;
    const size_t smem = (size_t)8 * rows_per_warp * PW * sizeof(float);        // 8 warps x 512 floats = 16 KB whatever PW is
    const int blocks = std::max(1, std::min(div_up(div_up(batch_size, rows_per_warp), 8), num_sms() * 8));
    B200_LAUNCH(gather_packed_kernel, blocks, 256, smem, idx, offset, limit, batch_size, S, A, PW, packed, states, actions, logprobs, advantages,
                adv_targets);
}

void host_shuffle(int* idx, int limit) {
    for (int i = 0; i < limit; i++) idx[i] = i;
    for (int i = 0; i < limit; i++) {
        const int j = rand() % limit;
        const int t = idx[i];
        idx[i] = idx[j];
        idx[j] = t;
    }
}

// ---- device permutation: 4-round Feistel network on ceil(log2 n) bits + cycle walking -------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256)
permutation_kernel(int* __restrict__ idx, int n, int half_bits, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3) {
    const uint32_t mask = (1u << half_bits) - 1u;
    const uint32_t keys[4] = {k0, k1, k2, k3};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t x = (uint32_t)i;
        do {   // bijection on [0, 2^(2*half_bits)); walk the cycle until we land back inside [0, n)
            uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
            for (int round = 0; round < 4; round++) {
                const uint32_t f = mix32(r ^ keys[round]) & mask;
                const uint32_t nl = r;
                r = l ^ f;
                l = nl;
            }
            x = (l << half_bits) | r;
        } while (x >= (uint32_t)n);
        idx[i] = (int)x;
    }
}

static uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// ---- accessors (src/trajectory_buffer.cu:5-39): follow the ACTIVE pointer set --------------------
static float* get_action(TrajectoryBuffer* b, int i) { return b->action_p + (size_t)i * b->action_size; }
static float* get_state(TrajectoryBuffer* b, int i) { return b->state_p + (size_t)i * b->state_size; }
static float* get_next_state(TrajectoryBuffer* b, int i) { return b->next_state_p + (size_t)i * b->state_size; }
static float* get_reward(TrajectoryBuffer* b, int i) { return b->reward_p + i; }
static float* get_logprob(TrajectoryBuffer* b, int i) { return b->logprob_p + i; }
static float* get_advantage(TrajectoryBuffer* b, int i) { return b->advantage_p + i; }
static float* get_adv_target(TrajectoryBuffer* b, int i) { return b->adv_target_p + i; }
static bool* get_terminated(TrajectoryBuffer* b, int i) { return b->terminated_p + i; }
static bool* get_truncated(TrajectoryBuffer* b, int i) { return b->truncated_p + i; }

static void activate(TrajectoryBuffer* b, bool device) {
    b->action_p = device ? b->d_action_p : b->h_action_p;
    b->state_p = device ? b->d_state_p : b->h_state_p;
    b->next_state_p = device ? b->d_next_state_p : b->h_next_state_p;
    b->reward_p = device ? b->d_reward_p : b->h_reward_p;
    b->logprob_p = device ? b->d_logprob_p : b->h_logprob_p;
    b->advantage_p = device ? b->d_advantage_p : b->h_advantage_p;
    b->adv_target_p = device ? b->d_adv_target_p : b->h_adv_target_p;
    b->terminated_p = device ? b->d_terminated_p : b->h_terminated_p;
    b->truncated_p = device ? b->d_truncated_p : b->h_truncated_p;
}

// shuffle staging: pinned host permutation + device copy, reused across calls
struct ShuffleStage { int* h = nullptr; int cap = 0; bool idx_is_device = false; };
static ShuffleStage g_shuffle;

}  // namespace b200

using namespace b200;

extern "C" {

void ppo_b200_gather(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* state,
                     const float* action, const float* logprob, const float* advantage, const float* adv_target,
                     float* states, float* actions, float* logprobs, float* advantages, float* adv_targets) {
    launch_gather(idx, offset, limit, batch_size, S, A, state, action, logprob, advantage, adv_target, states,
                  actions, logprobs, advantages, adv_targets);
}

int ppo_b200_packed_row_floats(int S, int A) { return packed_row_floats(S, A); }

void ppo_b200_pack_rows(float* packed, long long rows, int S, int A, const float* state, const float* action, const float* logprob,
                        const float* advantage, const float* adv_target) {
    launch_pack_rows(packed, rows, S, A, state, action, logprob, advantage, adv_target);
}

void ppo_b200_gather_packed(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* packed, float* states,
                            float* actions, float* logprobs, float* advantages, float* adv_targets) {
    launch_gather_packed(idx, offset, limit, batch_size, S, A, packed, states, actions, logprobs, advantages, adv_targets);
}

void ppo_b200_permutation(int* idx, int n, unsigned long long seed, unsigned long long epoch) {
    if (n <= 0) return;
    int bits = 2;
    while ((1ll << bits) < n) bits++;
    if (bits & 1) bits++;
    uint64_t s = seed * 0x9e3779b97f4a7c15ull + epoch + 1;
    const uint64_t a = splitmix64(s), b = splitmix64(s);
    const int blocks = (int)std::min<long long>(div_up(n, 256), (long long)num_sms() * 8);
    B200_LAUNCH(permutation_kernel, blocks, 256, 0, idx, n, bits / 2, (uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b,
                (uint32_t)(b >> 32));
}

TrajectoryBuffer* create_trajectory_buffer(int capacity, int state_size, int action_size) {
    ensure_device();
    TrajectoryBuffer* b = (TrajectoryBuffer*)malloc(sizeof(TrajectoryBuffer));
    b->capacity = capacity;
    b->idx = 0;
    b->state_size = state_size;
    b->action_size = action_size;
    b->full = false;
    const size_t n = capacity;
    b->h_action_p = hmalloc_pinned<float>(n * action_size);
    b->h_state_p = hmalloc_pinned<float>(n * state_size);
    b->h_next_state_p = hmalloc_pinned<float>(n * state_size);
    b->h_reward_p = hmalloc_pinned<float>(n);
    b->h_logprob_p = hmalloc_pinned<float>(n);
    b->h_advantage_p = hmalloc_pinned<float>(n);
    b->h_adv_target_p = hmalloc_pinned<float>(n);
    b->h_terminated_p = hmalloc_pinned<bool>(n);
    b->h_truncated_p = hmalloc_pinned<bool>(n);
    b->d_action_p = dmalloc<float>(n * action_size);
    b->d_state_p = dmalloc<float>(n * state_size);
    b->d_next_state_p = dmalloc<float>(n * state_size);
    b->d_reward_p = dmalloc<float>(n);
    b->d_logprob_p = dmalloc<float>(n);
    b->d_advantage_p = dmalloc<float>(n);
    b->d_adv_target_p = dmalloc<float>(n);
    b->d_terminated_p = dmalloc<bool>(n);
    b->d_truncated_p = dmalloc<bool>(n);
    memset(b->h_advantage_p, 0, n * sizeof(float));
    memset(b->h_adv_target_p, 0, n * sizeof(float));
    activate(b, false);
    b->random_idx = nullptr;
    b->action = get_action;
    b->state = get_state;
    b->next_state = get_next_state;
    b->reward = get_reward;
    b->logprob = get_logprob;
    b->advantage = get_advantage;
    b->adv_target = get_adv_target;
    b->terminated = get_terminated;
    b->truncated = get_truncated;
    return b;
}

void free_trajectory_buffer(TrajectoryBuffer* b, bool use_cuda) {
    if (!b) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    float* hf[] = {b->h_action_p, b->h_state_p, b->h_next_state_p, b->h_reward_p, b->h_logprob_p, b->h_advantage_p, b->h_adv_target_p};
    for (float* p : hf) CUDA_CHECK(cudaFreeHost(p));
    CUDA_CHECK(cudaFreeHost(b->h_terminated_p));
    CUDA_CHECK(cudaFreeHost(b->h_truncated_p));
    float* df[] = {b->d_action_p, b->d_state_p, b->d_next_state_p, b->d_reward_p, b->d_logprob_p, b->d_advantage_p, b->d_adv_target_p};
    for (float* p : df) CUDA_CHECK(cudaFree(p));
    CUDA_CHECK(cudaFree(b->d_terminated_p));
    CUDA_CHECK(cudaFree(b->d_truncated_p));
    if (b->random_idx) {   // src/trajectory_buffer.cu:108-112: device array after shuffle_buffer_cuda, host array otherwise
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, b->random_idx) == cudaSuccess && attr.type == cudaMemoryTypeDevice) CUDA_CHECK(cudaFree(b->random_idx));
        else { cudaGetLastError(); free(b->random_idx); }
    }
    (void)use_cuda;
    free(b);
}

void shuffle_buffer(TrajectoryBuffer* b) {     // src/trajectory_buffer.cu:126-142 (host index array)
    const int limit = b->full ? b->capacity : b->idx;
    if (b->random_idx) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, b->random_idx) == cudaSuccess && attr.type == cudaMemoryTypeDevice) CUDA_CHECK(cudaFree(b->random_idx));
        else { cudaGetLastError(); free(b->random_idx); }
    }
    b->random_idx = (int*)malloc((size_t)limit * sizeof(int));
    host_shuffle(b->random_idx, limit);
}

void shuffle_buffer_cuda(TrajectoryBuffer* b) { // src/trajectory_buffer.cu:144-166 (device index array)
    const int limit = b->full ? b->capacity : b->idx;
    if (limit > g_shuffle.cap) {
        CUDA_CHECK(cudaStreamSynchronize(stream()));
        if (g_shuffle.h) CUDA_CHECK(cudaFreeHost(g_shuffle.h));
        g_shuffle.h = hmalloc_pinned<int>(limit);
        g_shuffle.cap = limit;
    }
    CUDA_CHECK(cudaStreamSynchronize(stream()));   // previous upload from the pinned stage must be done
    host_shuffle(g_shuffle.h, limit);
    if (b->random_idx) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, b->random_idx) == cudaSuccess && attr.type == cudaMemoryTypeDevice) CUDA_CHECK(cudaFree(b->random_idx));
        else { cudaGetLastError(); free(b->random_idx); }
    }
    b->random_idx = dmalloc<int>(limit);
    CUDA_CHECK(cudaMemcpyAsync(b->random_idx, g_shuffle.h, (size_t)limit * sizeof(int), cudaMemcpyHostToDevice, stream()));
}

void get_batch_cuda(TrajectoryBuffer* b, int batch_idx, int batch_size, float* states, float* actions,
                    float* logprobs, float* advantages, float* adv_targets) {
    const int limit = b->full ? b->capacity : b->idx;
    launch_gather(b->random_idx, batch_idx * batch_size, limit, batch_size, b->state_size, b->action_size, b->state_p,
                  b->action_p, b->logprob_p, b->advantage_p, b->adv_target_p, states, actions, logprobs, advantages,
                  adv_targets);
}

// host-pointer twin (src/trajectory_buffer.cu:202-220): the HOST arrays + host index list are staged
// to the device, gathered by the kernel, and the minibatch copied back.
void get_batch(TrajectoryBuffer* b, int batch_idx, int batch_size, float* states, float* actions,
               float* logprobs, float* advantages, float* adv_targets) {
    const int limit = b->full ? b->capacity : b->idx;
    const int S = b->state_size, A = b->action_size;
    HostStage st;
    st.add(b->random_idx, (size_t)limit * 4, true, false);
    st.add(b->state_p, (size_t)limit * S * 4, true, false);
    st.add(b->action_p, (size_t)limit * A * 4, true, false);
    st.add(b->logprob_p, (size_t)limit * 4, true, false);
    st.add(b->advantage_p, (size_t)limit * 4, true, false);
    st.add(b->adv_target_p, (size_t)limit * 4, true, false);
    st.add(states, (size_t)batch_size * S * 4, false, true);
    st.add(actions, (size_t)batch_size * A * 4, false, true);
    st.add(logprobs, (size_t)batch_size * 4, false, true);
    st.add(advantages, (size_t)batch_size * 4, false, true);
    st.add(adv_targets, (size_t)batch_size * 4, false, true);
    st.upload();
    launch_gather(st.dev<int>(0), batch_idx * batch_size, limit, batch_size, S, A, st.dev<float>(1), st.dev<float>(2),
                  st.dev<float>(3), st.dev<float>(4), st.dev<float>(5), st.dev<float>(6), st.dev<float>(7),
                  st.dev<float>(8), st.dev<float>(9), st.dev<float>(10));
    st.download();
}

void reset_buffer(TrajectoryBuffer* b) {        // src/trajectory_buffer.cu:222-225
    b->idx = 0;
    b->full = false;
}

static void copy_all(TrajectoryBuffer* b, bool to_device) {
    const size_t n = b->capacity, S = b->state_size, A = b->action_size;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    auto cp = [&](void* d, void* h, size_t bytes) {
        CUDA_CHECK(cudaMemcpyAsync(to_device ? d : h, to_device ? h : d, bytes, kind, stream()));
    };
    cp(b->d_action_p, b->h_action_p, n * A * 4);
    cp(b->d_state_p, b->h_state_p, n * S * 4);
    cp(b->d_next_state_p, b->h_next_state_p, n * S * 4);
    cp(b->d_reward_p, b->h_reward_p, n * 4);
    cp(b->d_logprob_p, b->h_logprob_p, n * 4);
    cp(b->d_advantage_p, b->h_advantage_p, n * 4);
    cp(b->d_adv_target_p, b->h_adv_target_p, n * 4);
    cp(b->d_terminated_p, b->h_terminated_p, n);
    cp(b->d_truncated_p, b->h_truncated_p, n);
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

}  // extern "C"

namespace b200 {
// Update phase on a host-filled buffer (ppo_b200_update): GAE overwrites advantage / adv_target and nothing else in the
// buffer changes, so the mirrors move only what is needed - the seven input arrays up, the two output arrays down -
// instead of the reference's nine blocking copies each way (src/trajectory_buffer.cu:227-273).  The post-condition is the
// reference's: on return the host set is active and holds exactly what the device holds.
void buffer_upload_inputs(TrajectoryBuffer* b) {
    const size_t n = b->capacity, S = b->state_size, A = b->action_size;
    auto up = [&](void* d, const void* h, size_t bytes) { CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream())); };
    up(b->d_state_p, b->h_state_p, n * S * 4);
    up(b->d_next_state_p, b->h_next_state_p, n * S * 4);
    up(b->d_reward_p, b->h_reward_p, n * 4);
    up(b->d_terminated_p, b->h_terminated_p, n);
    up(b->d_truncated_p, b->h_truncated_p, n);
    up(b->d_action_p, b->h_action_p, n * A * 4);
    up(b->d_logprob_p, b->h_logprob_p, n * 4);
    activate(b, true);
}
void buffer_download_outputs(TrajectoryBuffer* b) {
    const size_t n = b->capacity;
    CUDA_CHECK(cudaMemcpyAsync(b->h_advantage_p, b->d_advantage_p, n * 4, cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaMemcpyAsync(b->h_adv_target_p, b->d_adv_target_p, n * 4, cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    activate(b, false);
}
}  // namespace b200

extern "C" {

void buffer_to_device(TrajectoryBuffer* b) {    // src/trajectory_buffer.cu:227-249
    copy_all(b, true);
    activate(b, true);
}

void buffer_to_host(TrajectoryBuffer* b) {      // src/trajectory_buffer.cu:251-273
    copy_all(b, false);
    activate(b, false);
}

}  // extern "C"
