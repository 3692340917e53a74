// env.cu — environments and rollout kernels.
//
//  * create_simple_env      : the reference's toy env (src/env.c) behind the same hooks.
//  * create_pendulum_env    : native host Pendulum-v1 behind the reference hooks.  The reference
//                             reaches Pendulum through embedded CPython + gymnasium (src/gym_env.c,
//                             scripts/gym_env.py:12); the dynamics follow the public gymnasium
//                             definition: parity with gymnasium itself is UNPINNED
//                             (not installable here), pinned only by hand-derived known answers.
//  * create_pendulum_env_cuda: n vectorised Pendulum envs resident on the device, stepped by
//      rollout_kernel — policy forward (weights staged once in shared memory), Box-Muller sampling
//      from a counter-based Philox stream, log-prob, env step, and the env-major buffer write for
//      all T steps in ONE launch (the reference does 1 env, 1 step at a time on the host with a
//      Python call per step, src/ppo.cu:54-79).
//  * sample_action_kernel   : the per-step policy sample for opaque host envs; consumes the two
//      glibc rand() draws the host made, so the RNG stream is the reference's (src/policy.cu:46-89).
#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr double kPiD = 3.14159265358979323846;

// ================================ Pendulum dynamics (host + device) ===============================
struct PendulumState { double th, thd; };

__host__ __device__ inline double angle_normalize(double x) {
    double y = fmod(x + kPiD, 2 * kPiD);
    if (y < 0) y += 2 * kPiD;
    return y - kPiD;
}

__host__ __device__ inline void pendulum_obs(const PendulumState& s, float* obs) {
    obs[0] = (float)cos(s.th);
    obs[1] = (float)sin(s.th);
    obs[2] = (float)s.thd;
}

// gymnasium PendulumEnv.step: g=10, m=1, l=1, dt=0.05, max_speed=8, max_torque=2
__host__ __device__ inline float pendulum_step(PendulumState& s, float action) {
    double u = action;
    u = u > 2.0 ? 2.0 : (u < -2.0 ? -2.0 : u);
    const double an = angle_normalize(s.th);
    const double cost = an * an + 0.1 * s.thd * s.thd + 0.001 * u * u;
    double nthd = s.thd + (3 * 10.0 / (2 * 1.0) * sin(s.th) + 3.0 / (1.0 * 1.0 * 1.0) * u) * 0.05;
    nthd = nthd > 8.0 ? 8.0 : (nthd < -8.0 ? -8.0 : nthd);
    s.th = s.th + nthd * 0.05;
    s.thd = nthd;
    return (float)(-cost);
}

// ================================ Philox4x32-10 ======================================================
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

// ================================ cooperative small-MLP forward =======================================
// E envs (lanes) x units split over the warps of the CTA.  hin/hout are k-major [width][E] in shared
// memory; W/b point to the layer's parameters (shared or global).
template <int E>
__device__ __forceinline__ void layer_forward_tile(const float* __restrict__ W, const float* __restrict__ b,
                                                   const float* hin, float* hout, int n, int l, int act,
                                                   int warp, int nwarps, int lane) {
    for (int j0 = warp * 4; j0 < l; j0 += nwarps * 4) {
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = (j0 + u < l) ? b[j0 + u] : 0.f;
        for (int k = 0; k < n; k++) {
            const float hv = hin[k * E + lane];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (j0 + u < l) acc[u] = fmaf(W[(size_t)(j0 + u) * n + k], hv, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (j0 + u < l) hout[(j0 + u) * E + lane] = act_apply(acc[u], act);
    }
}

struct NetView {
    const float* params;   // flat W0,b0,W1,b1...
    int num_layers;        // number of sizes
    int sizes[8];
    int acts[8];
    int param_count;
    int max_width;
};

static NetView make_view(NeuralNetwork* nn) {
    NetDev* nd = net_dev(nn);
    if (nn->num_layers > 8) B200_FATAL("rollout kernels support at most 7 weight layers");
    NetView v{};
    v.params = nd->params;
    v.num_layers = nn->num_layers;
    v.param_count = (int)nd->param_count;
    v.max_width = 0;
    for (int i = 0; i < nn->num_layers; i++) {
        v.sizes[i] = nd->sizes[i];
        v.max_width = std::max(v.max_width, nd->sizes[i]);
        if (i < nn->num_layers - 1) v.acts[i] = nd->acts[i];
    }
    return v;
}

// log-prob of one action row, same arithmetic as policy.cu:log_prob_row (src/policy.cu:67-74)
__device__ __forceinline__ float log_prob_dev(const float* mu, const float* log_std, const float* action, int A) {
    float logprob = (float)(-0.5 * A * (double)logf((float)(2 * kPiD)));
    for (int j = 0; j < A; j++) {
        const float z = __fdiv_rn(__fsub_rn(action[j], mu[j]), expf(log_std[j]));
        logprob = (float)((double)logprob - ((double)log_std[j] + 0.5 * (double)__fmul_rn(z, z)));
    }
    return logprob;
}

// ================================ device env + fused rollout ===========================================
struct DeviceEnv {
    Env base;                   // MUST be first: callers hold Env*
    unsigned long long magic;
    int n_envs;
    unsigned long long seed;
    unsigned long long rollouts;   // number of rollouts so far (RNG counter)
    PendulumState host_state;      // env 0 driven through the hooks
    int host_steps;
    float* d_ret_sum;              // [n_envs] sum of rewards of finished episodes
    int* d_ret_cnt;                // [n_envs]
    // running observation normalisation (new capability; merge formula of include/welford_var.h:33-40,58-66)
    bool obs_norm;
    double* d_obs_stat;            // [3 features][mean, M2, n] running Welford state (float64)
    float* d_obs_mean;             // [3]
    float* d_obs_inv_std;          // [3]  1 / (std + 1e-8)
    float* d_obs_partial;          // [CTAs][3][mean, M2, n] per-CTA Welford triples of the last rollout
    int obs_partial_cap;
};
constexpr unsigned long long kDeviceEnvMagic = 0xB200E17Full;
static DeviceEnv* g_device_env = nullptr;

DeviceEnv* as_device_env(Env* env) {
    DeviceEnv* e = reinterpret_cast<DeviceEnv*>(env);
    return (g_device_env == e && e && e->magic == kDeviceEnvMagic) ? e : nullptr;
}
int device_env_count(DeviceEnv* e) { return e->n_envs; }
void device_env_reset_obs_norm(DeviceEnv* e) {
    const float one[3] = {1.f, 1.f, 1.f};
    CUDA_CHECK(cudaMemsetAsync(e->d_obs_stat, 0, 9 * sizeof(double), stream()));
    CUDA_CHECK(cudaMemsetAsync(e->d_obs_mean, 0, 3 * sizeof(float), stream()));
    CUDA_CHECK(cudaMemcpyAsync(e->d_obs_inv_std, one, sizeof(one), cudaMemcpyHostToDevice, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void device_env_set_obs_norm(bool enabled) {
    if (!g_device_env) B200_FATAL("ppo_b200_set_obs_norm: create the device env first (create_pendulum_env_cuda)");
    if (enabled && !g_device_env->obs_norm) device_env_reset_obs_norm(g_device_env);
    g_device_env->obs_norm = enabled;
}


constexpr int kRollE = 32;        // envs per CTA (one per lane)
constexpr int kRollWarps = 4;

struct RolloutArgs {
    NetView net;
    const float* log_std;
    int n_envs, T, S, A;
    unsigned long long seed, rollout;
    float *state, *next_state, *action, *reward, *logprob;
    unsigned char *terminated, *truncated;
    const float* obs_mean;
    const float* obs_inv_std;
    float* ret_sum;
    int* ret_cnt;
    float* raw_obs_partial;   // [blocks][S][2] sum / sum of squares of RAW observations (obs-norm), may be null
    int weights_in_smem;
    int horizon;
};

__global__ void __launch_bounds__(kRollE * kRollWarps)
rollout_kernel(const RolloutArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* hA = smem;                                   // [max_width][E]
    float* hB = hA + p.net.max_width * kRollE;
    float* wsm = hB + p.net.max_width * kRollE;         // parameters (optional)
    const float* params = p.net.params;
    if (p.weights_in_smem) {
        for (int i = threadIdx.x; i < p.net.param_count; i += blockDim.x) wsm[i] = p.net.params[i];
        params = wsm;
    }
    __syncthreads();

    const int env = blockIdx.x * kRollE + lane;
    const bool live = env < p.n_envs;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    PendulumState st{0.0, 0.0};
    int ep_steps = 0, episode = 0;
    float ep_ret = 0.f, ret_sum = 0.f;
    int ret_cnt = 0;
    float obs[3];
    float osum[3] = {0.f, 0.f, 0.f}, osq[3] = {0.f, 0.f, 0.f};

    auto reset = [&]() {   // theta ~ U(-pi,pi), theta_dot ~ U(-1,1)  (gymnasium reset)
        const uint4 r = philox4x32(make_uint4((uint32_t)env, (uint32_t)episode, (uint32_t)p.rollout, 0xFFFFFFFFu), key);
        st.th = (2.0 * (double)u01_open(r.x) - 1.0) * kPiD;
        st.thd = 2.0 * (double)u01_open(r.y) - 1.0;
        ep_steps = 0;
        ep_ret = 0.f;
        pendulum_obs(st, obs);
    };
    auto norm = [&](int k, float x) { return p.obs_mean ? (x - p.obs_mean[k]) * p.obs_inv_std[k] : x; };
    if (warp == 0 && live) reset();   // src/ppo.cu:55: every collect starts with a reset

    for (int t = 0; t < p.T; t++) {
        const size_t row = (size_t)env * p.T + t;   // env-major flat index
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float x = live ? norm(k, obs[k]) : 0.f;
                hA[k * kRollE + lane] = x;
                if (live) { p.state[row * 3 + k] = x; osum[k] += obs[k]; osq[k] += obs[k] * obs[k]; }
            }
        }
        __syncthreads();
        // ---- policy forward: hA -> ... -> mu
        float* hin = hA;
        float* hout = hB;
        const float* w = params;
        for (int i = 0; i < p.net.num_layers - 1; i++) {
            const int n = p.net.sizes[i], l = p.net.sizes[i + 1];
            layer_forward_tile<kRollE>(w, w + (size_t)n * l, hin, hout, n, l, p.net.acts[i], warp, kRollWarps, lane);
            w += (size_t)n * l + l;
            __syncthreads();
            float* tmp = hin; hin = hout; hout = tmp;
        }
        // ---- sample, log-prob, env step (A == 1 for Pendulum), buffer write
        if (warp == 0 && live) {
            const float mu = hin[lane];
            const uint4 r = philox4x32(make_uint4((uint32_t)env, (uint32_t)t, (uint32_t)p.rollout, 0u), key);
            const float z = sqrtf(-2.f * logf(u01_open(r.x))) * cosf(6.283185307179586f * u01_open(r.y));
            const float ls = p.log_std[0];
            const float a = mu + z * expf(ls);                       // src/policy.cu:85
            const float lp = log_prob_dev(&mu, &ls, &a, 1);
            const float rew = pendulum_step(st, a);
            ep_steps++;
            ep_ret += rew;
            bool trunc = ep_steps >= p.horizon;                       // TimeLimit(200)
            pendulum_obs(st, obs);
            p.action[row] = a;
            p.logprob[row] = lp;
            p.reward[row] = rew;
#pragma unroll
            for (int k = 0; k < 3; k++) p.next_state[row * 3 + k] = norm(k, obs[k]);
            if (trunc) { ret_sum += ep_ret; ret_cnt++; episode++; }
            if (t == p.T - 1) trunc = true;                           // src/ppo.cu:70-74
            p.terminated[row] = 0;
            p.truncated[row] = trunc ? 1 : 0;
            if (trunc && t < p.T - 1) reset();                        // src/ppo.cu:64-66
        }
        __syncthreads();
    }
    if (warp == 0) {
        if (live) { p.ret_sum[env] = ret_sum; p.ret_cnt[env] = ret_cnt; }
        if (p.raw_obs_partial) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float s = warp_sum(live ? osum[k] : 0.f), q = warp_sum(live ? osq[k] : 0.f);
                if (lane == 0) { p.raw_obs_partial[(blockIdx.x * 3 + k) * 2] = s; p.raw_obs_partial[(blockIdx.x * 3 + k) * 2 + 1] = q; }
            }
        }
    }
}


// ================================ 64-wide fused rollout ===============================================
// Same contract as rollout_kernel, for nets the 64-wide tile kernels support (all widths <= 64): the weight
// image [Wt_l | biases] maintained by the Adam kernel is staged once per CTA; per step each thread owns a
// 4 envs x 4 units register tile of the hidden layers (2 LDS.128 per 16 FFMAs instead of 1 scalar load per
// FFMA), the action-mean layer is split over the four warps, and warp 0 (lane = env) samples, steps the
// env in float64 and writes the env-major buffer row.  sincos(theta) is computed once per step and shared
// by the observation and the next step's dynamics.
constexpr int kR64E = 32;
constexpr int kR64Threads = 128;

struct Rollout64Args {
    FusedNet net;
    const float* image;
    const float* log_std;
    int n_envs, T;
    unsigned long long seed, rollout;
    float *state, *next_state, *action, *reward, *logprob;
    unsigned char *terminated, *truncated;
    const float* obs_mean;        // null: no observation normalisation
    const float* obs_inv_std;
    float* ret_sum;
    int* ret_cnt;
    float* obs_partial;           // [blocks][3][3] Welford triples of the RAW observations, may be null
    int horizon;
};

__device__ __forceinline__ void welford_merge(float& mean, float& m2, float& n, float mb, float m2b, float nb) {
    if (nb <= 0.f) return;
    const float nn = n + nb, delta = mb - mean;          // include/welford_var.h:33-40
    mean += delta * nb / nn;
    m2 += m2b + delta * delta * n * nb / nn;
    n = nn;
}

__global__ void __launch_bounds__(kR64Threads) rollout64_kernel(const Rollout64Args p) {
    constexpr int E = kR64E;
    extern __shared__ __align__(16) float smem[];
    const FusedNet& net = p.net;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* img = smem;
    const int HW = net.max_width_pad;       // rows of the activation tiles (64 or 128)
    float* hA = img + net.img_floats;       // [HW][E]: feature-major, one column per env
    float* hB = hA + HW * E;
    float* part = hB + HW * E;              // [4][8][E] partial action means
    for (int i = tid * 4; i < net.img_floats; i += kR64Threads * 4)
        *reinterpret_cast<float4*>(img + i) = *reinterpret_cast<const float4*>(p.image + i);
    const int L = net.L, A = net.sizes[L];
    const int env = blockIdx.x * E + lane;
    const bool live = env < p.n_envs;
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    double th = 0.0, thd = 0.0, sn = 0.0, cs = 1.0;
    int ep_steps = 0, episode = 0, ret_cnt = 0;
    float ep_ret = 0.f, ret_sum = 0.f;
    float wmean[3] = {0.f, 0.f, 0.f}, wm2[3] = {0.f, 0.f, 0.f}, wn = 0.f;
    float om[3] = {0.f, 0.f, 0.f}, oi[3] = {1.f, 1.f, 1.f};
    if (p.obs_mean) {
#pragma unroll
        for (int k = 0; k < 3; k++) { om[k] = p.obs_mean[k]; oi[k] = p.obs_inv_std[k]; }
    }
    auto reset = [&]() {   // theta ~ U(-pi,pi), theta_dot ~ U(-1,1)  (gymnasium reset)
        const uint4 r = philox4x32(make_uint4((uint32_t)env, (uint32_t)episode, (uint32_t)p.rollout, 0xFFFFFFFFu), key);
        th = (2.0 * (double)u01_open(r.x) - 1.0) * kPiD;
        thd = 2.0 * (double)u01_open(r.y) - 1.0;
        sincos(th, &sn, &cs);
        ep_steps = 0;
        ep_ret = 0.f;
    };
    if (warp == 0 && live) reset();         // src/ppo.cu:55: every collect starts with a reset
    // zero the padding row of the input tile once (the first layer reads pad4(S) = 4 feature rows)
    if (tid < E) hA[3 * E + tid] = 0.f;
    __syncthreads();

    const int eg = tid & 7, ug = tid >> 3;  // 4 envs x 4 units per thread
    for (int t = 0; t < p.T; t++) {
        const size_t row = (size_t)env * p.T + t;   // env-major flat index
        if (warp == 0) {
            const float raw[3] = {(float)cs, (float)sn, (float)thd};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float x = live ? (raw[k] - om[k]) * oi[k] : 0.f;
                hA[k * E + lane] = x;
                if (live) p.state[row * 3 + k] = x;
            }
            if (live) {                     // Welford update with the raw observation
                wn += 1.f;
#pragma unroll
                for (int k = 0; k < 3; k++) { const float d = raw[k] - wmean[k]; wmean[k] += d / wn; wm2[k] += d * (raw[k] - wmean[k]); }
            }
        }
        __syncthreads();
        // ---- hidden layers: hin -> hout, 4x4 register tiles
        float* hin = hA;
        float* hout = hB;
        for (int l = 0; l < L - 1; l++) {
            const int n_in = net.sizes[l], n_out = net.sizes[l + 1];
            for (int ub = 0; ub < pad4(n_out); ub += 64)        // layers wider than 64 units: 64-unit blocks
            if (ub + 4 * ug < pad4(n_out)) {
                const float* wp = img + net.wt_off[l] + ub + 4 * ug;
                const int ldw = net.ldw[l];
                const float4 b = *reinterpret_cast<const float4*>(img + net.bs_off[l] + ub + 4 * ug);
                float2 acc2[4][2];               // [unit][env pair]: packed FFMA2
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int c = 0; c < 4; c++) { acc2[c][0] = make_float2(bv[c], bv[c]); acc2[c][1] = make_float2(bv[c], bv[c]); }
                const float* xp = hin + 4 * eg;
#pragma unroll 4
                for (int k = 0; k < n_in; k++) {
                    const float4 a = *reinterpret_cast<const float4*>(xp + k * E);
                    const float4 w = *reinterpret_cast<const float4*>(wp + k * ldw);
                    const float2 ap[2] = {make_float2(a.x, a.y), make_float2(a.z, a.w)};
                    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        acc2[c][0] = __ffma2_rn(ap[0], make_float2(wv[c], wv[c]), acc2[c][0]);
                        acc2[c][1] = __ffma2_rn(ap[1], make_float2(wv[c], wv[c]), acc2[c][1]);
                    }
                }
                float acc[4][4];
#pragma unroll
                for (int c = 0; c < 4; c++) { acc[c][0] = acc2[c][0].x; acc[c][1] = acc2[c][0].y; acc[c][2] = acc2[c][1].x; acc[c][3] = acc2[c][1].y; }
                const int act = net.acts[l];
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<float4*>(hout + (ub + 4 * ug + c) * E + 4 * eg) =
                        make_float4(act_apply(acc[c][0], act), act_apply(acc[c][1], act), act_apply(acc[c][2], act), act_apply(acc[c][3], act));
            }
            __syncthreads();
            float* tmp = hin; hin = hout; hout = tmp;
        }
        // ---- action-mean layer (A <= 8): each warp takes a quarter of k, lane = env
        {
            const int n_in = net.sizes[L - 1];
            const int kq = (n_in + 3) >> 2;
            const int k0 = min(n_in, warp * kq), k1 = min(n_in, k0 + kq);
            const float* wt = img + net.wt_off[L - 1];
            const int ldw = net.ldw[L - 1];
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = 0.f;
            for (int k = k0; k < k1; k++) {
                const float x = hin[k * E + lane];
#pragma unroll
                for (int j = 0; j < 8; j++) if (j < A) acc[j] = fmaf(x, wt[k * ldw + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) if (j < A) part[(warp * 8 + j) * E + lane] = acc[j];
        }
        __syncthreads();
        // ---- sample, log-prob, env step (A == 1 for Pendulum), buffer write
        if (warp == 0 && live) {
            float mu = ((part[lane] + part[8 * E + lane]) + part[16 * E + lane]) + part[24 * E + lane] + img[net.bs_off[L - 1]];
            mu = act_apply(mu, net.acts[L - 1]);
            const uint4 r = philox4x32(make_uint4((uint32_t)env, (uint32_t)t, (uint32_t)p.rollout, 0u), key);
            const float z = sqrtf(-2.f * logf(u01_open(r.x))) * cosf(6.283185307179586f * u01_open(r.y));
            const float ls = p.log_std[0];
            const float a = mu + z * expf(ls);                       // src/policy.cu:85
            const float lp = log_prob_dev(&mu, &ls, &a, 1);
            // gymnasium PendulumEnv.step (same arithmetic as pendulum_step above, sin(theta) reused from the observation)
            double u = a;
            u = u > 2.0 ? 2.0 : (u < -2.0 ? -2.0 : u);
            const double an = angle_normalize(th);
            const double cost = an * an + 0.1 * thd * thd + 0.001 * u * u;
            double nthd = thd + (3 * 10.0 / (2 * 1.0) * sn + 3.0 / (1.0 * 1.0 * 1.0) * u) * 0.05;
            nthd = nthd > 8.0 ? 8.0 : (nthd < -8.0 ? -8.0 : nthd);
            th = th + nthd * 0.05;
            thd = nthd;
            sincos(th, &sn, &cs);
            const float rew = (float)(-cost);
            ep_steps++;
            ep_ret += rew;
            bool trunc = ep_steps >= p.horizon;                       // TimeLimit(200)
            p.action[row] = a;
            p.logprob[row] = lp;
            p.reward[row] = rew;
            p.next_state[row * 3 + 0] = ((float)cs - om[0]) * oi[0];
            p.next_state[row * 3 + 1] = ((float)sn - om[1]) * oi[1];
            p.next_state[row * 3 + 2] = ((float)thd - om[2]) * oi[2];
            if (trunc) { ret_sum += ep_ret; ret_cnt++; episode++; }
            if (t == p.T - 1) trunc = true;                           // src/ppo.cu:70-74
            p.terminated[row] = 0;
            p.truncated[row] = trunc ? 1 : 0;
            if (trunc && t < p.T - 1) reset();                        // src/ppo.cu:64-66
        }
        // no barrier needed here: the next writers of hA (warp 0, rows 0..2) only race with readers of the
        // action-mean layer, which all passed the barrier above
    }
    if (warp == 0) {
        if (live) { p.ret_sum[env] = ret_sum; p.ret_cnt[env] = ret_cnt; }
        if (p.obs_partial) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float mean = wmean[k], m2 = wm2[k], n = wn;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {       // fixed butterfly order -> deterministic
                    const float mb = __shfl_xor_sync(kFull, mean, o), m2b = __shfl_xor_sync(kFull, m2, o), nb = __shfl_xor_sync(kFull, n, o);
                    if ((lane & o) == 0) welford_merge(mean, m2, n, mb, m2b, nb);
                    else { float ma = mb, m2a = m2b, na = nb; welford_merge(ma, m2a, na, mean, m2, n); mean = ma; m2 = m2a; n = na; }
                }
                if (lane == 0) {
                    float* o3 = p.obs_partial + ((size_t)blockIdx.x * 3 + k) * 3;
                    o3[0] = mean; o3[1] = m2; o3[2] = n;
                }
            }
        }
    }
}

// Fold the per-CTA triples of one rollout into the running float64 state (fixed order) and refresh mean / 1/(std+eps).
__global__ void obs_stats_merge_kernel(const float* __restrict__ partial, int blocks, double* __restrict__ stat,
                                       float* __restrict__ mean_out, float* __restrict__ inv_std_out) {
    const int k = threadIdx.x;
    if (k >= 3) return;
    double mean = stat[k * 3], m2 = stat[k * 3 + 1], n = stat[k * 3 + 2];
    for (int b = 0; b < blocks; b++) {
        const float* q = partial + ((size_t)b * 3 + k) * 3;
        const double mb = q[0], m2b = q[1], nb = q[2];
        if (nb > 0.0) {
            const double delta = mb - mean, nn = n + nb;
            mean += delta * nb / nn;
            m2 += m2b + delta * delta * n * nb / nn;
            n = nn;
        }
    }
    stat[k * 3] = mean; stat[k * 3 + 1] = m2; stat[k * 3 + 2] = n;
    mean_out[k] = (float)mean;
    inv_std_out[k] = (float)(1.0 / (sqrt(n > 0.0 ? m2 / n : 1.0) + 1e-8));
}

// fixed-order reduction of the per-env episode returns -> {sum of returns, episodes}
__global__ void __launch_bounds__(1024) return_stats_kernel(const float* ret_sum, const int* ret_cnt, int n, float* out) {
    __shared__ double s_sum[1024];
    __shared__ int s_cnt[1024];
    double s = 0.0;
    int c = 0;
    for (int i = threadIdx.x; i < n; i += 1024) { s += ret_sum[i]; c += ret_cnt[i]; }
    s_sum[threadIdx.x] = s; s_cnt[threadIdx.x] = c;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) { s_sum[threadIdx.x] += s_sum[threadIdx.x + k]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = (float)s_sum[0]; out[1] = (float)s_cnt[0]; }
}

void device_rollout(DeviceEnv* e, GaussianPolicy* policy, TrajectoryBuffer* buffer, int T,
                    const float* obs_mean, const float* obs_inv_std, float* return_stats) {
    if (policy->state_size != 3 || policy->action_size != 1) B200_FATAL("device Pendulum needs state_size 3 / action_size 1");
    if ((long long)e->n_envs * T > buffer->capacity) B200_FATAL("buffer capacity %d < n_envs*T = %lld", buffer->capacity, (long long)e->n_envs * T);
    (void)obs_mean; (void)obs_inv_std;
    static int use64 = -1;
    if (use64 < 0) { const char* v = getenv("PPO_B200_ROLLOUT"); use64 = (v && strcmp(v, "old") == 0) ? 0 : 1; }
    Rollout64Args r{};
    bool narrow = true;                       // rollout64_kernel handles layers of up to 128 units (64-unit blocks)
    {
        NetDev* ndp = net_dev(policy->mu);
        for (int w : ndp->sizes) narrow = narrow && w <= 128;
    }
    if (use64 && narrow && fused_image64(policy->mu, &r.net, &r.image)) {
        const int blocks = div_up(e->n_envs, kR64E);
        if (e->obs_norm && blocks > e->obs_partial_cap) {
            CUDA_CHECK(cudaStreamSynchronize(stream()));
            if (e->d_obs_partial) CUDA_CHECK(cudaFree(e->d_obs_partial));
            e->d_obs_partial = dmalloc<float>((size_t)blocks * 9);
            e->obs_partial_cap = blocks;
        }
        r.log_std = policy->d_log_std;
        r.n_envs = e->n_envs; r.T = T;
        r.seed = e->seed; r.rollout = e->rollouts++;
        r.state = buffer->d_state_p; r.next_state = buffer->d_next_state_p; r.action = buffer->d_action_p;
        r.reward = buffer->d_reward_p; r.logprob = buffer->d_logprob_p;
        r.terminated = reinterpret_cast<unsigned char*>(buffer->d_terminated_p);
        r.truncated = reinterpret_cast<unsigned char*>(buffer->d_truncated_p);
        r.obs_mean = e->obs_norm ? e->d_obs_mean : nullptr;
        r.obs_inv_std = e->obs_norm ? e->d_obs_inv_std : nullptr;
        r.ret_sum = e->d_ret_sum; r.ret_cnt = e->d_ret_cnt;
        r.obs_partial = e->obs_norm ? e->d_obs_partial : nullptr;
        r.horizon = e->base.horizon;
        const size_t smem = ((size_t)r.net.img_floats + 2 * r.net.max_width_pad * kR64E + 4 * 8 * kR64E) * sizeof(float);
        static size_t configured64 = 0;
        if (smem > configured64) {
            CUDA_CHECK(cudaFuncSetAttribute(rollout64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured64 = smem;
        }
        B200_LAUNCH(rollout64_kernel, blocks, kR64Threads, smem, r);
        if (e->obs_norm) B200_LAUNCH(obs_stats_merge_kernel, 1, 32, 0, e->d_obs_partial, blocks, e->d_obs_stat, e->d_obs_mean, e->d_obs_inv_std);
        if (return_stats) B200_LAUNCH(return_stats_kernel, 1, 1024, 0, e->d_ret_sum, e->d_ret_cnt, e->n_envs, return_stats);
        return;
    }
    if (e->obs_norm) B200_FATAL("observation normalisation needs a policy net the 64-wide kernels support");
    RolloutArgs a{};
    a.net = make_view(policy->mu);
    a.log_std = policy->d_log_std;
    a.n_envs = e->n_envs; a.T = T; a.S = 3; a.A = 1;
    a.seed = e->seed; a.rollout = e->rollouts++;
    a.state = buffer->d_state_p; a.next_state = buffer->d_next_state_p; a.action = buffer->d_action_p;
    a.reward = buffer->d_reward_p; a.logprob = buffer->d_logprob_p;
    a.terminated = reinterpret_cast<unsigned char*>(buffer->d_terminated_p);
    a.truncated = reinterpret_cast<unsigned char*>(buffer->d_truncated_p);
    a.obs_mean = obs_mean; a.obs_inv_std = obs_inv_std;
    a.ret_sum = e->d_ret_sum; a.ret_cnt = e->d_ret_cnt;
    a.raw_obs_partial = nullptr;
    a.horizon = e->base.horizon;
    size_t smem = (size_t)2 * a.net.max_width * kRollE * sizeof(float);
    a.weights_in_smem = (smem + (size_t)a.net.param_count * sizeof(float)) <= 200 * 1024;
    if (a.weights_in_smem) smem += (size_t)a.net.param_count * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    B200_LAUNCH(rollout_kernel, div_up(e->n_envs, kRollE), kRollE * kRollWarps, smem, a);
    if (return_stats) B200_LAUNCH(return_stats_kernel, 1, 1024, 0, e->d_ret_sum, e->d_ret_cnt, e->n_envs, return_stats);
}

// ================================ per-step sample for opaque host envs ================================
struct SampleArgs {
    NetView net;
    const float* log_std;
    const float* state;     // host-mapped or device, [m][S]
    float* action;          // [m][A]
    float* logprob;         // [m]
    int S, A;
    int draws[16];          // raw rand() values, consumed in the reference's order
    const int* draws_ptr;   // when more than 16 draws are needed
    int stage_params;       // parameters fit in shared memory next to the two activation vectors
};

__global__ void __launch_bounds__(128) sample_action_kernel(const SampleArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row = blockIdx.x;
    float* hA = smem;
    float* hB = smem + p.net.max_width;
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) hA[k] = p.state[(size_t)row * p.S + k];
    // This kernel is pure latency (one CTA, one env step): fetch ALL parameters with independent coalesced loads in one
    // round trip instead of chasing them unit by unit through L2 (12 us -> 4 us per step of the host-env rollout).
    const float* w = p.net.params;
    if (p.stage_params) {
        float* wsm = smem + 2 * p.net.max_width;
        for (int i = threadIdx.x; i < p.net.param_count; i += blockDim.x) wsm[i] = __ldg(p.net.params + i);
        w = wsm;
    }
    __syncthreads();
    float* hin = hA;
    float* hout = hB;
    for (int i = 0; i < p.net.num_layers - 1; i++) {
        const int n = p.net.sizes[i], l = p.net.sizes[i + 1];
        for (int j = warp; j < l; j += 4) {          // one warp per output unit, lanes over k
            float acc = 0.f;
            for (int k = lane; k < n; k += 32) acc = fmaf(w[(size_t)j * n + k], hin[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) hout[j] = act_apply(acc + w[(size_t)n * l + j], p.net.acts[i]);
        }
        w += (size_t)n * l + l;
        __syncthreads();
        float* tmp = hin; hin = hout; hout = tmp;
    }
    if (threadIdx.x == 0) {
        // generate_gaussian_noise, src/policy.cu:46-65 (loop bound corrected for n >= 6),
        // fed with the host's rand() draws; noise index base = row * A
        const int A = p.A;
        const int* dr = p.draws_ptr ? p.draws_ptr : p.draws;
        const float rmax = 2147483647.f;   // (float)RAND_MAX
        float noise[32];
        const int n_total = gridDim.x * A;
        for (int j = 0; j < A; j++) {
            const int gi = row * A + j;     // global noise index
            float z;
            if (n_total == 1 || ((n_total & 1) && gi == n_total - 1)) {
                const int base = (n_total == 1) ? 0 : (n_total - 1);   // the odd tail draws after the pairs
                const float u1 = __fdiv_rn((float)dr[base], rmax);
                const double th = 2 * kPiD * (double)(float)dr[base + 1] / 2147483647.0;
                z = sqrtf(-2.f * logf(u1)) * cosf((float)th);
            } else {
                const int pair = gi >> 1;
                const float u1 = __fdiv_rn((float)dr[2 * pair], rmax), u2 = __fdiv_rn((float)dr[2 * pair + 1], rmax);
                const float r = sqrtf(-2.f * logf(u1));
                const float theta = (float)(2 * kPiD * (double)u2);
                z = (gi & 1) ? r * sinf(theta) : r * cosf(theta);
            }
            noise[j] = z;
        }
        float act[32];
        for (int j = 0; j < A; j++) {
            act[j] = hin[j] + noise[j] * expf(p.log_std[j]);       // src/policy.cu:85
            p.action[(size_t)row * A + j] = act[j];
        }
        p.logprob[row] = log_prob_dev(hin, p.log_std, act, A);
    }
}

// dynamic shared memory of sample_action_kernel; decides whether the parameters are staged
static size_t sample_smem(SampleArgs& a) {
    const size_t base = 2 * (size_t)a.net.max_width * sizeof(float);
    const size_t with_params = base + (size_t)a.net.param_count * sizeof(float);
    a.stage_params = with_params <= 160 * 1024 ? 1 : 0;
    const size_t need = a.stage_params ? with_params : base;
    static size_t configured = 48 * 1024;
    if (need > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(sample_action_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        configured = need;
    }
    return need;
}

void launch_sample_action(GaussianPolicy* policy, const float* state, float* action, float* logprob,
                          const int* rand_draws, int n_draws) {
    SampleArgs a{};
    a.net = make_view(policy->mu);
    a.log_std = policy->d_log_std;
    a.state = state; a.action = action; a.logprob = logprob;
    a.S = policy->state_size; a.A = policy->action_size;
    if (a.A > 32) B200_FATAL("action_size > 32 unsupported");
    if (n_draws > 16) B200_FATAL("launch_sample_action: more than 16 draws need sample_action()");
    for (int i = 0; i < n_draws; i++) a.draws[i] = rand_draws[i];
    a.draws_ptr = nullptr;
    const size_t sm = sample_smem(a);
    B200_LAUNCH(sample_action_kernel, 1, 128, sm, a);
}

// ================================ persistent sampler for opaque host envs ==============================
// collect_trajectories with an opaque host env (the reference's hook signature: one env, host pointers, one step at a
// time, src/ppo.cu:54-79) needs one policy sample per env step.  One kernel launch + stream sync per step costs ~20 us; the
// reference's default run makes 3000 of them per iteration.  Instead ONE single-CTA kernel stays resident for the whole
// rollout and talks to the host through a mailbox in mapped pinned memory:
//   request  words  req[0] = {seq : cmd}, req[1..S] = {seq : state bits}, req[1+S..] = {seq : rand() draw}
//   response words  resp[0..A-1] = {seq : action bits}, resp[A] = {seq : log-prob bits}
// Every word is one naturally atomic 64-bit store carrying its own sequence tag (no fences, no separate flag): the
// device polls its request words with ONE warp-wide system-scope load per PCIe round trip, the host polls its own memory.
// The weights are staged once per launch in shared memory in the k-major image layout (thread j owns output unit j, the
// k-loop runs sequentially with separate mul and add = the reference's sequential-k sgemv order, src/mat_mul.cu:28-44).
// The kernel leaves on a stop request or after `idle_limit` clocks without one (the host relaunches it on demand), so a
// host that dies mid-rollout never leaves a kernel spinning.
constexpr int kMailReq = 192, kMailResp = 64;        // 64-bit words
struct SamplerMailbox { unsigned long long req[kMailReq]; unsigned long long resp[kMailResp]; unsigned long long exit_word; };
struct HostSamplerArgs {
    FusedNet net;
    const float* image;
    const float* log_std;
    SamplerMailbox* box;          // mapped pinned host memory
    int S, A, n_draws;
    unsigned int first_seq, launch_id;
    long long idle_limit;
};
enum { kSamplerStep = 1, kSamplerStop = 2 };

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(128) host_sampler_kernel(const HostSamplerArgs p) {
    extern __shared__ __align__(16) float smem[];
    const FusedNet& net = p.net;
    float* img = smem;
    float* hA = smem + net.img_floats;
    float* hB = hA + 128;
    int* draws = reinterpret_cast<int*>(hB + 128);       // 16 ints
    const int tid = threadIdx.x;
    for (int i = tid; i < net.img_floats; i += blockDim.x) img[i] = __ldg(p.image + i);
    __syncthreads();
    const int n_words = 1 + p.S + p.n_draws;
    unsigned int seq = p.first_seq;
    long long idle_since = clock64();
    for (;;) {
        // ---- wait for request `seq`
        bool stop = false;
        for (;;) {
            bool ok = true;
            int cmd = 0;
            for (int w = tid; w < n_words; w += blockDim.x) {
                const unsigned long long v = ld_sys_u64(p.box->req + w);
                if ((unsigned int)(v >> 32) != seq) ok = false;
                else if (w == 0) cmd = (int)(unsigned int)v;
                else if (w <= p.S) hA[w - 1] = __uint_as_float((unsigned int)v);
                else draws[w - 1 - p.S] = (int)(unsigned int)v;
            }
            const bool timed_out = tid == 0 && clock64() - idle_since > p.idle_limit;
            if (__syncthreads_or((tid == 0 && cmd == kSamplerStop) || timed_out)) { stop = true; break; }
            if (__syncthreads_and(ok)) break;
        }
        if (stop) break;
        // ---- mu = net(state): thread j = output unit j
        float* hin = hA;
        float* hout = hB;
        for (int l = 0; l < net.L; l++) {
            const int n = net.sizes[l], o = net.sizes[l + 1];
            if (tid < o) {
                const float* wt = img + net.wt_off[l] + tid;
                float acc = 0.f;
                for (int k = 0; k < n; k++) acc = __fadd_rn(acc, __fmul_rn(hin[k], wt[k * net.ldw[l]]));
                hout[tid] = act_apply(__fadd_rn(acc, img[net.bs_off[l] + tid]), net.acts[l]);
            }
            __syncthreads();
            float* tmp = hin; hin = hout; hout = tmp;
        }
        if (tid == 0) {
            // generate_gaussian_noise (src/policy.cu:46-65, loop bound corrected for n >= 6) from the host's rand() draws
            const int A = p.A;
            const float rmax = 2147483647.f;
            float act[32];
            for (int j = 0; j < A; j++) {
                float z;
                if (A == 1 || ((A & 1) && j == A - 1)) {
                    const int base = (A == 1) ? 0 : (A - 1);
                    const float u1 = __fdiv_rn((float)draws[base], rmax);
                    const double th = 2 * kPiD * (double)(float)draws[base + 1] / 2147483647.0;
                    z = sqrtf(-2.f * logf(u1)) * cosf((float)th);
                } else {
                    const int pair = j >> 1;
                    const float u1 = __fdiv_rn((float)draws[2 * pair], rmax), u2 = __fdiv_rn((float)draws[2 * pair + 1], rmax);
                    const float r = sqrtf(-2.f * logf(u1));
                    const float theta = (float)(2 * kPiD * (double)u2);
                    z = (j & 1) ? r * sinf(theta) : r * cosf(theta);
                }
                act[j] = hin[j] + z * expf(p.log_std[j]);           // src/policy.cu:85
            }
            const float lp = log_prob_dev(hin, p.log_std, act, A);
            const unsigned long long tag = (unsigned long long)seq << 32;
            for (int j = 0; j < A; j++) st_sys_u64(p.box->resp + j, tag | __float_as_uint(act[j]));
            st_sys_u64(p.box->resp + A, tag | __float_as_uint(lp));
            idle_since = clock64();
        }
        __syncthreads();        // hA / draws are rewritten by the next poll
        ++seq;
    }
    if (tid == 0) st_sys_u64(&p.box->exit_word, ((unsigned long long)p.launch_id << 32) | seq);
}

struct HostSampler {
    SamplerMailbox* box = nullptr;
    unsigned int seq = 0, launch_id = 0;
    bool running = false;
    GaussianPolicy* policy = nullptr;
};
static HostSampler g_sampler;

static bool sampler_enabled() {
    static int cached = -1;
    if (cached < 0) { const char* e = getenv("PPO_B200_MAILBOX"); cached = (e && e[0] == '0') ? 0 : 1; }
    return cached == 1;
}

static void sampler_launch(GaussianPolicy* policy, int n_draws, unsigned int first_seq) {
    HostSamplerArgs a{};
    const float* image = nullptr;
    if (!fused_image64(policy->mu, &a.net, &image)) B200_FATAL("host sampler on an unsupported net");
    a.image = image;
    a.log_std = policy->d_log_std;
    a.box = g_sampler.box;
    a.S = policy->state_size; a.A = policy->action_size; a.n_draws = n_draws;
    a.first_seq = first_seq;
    a.launch_id = ++g_sampler.launch_id;
    a.idle_limit = 100000000ll;       // ~50 ms at 1.9 GHz
    const size_t smem = (size_t)(a.net.img_floats + 256 + 16) * sizeof(float);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(host_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    B200_LAUNCH(host_sampler_kernel, 1, 128, smem, a);
    g_sampler.running = true;
    g_sampler.policy = policy;
}

// true when this policy / env shape can use the mailbox kernel
bool host_sampler_supported(GaussianPolicy* policy) {
    if (!sampler_enabled()) return false;
    FusedNet net;
    const float* image;
    const int S = policy->state_size, A = policy->action_size;
    if (S > 128 || A > 8 || 1 + S + 16 > kMailReq) return false;
    return fused_image64(policy->mu, &net, &image);
}

static bool sampler_exited() {
    const unsigned long long w = __atomic_load_n(&g_sampler.box->exit_word, __ATOMIC_ACQUIRE);
    return (unsigned int)(w >> 32) == g_sampler.launch_id;
}

// One policy sample through the mailbox: state[S] (host) -> action[A], *logprob (host).
void host_sampler_step(GaussianPolicy* policy, const float* state, float* action, float* logprob, const int* draws, int n_draws) {
    HostSampler& hs = g_sampler;
    if (!hs.box) {
        hs.box = hmalloc_pinned<SamplerMailbox>(1);
        memset(hs.box, 0, sizeof(SamplerMailbox));
    }
    const int S = policy->state_size, A = policy->action_size;
    const unsigned int seq = ++hs.seq;
    if (seq == 0) B200_FATAL("host sampler sequence wrapped");
    const unsigned long long tag = (unsigned long long)seq << 32;
    for (int k = 0; k < S; k++) {
        unsigned int bits;
        memcpy(&bits, state + k, 4);
        __atomic_store_n(&hs.box->req[1 + k], tag | bits, __ATOMIC_RELAXED);
    }
    for (int d = 0; d < n_draws; d++) __atomic_store_n(&hs.box->req[1 + S + d], tag | (unsigned int)draws[d], __ATOMIC_RELAXED);
    __atomic_store_n(&hs.box->req[0], tag | (unsigned long long)kSamplerStep, __ATOMIC_RELEASE);
    if (hs.running && (hs.policy != policy || sampler_exited())) hs.running = false;
    if (!hs.running) sampler_launch(policy, n_draws, seq);
    // wait for the response; relaunch if the kernel idled out in between
    long long spins = 0;
    for (;;) {
        bool ok = true;
        for (int j = 0; j <= A && ok; j++)
            ok = (unsigned int)(__atomic_load_n(&hs.box->resp[j], __ATOMIC_ACQUIRE) >> 32) == seq;
        if (ok) break;
        if ((++spins & 0x3ff) == 0) {
            if (sampler_exited()) { hs.running = false; sampler_launch(policy, n_draws, seq); }
            if (spins > (1ll << 34)) B200_FATAL("host sampler: no response from the device");
            const cudaError_t err = cudaStreamQuery(stream());
            if (err != cudaSuccess && err != cudaErrorNotReady) B200_FATAL("host sampler kernel failed: %s", cudaGetErrorString(err));
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    for (int j = 0; j < A; j++) {
        const unsigned int bits = (unsigned int)hs.box->resp[j];
        memcpy(action + j, &bits, 4);
    }
    const unsigned int lb = (unsigned int)hs.box->resp[A];
    memcpy(logprob, &lb, 4);
}

// End of a rollout: tell the resident kernel to leave (it would also leave by itself after the idle limit).
void host_sampler_stop() {
    HostSampler& hs = g_sampler;
    if (!hs.box || !hs.running) return;
    if (!sampler_exited()) {
        const unsigned int seq = ++hs.seq;
        __atomic_store_n(&hs.box->req[0], ((unsigned long long)seq << 32) | (unsigned long long)kSamplerStop, __ATOMIC_RELEASE);
    }
    hs.running = false;
}

__global__ void pendulum_step_kernel(double* theta, double* theta_dot, const float* action, float* obs, float* reward, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PendulumState s{theta[i], theta_dot[i]};
    reward[i] = pendulum_step(s, action[i]);
    theta[i] = s.th; theta_dot[i] = s.thd;
    pendulum_obs(s, obs + 3 * (size_t)i);
}

// ================================ host envs ================================================================
// toy env, src/env.c:6-33 (file-global state, like the reference)
static float toy_state = 0;
static int toy_step = 0;
static void reset_simple_env(float* obs) { toy_state = 0; toy_step = 0; obs[0] = 0; }
static void step_simple_env(float* action, float* obs, float* reward, bool* terminated, bool* truncated, int action_size) {
    (void)action_size;
    toy_state += fmaxf(fminf(action[0], 1), -1);
    obs[0] = toy_state;
    toy_step += 1;
    if (toy_state >= 5) { reward[0] = 1; terminated[0] = true; truncated[0] = false; }
    else if (toy_step >= 15) { reward[0] = 0; terminated[0] = false; truncated[0] = true; }
    else { reward[0] = 0; terminated[0] = false; truncated[0] = false; }
}
static void free_simple_env() {}

// native host Pendulum (one instance, context-free hooks like src/gym_env.c:3)
static PendulumState pend_state;
static int pend_steps = 0;
static uint64_t pend_rng = 0;
static uint64_t next_u64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double next_unit(uint64_t& s) { return (double)(next_u64(s) >> 11) * (1.0 / 9007199254740992.0); }
static void pend_reset_state(PendulumState& st, int& steps, uint64_t& rng) {
    st.th = (2.0 * next_unit(rng) - 1.0) * kPiD;
    st.thd = 2.0 * next_unit(rng) - 1.0;
    steps = 0;
}
static void reset_pendulum_env(float* obs) { pend_reset_state(pend_state, pend_steps, pend_rng); pendulum_obs(pend_state, obs); }
static void step_pendulum_env(float* action, float* obs, float* reward, bool* terminated, bool* truncated, int action_size) {
    (void)action_size;
    *reward = pendulum_step(pend_state, action[0]);
    pend_steps++;
    pendulum_obs(pend_state, obs);
    *terminated = false;
    *truncated = pend_steps >= 200;
}
static void free_pendulum_env() {}

// hooks of the device env: env 0's twin on the host (used only when a caller drives the Env through
// the reference's one-step API, e.g. eval_ppo)
static uint64_t dev_host_rng = 0;
static void reset_device_env(float* obs) {
    DeviceEnv* e = g_device_env;
    pend_reset_state(e->host_state, e->host_steps, dev_host_rng);
    pendulum_obs(e->host_state, obs);
}
static void step_device_env(float* action, float* obs, float* reward, bool* terminated, bool* truncated, int action_size) {
    (void)action_size;
    DeviceEnv* e = g_device_env;
    *reward = pendulum_step(e->host_state, action[0]);
    e->host_steps++;
    pendulum_obs(e->host_state, obs);
    *terminated = false;
    *truncated = e->host_steps >= 200;
}
static void free_device_env() {
    DeviceEnv* e = g_device_env;
    if (!e) return;
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    CUDA_CHECK(cudaFree(e->d_ret_sum));
    CUDA_CHECK(cudaFree(e->d_ret_cnt));
    CUDA_CHECK(cudaFree(e->d_obs_stat));
    CUDA_CHECK(cudaFree(e->d_obs_mean));
    CUDA_CHECK(cudaFree(e->d_obs_inv_std));
    if (e->d_obs_partial) CUDA_CHECK(cudaFree(e->d_obs_partial));
    e->magic = 0;
    g_device_env = nullptr;
}

}  // namespace b200

using namespace b200;

extern "C" {

Env* create_simple_env(int id, int seed) {     // src/env.c:41-51
    (void)id; (void)seed;
    Env* env = (Env*)malloc(sizeof(Env));
    env->state_size = 1;
    env->action_size = 1;
    env->reset_env = reset_simple_env;
    env->step_env = step_simple_env;
    env->free_env = free_simple_env;
    env->horizon = 15;
    env->gamma = 0.99;
    return env;
}

Env* create_pendulum_env(int id, int seed) {
    (void)id;
    Env* env = (Env*)malloc(sizeof(Env));
    env->state_size = 3;
    env->action_size = 1;
    env->reset_env = reset_pendulum_env;
    env->step_env = step_pendulum_env;
    env->free_env = free_pendulum_env;
    env->horizon = 200;           // scripts/gym_env.py:21 (spec.max_episode_steps)
    env->gamma = 0.99;            // src/gym_env.c:102
    pend_rng = 0x5eed0000ull + (uint64_t)(uint32_t)seed;
    return env;
}

Env* create_gym_env(int id, int seed) {
    if (id != 0) B200_FATAL("create_gym_env: only id 0 (Pendulum-v1) has a native implementation");
    return create_pendulum_env(id, seed);
}

extern "C" void ppo_b200_get_obs_norm(const Env* env, float* mean, float* std_, double* count) {
    b200::DeviceEnv* e = b200::as_device_env(const_cast<Env*>(env));
    if (!e) B200_FATAL("ppo_b200_get_obs_norm: not a device env");
    double st[9];
    CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    CUDA_CHECK(cudaMemcpy(st, e->d_obs_stat, sizeof(st), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++) {
        mean[k] = (float)st[k * 3];
        std_[k] = (float)sqrt(st[k * 3 + 2] > 0 ? st[k * 3 + 1] / st[k * 3 + 2] : 1.0);
    }
    if (count) *count = st[2];
}

Env* create_pendulum_env_cuda(int n_envs, int seed) {
    ensure_device();
    if (g_device_env) B200_FATAL("only one device env may exist at a time (context-free Env hooks)");
    DeviceEnv* e = (DeviceEnv*)calloc(1, sizeof(DeviceEnv));
    e->base.state_size = 3;
    e->base.action_size = 1;
    e->base.horizon = 200;
    e->base.gamma = 0.99;
    e->base.reset_env = reset_device_env;
    e->base.step_env = step_device_env;
    e->base.free_env = free_device_env;
    e->magic = kDeviceEnvMagic;
    e->n_envs = n_envs;
    e->seed = 0xB200ull * 0x9e3779b97f4a7c15ull + (uint64_t)(uint32_t)seed;
    e->rollouts = 0;
    e->d_ret_sum = dmalloc<float>(n_envs);
    e->d_ret_cnt = dmalloc<int>(n_envs);
    e->obs_norm = false;
    e->d_obs_stat = dmalloc<double>(9);
    e->d_obs_mean = dmalloc<float>(3);
    e->d_obs_inv_std = dmalloc<float>(3);
    e->d_obs_partial = nullptr;
    e->obs_partial_cap = 0;
    device_env_reset_obs_norm(e);
    dev_host_rng = 0xd00d0000ull + (uint64_t)(uint32_t)seed;
    g_device_env = e;
    return &e->base;
}

int ppo_b200_env_is_device(const Env* env) { return as_device_env(const_cast<Env*>(env)) != nullptr; }
int ppo_b200_env_num_envs(const Env* env) {
    DeviceEnv* e = as_device_env(const_cast<Env*>(env));
    return e ? e->n_envs : 1;
}

void ppo_b200_pendulum_step(double* theta, double* theta_dot, const float* action, float* obs, float* reward, int n) {
    if (n <= 0) return;
    B200_LAUNCH(pendulum_step_kernel, div_up(n, 256), 256, 0, theta, theta_dot, action, obs, reward, n);
}

// src/policy.cu:76-89 with host pointers, any m: noise draws made on the host with rand() in the
// reference's order, everything else on the device.
void sample_action(GaussianPolicy* policy, float* state, float* action, float* log_prob, int m) {
    const int S = policy->state_size, A = policy->action_size, n = m * A;
    std::vector<int> draws;
    if (n == 1) { draws.push_back(rand()); draws.push_back(rand()); }
    else {
        for (int i = 0; i + 1 < n; i += 2) { draws.push_back(rand()); draws.push_back(rand()); }
        if (n & 1) { draws.push_back(rand()); draws.push_back(rand()); }
    }
    HostStage st;
    st.add(state, (size_t)m * S * 4, true, false);
    st.add(action, (size_t)m * A * 4, false, true);
    st.add(log_prob, (size_t)m * 4, false, true);
    st.add(draws.data(), draws.size() * 4, true, false);
    st.upload();
    SampleArgs a{};
    a.net = make_view(policy->mu);
    a.log_std = policy->d_log_std;
    a.state = st.dev<float>(0); a.action = st.dev<float>(1); a.logprob = st.dev<float>(2);
    a.S = S; a.A = A;
    a.draws_ptr = st.dev<int>(3);
    if (A > 32) B200_FATAL("action_size > 32 unsupported");
    const size_t sm = sample_smem(a);
    B200_LAUNCH(sample_action_kernel, m, 128, sm, a);
    st.download();
}

}  // extern "C"
