// policy.cu — Gaussian policy + loss kernels.
//
// Stage kernels behind the reference entry points (include/policy.h, include/loss.h, ppo.h):
//   log-prob              src/policy.cu:67-74, 113-125 (the CUDA twin there is only right for A==1)
//   log-prob backward     src/policy.cu:101-111, 141-158 (float atomicAdd -> here a fixed-order sum)
//   PPO-clip loss + grad  src/ppo.cu:82-107, 109-169
//   MSE + derivative      src/loss.cu:5-23, 25-83
// and the two FUSED heads the training path uses (one launch each, no host sync, no cudaMalloc):
//   value head   : grad = 2 (y - target) / m ; loss accumulated on the device
//   policy head  : log-prob -> ratio -> clipped surrogate -> grad_mu[m][A], grad_log_std[A], loss
// Reductions are two-level (block partials, last block sums them in block order), so every result
// is deterministic.  Arithmetic follows the reference's mixed float/double expressions.
#include <unordered_map>

#include "common.cuh"
#include "internal.h"

namespace b200 {

constexpr double kPi = 3.14159265358979323846;  // include/policy.h:7

__device__ __forceinline__ float log_prob_row(const float* mu, const float* log_std, const float* action, int A) {
    // src/policy.cu:67-74: float logprob = -0.5 * A * logf(2*pi); logprob -= log_std + 0.5 * powf(z, 2)
    float logprob = (float)(-0.5 * A * (double)logf((float)(2 * kPi)));
    for (int j = 0; j < A; j++) {
        const float z = __fdiv_rn(__fsub_rn(action[j], mu[j]), expf(log_std[j]));
        logprob = (float)((double)logprob - ((double)log_std[j] + 0.5 * (double)__fmul_rn(z, z)));
    }
    return logprob;
}

__global__ void __launch_bounds__(256)
log_prob_kernel(const float* __restrict__ mu, const float* __restrict__ log_std,
                const float* __restrict__ action, float* __restrict__ out, int m, int A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = log_prob_row(mu + (size_t)i * A, log_std, action + (size_t)i * A, A);
}

// Block-level sum of one float per thread; result valid in thread 0.
__device__ __forceinline__ float block_sum_256(float v, float* red /* [8] */) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    return t;
}

// ppo-clip per-sample math, src/ppo.cu:89-98
__device__ __forceinline__ void ppo_clip(float adv, float logp, float logp_old, float epsilon, int m_total,
                                         float& loss_term, float& grad) {
    const float ratio = expf(__fsub_rn(logp, logp_old));
    const bool adv_pos = adv > 0.f;
    const bool hi = ratio > 1.f + epsilon, lo = ratio < 1.f - epsilon;
    const float sel = adv_pos ? (hi ? 1.f + epsilon : ratio) : (lo ? 1.f - epsilon : ratio);
    loss_term = __fmul_rn(adv, sel);
    const int keep = adv_pos ? !hi : !lo;
    grad = __fdiv_rn(__fmul_rn(__fmul_rn((float)(-keep), adv), ratio), (float)m_total);
}

// partial[block] = sum of adv*sel ; grad_logprob per sample
__global__ void __launch_bounds__(256)
policy_loss_kernel(float* __restrict__ partial, float* __restrict__ grad_logprob, const float* __restrict__ adv,
                   const float* __restrict__ logp, const float* __restrict__ logp_old, float epsilon, int m) {
    __shared__ float red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float term = 0.f;
    if (i < m) {
        float g;
        ppo_clip(adv[i], logp[i], logp_old[i], epsilon, m, term, g);
        grad_logprob[i] = g;
    }
    const float s = block_sum_256(term, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
mse_kernel(float* __restrict__ partial, const float* __restrict__ y, const float* __restrict__ t, int count) {
    __shared__ float red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float term = 0.f;
    if (i < count) { const float d = __fsub_rn(t[i], y[i]); term = __fmul_rn(d, d); }
    const float s = block_sum_256(term, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
mse_grad_kernel(float* __restrict__ grad, const float* __restrict__ y, const float* __restrict__ t, int count, int denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) grad[i] = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(y[i], t[i])), (float)denom);  // src/loss.cu:20
}

// grad_mu + per-block partials of grad_log_std (src/policy.cu:104-110 with per-sample grad_in[i])
__global__ void __launch_bounds__(256)
log_prob_backward_kernel(const float* __restrict__ grad_in, const float* __restrict__ mu,
                         const float* __restrict__ log_std, const float* __restrict__ action,
                         float* __restrict__ grad_mu, float* __restrict__ partial /* [blocks][A] */, int m, int A) {
    __shared__ float red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int j = 0; j < A; j++) {
        float term = 0.f;
        if (i < m) {
            const float e2 = expf(-2.f * log_std[j]);
            const float diff = __fsub_rn(action[(size_t)i * A + j], mu[(size_t)i * A + j]);
            const float g = grad_in[i];
            grad_mu[(size_t)i * A + j] = __fmul_rn(__fmul_rn(diff, e2), g);
            term = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff, diff), e2)), g);
        }
        const float s = block_sum_256(term, red);
        if (threadIdx.x == 0) partial[(size_t)blockIdx.x * A + j] = s;
    }
}

// out[j] = add[j] + sum_b partial[b][j]  (block order -> deterministic)
__global__ void sum_partials_kernel(float* __restrict__ out, const float* __restrict__ partial, int blocks, int width, float add) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    float s = 0.f;
    for (int b = 0; b < blocks; b++) s += partial[(size_t)b * width + j];
    out[j] = s + add;
}

// ---- fused heads (training path) ----------------------------------------------------------------
struct HeadWork {
    float* partial;       // [blocks][width]
    unsigned* counter;    // arrival counter, self-resetting
};

__global__ void __launch_bounds__(256)
value_head_kernel(const float* __restrict__ y, const float* __restrict__ target, float* __restrict__ grad, int m,
                  int m_total, float* __restrict__ partial, unsigned* counter, float* loss_slot) {
    __shared__ float red[8];
    __shared__ bool last;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float term = 0.f;
    if (i < m) {
        const float yi = y[i], ti = target[i];
        grad[i] = __fdiv_rn(__fmul_rn(2.f, __fsub_rn(yi, ti)), (float)m_total);
        const float d = __fsub_rn(ti, yi);
        term = __fmul_rn(d, d);
    }
    const float s = block_sum_256(term, red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float t = 0.f;
        for (unsigned b = 0; b < gridDim.x; b++) t += ((volatile float*)partial)[b];
        *loss_slot += t / (float)m_total;
        *counter = 0;
    }
}

template <int MAXA>
__global__ void __launch_bounds__(256)
policy_head_kernel(const float* __restrict__ mu, const float* __restrict__ log_std,
                   const float* __restrict__ action, const float* __restrict__ logp_old,
                   const float* __restrict__ adv, int m, int A, int m_total, float epsilon, float ent_coeff,
                   float* __restrict__ logp_out, float* __restrict__ grad_mu, float* __restrict__ grad_log_std,
                   float* __restrict__ partial /* [blocks][A+1] */, unsigned* counter, float* loss_slot) {
    __shared__ float red[8];
    __shared__ bool last;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float term = 0.f, g = 0.f;
    float diff[MAXA];
    if (i < m) {
        const float lp = log_prob_row(mu + (size_t)i * A, log_std, action + (size_t)i * A, A);
        if (logp_out) logp_out[i] = lp;
        ppo_clip(adv[i], lp, logp_old[i], epsilon, m_total, term, g);
#pragma unroll
        for (int j = 0; j < MAXA; j++)
            if (j < A) diff[j] = __fsub_rn(action[(size_t)i * A + j], mu[(size_t)i * A + j]);
    }
    const float sl = block_sum_256(term, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (A + 1) + A] = sl;
#pragma unroll
    for (int j = 0; j < MAXA; j++) {
        if (j < A) {   // A is uniform across the block
            float t = 0.f;
            if (i < m) {
                const float e2 = expf(-2.f * log_std[j]);
                grad_mu[(size_t)i * A + j] = __fmul_rn(__fmul_rn(diff[j], e2), g);
                t = __fmul_rn(__fadd_rn(-1.f, __fmul_rn(__fmul_rn(diff[j], diff[j]), e2)), g);
            }
            const float s = block_sum_256(t, red);
            if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (A + 1) + j] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        const volatile float* vp = partial;
        if ((int)threadIdx.x < A) {
            float t = 0.f;
            for (unsigned b = 0; b < gridDim.x; b++) t += vp[(size_t)b * (A + 1) + threadIdx.x];
            // src/ppo.cu:436-438: the plain-C twin adds grad_entropy = -ent_coeff to every log_std grad
            grad_log_std[threadIdx.x] = t + (-ent_coeff);
        }
        if (threadIdx.x == 32) {
            float t = 0.f;
            for (unsigned b = 0; b < gridDim.x; b++) t += vp[(size_t)b * (A + 1) + A];
            // entropy, src/policy.cu:171-178
            float entropy = (float)(A * 0.5 * (1 + log(2 * kPi)));
            for (int j = 0; j < A; j++) entropy += log_std[j];
            *loss_slot += -t / (float)m_total - ent_coeff * entropy;
        }
        __syncthreads();
        if (threadIdx.x == 0) *counter = 0;
    }
}

static HeadWork head_work(int blocks, int width) {
    // [counter 256 B][partials]
    char* ws = static_cast<char*>(scratch(kScratchLoss, 256 + (size_t)blocks * width * sizeof(float)));
    return HeadWork{reinterpret_cast<float*>(ws + 256), reinterpret_cast<unsigned*>(ws)};
}

void launch_log_prob(const float* mu, const float* log_std, const float* action, float* out, int m, int A) {
    if (m <= 0) return;
    B200_LAUNCH(log_prob_kernel, div_up(m, 256), 256, 0, mu, log_std, action, out, m, A);
}

void launch_value_head(const float* y, const float* target, float* grad, int m, int m_total, float* loss_slot) {
    if (m <= 0) return;
    const int blocks = div_up(m, 256);
    HeadWork w = head_work(blocks, 1);
    B200_LAUNCH(value_head_kernel, blocks, 256, 0, y, target, grad, m, m_total, w.partial, w.counter, loss_slot);
}

void launch_policy_head(const float* mu, const float* log_std, const float* action, const float* logp_old,
                        const float* adv, int m, int A, int m_total, float epsilon, float ent_coeff,
                        float* logp_out, float* grad_mu, float* grad_log_std, float* loss_slot) {
    if (m <= 0) return;
    const int blocks = div_up(m, 256);
    HeadWork w = head_work(blocks, A + 1);
    if (A <= 1)
        B200_LAUNCH(policy_head_kernel<1>, blocks, 256, 0, mu, log_std, action, logp_old, adv, m, A, m_total, epsilon,
                    ent_coeff, logp_out, grad_mu, grad_log_std, w.partial, w.counter, loss_slot);
    else if (A <= 8)
        B200_LAUNCH(policy_head_kernel<8>, blocks, 256, 0, mu, log_std, action, logp_old, adv, m, A, m_total, epsilon,
                    ent_coeff, logp_out, grad_mu, grad_log_std, w.partial, w.counter, loss_slot);
    else if (A <= 32)
        B200_LAUNCH(policy_head_kernel<32>, blocks, 256, 0, mu, log_std, action, logp_old, adv, m, A, m_total, epsilon,
                    ent_coeff, logp_out, grad_mu, grad_log_std, w.partial, w.counter, loss_slot);
    else
        B200_FATAL("action_size %d > 32 is not supported by the fused policy head", A);
}

static float entropy_from_host(const float* log_std, int A) {
    float entropy = A * 0.5 * (1 + log(2 * kPi));   // src/policy.cu:172, a constant plus A adds
    for (int j = 0; j < A; j++) entropy += log_std[j];
    return entropy;
}

static float sum_block_partials(const float* d_partial, int blocks) {
    std::vector<float> h(blocks);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), d_partial, blocks * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    float s = 0.f;
    for (int b = 0; b < blocks; b++) s += h[b];   // the reference also adds its block sums on the host (src/ppo.cu:161-164)
    return s;
}

}  // namespace b200

using namespace b200;

extern "C" {

// ---- include/policy.h ----------------------------------------------------------------------------------
GaussianPolicy* create_gaussian_policy(int* layer_sizes, char** activation_functions, int num_layers, float init_std) {
    GaussianPolicy* policy = (GaussianPolicy*)malloc(sizeof(GaussianPolicy));
    policy->state_size = layer_sizes[0];
    policy->action_size = layer_sizes[num_layers - 1];
    policy->mu = create_neural_network(layer_sizes, activation_functions, num_layers);
    const int A = policy->action_size;
    policy->log_std = (float*)malloc(A * sizeof(float));
    policy->log_std_grad = (float*)calloc(A, sizeof(float));
    // d_log_std and its grad share one allocation so that {log_std} is a flat Adam vector too
    policy->d_log_std = dmalloc<float>(A);
    policy->d_log_std_grad = dmalloc<float>(A);
    policy->input_action = nullptr;
    policy->d_input_action = nullptr;
    for (int i = 0; i < A; i++) policy->log_std[i] = logf(init_std);   // src/policy.cu:22-24
    CUDA_CHECK(cudaMemcpy(policy->d_log_std, policy->log_std, A * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(policy->d_log_std_grad, 0, A * sizeof(float)));
    return policy;
}

void free_gaussian_policy(GaussianPolicy* policy) {
    if (!policy) return;
    free_neural_network(policy->mu);
    free(policy->log_std);
    free(policy->log_std_grad);
    // input_action / d_input_action are BORROWED from the caller (src/policy.cu:92,128); the
    // reference frees them (src/policy.cu:37,41), which is a latent double free — not replicated.
    CUDA_CHECK(cudaFree(policy->d_log_std));
    CUDA_CHECK(cudaFree(policy->d_log_std_grad));
    free(policy);
}

void compute_log_prob_cuda(GaussianPolicy* policy, float* out, float* state, float* action, int m) {
    policy->d_input_action = action;
    net_forward(policy->mu, state, m, false);
    launch_log_prob(policy->mu->d_output, policy->d_log_std, action, out, m, policy->action_size);
}

void log_prob_backwards_cuda(GaussianPolicy* policy, float* grad_in, float* grad_mu, float* grad_log_std, int m) {
    const int A = policy->action_size, blocks = div_up(m, 256);
    float* partial = static_cast<float*>(scratch(kScratchLoss, 256 + (size_t)blocks * A * sizeof(float))) + 64;
    B200_LAUNCH(log_prob_backward_kernel, blocks, 256, 0, grad_in, policy->mu->d_output, policy->d_log_std,
                policy->d_input_action, grad_mu, partial, m, A);
    B200_LAUNCH(sum_partials_kernel, div_up(A, 64), 64, 0, grad_log_std, partial, blocks, A, 0.f);
}

float compute_entropy_cuda(GaussianPolicy* policy) {   // src/policy.cu:180-193 (blocking D2H, like the reference)
    std::vector<float> h(policy->action_size);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), policy->d_log_std, h.size() * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
    return entropy_from_host(h.data(), policy->action_size);
}

float compute_entropy(GaussianPolicy* policy) { return entropy_from_host(policy->log_std, policy->action_size); }

void policy_to_host(GaussianPolicy* policy) {          // src/policy.cu:195-198
    nn_write_weights_to_host(policy->mu);
    CUDA_CHECK(cudaMemcpyAsync(policy->log_std, policy->d_log_std, policy->action_size * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// host-pointer twins, staged (src/policy.cu:91-111)
void compute_log_prob(GaussianPolicy* policy, float* out, float* state, float* action, int m) {
    policy->input_action = action;
    const int S = policy->state_size, A = policy->action_size;
    nn_write_weights_to_device(policy->mu);
    CUDA_CHECK(cudaMemcpyAsync(policy->d_log_std, policy->log_std, A * sizeof(float), cudaMemcpyHostToDevice, stream()));
    HostStage st;
    st.add(state, (size_t)m * S * 4, true, false);
    st.add(action, (size_t)m * A * 4, true, false);
    st.add(out, (size_t)m * 4, false, true);
    st.upload();
    net_forward(policy->mu, st.dev<float>(0), m, false);
    launch_log_prob(policy->mu->d_output, policy->d_log_std, st.dev<float>(1), st.dev<float>(2), m, A);
    free(policy->mu->output);
    policy->mu->output = (float*)malloc((size_t)m * A * sizeof(float));
    CUDA_CHECK(cudaMemcpyAsync(policy->mu->output, policy->mu->d_output, (size_t)m * A * sizeof(float), cudaMemcpyDeviceToHost, stream()));
    st.download();
}

void log_prob_backwards(GaussianPolicy* policy, float* grad_in, float* grad_mu, float* grad_log_std, int m) {
    const int A = policy->action_size;
    HostStage st;
    st.add(grad_in, (size_t)m * 4, true, false);
    st.add(policy->input_action, (size_t)m * A * 4, true, false);
    st.add(grad_mu, (size_t)m * A * 4, false, true);
    st.add(grad_log_std, (size_t)A * 4, false, true);
    st.upload();
    policy->d_input_action = st.dev<float>(1);
    log_prob_backwards_cuda(policy, st.dev<float>(0), st.dev<float>(2), st.dev<float>(3), m);
    st.download();
    policy->d_input_action = nullptr;
}

void save_policy(GaussianPolicy* policy, FILE* file) {  // src/policy.cu:201-205
    fwrite(policy->log_std, sizeof(float), policy->action_size, file);
    save_neural_network(policy->mu, file);
}

GaussianPolicy* load_policy(FILE* file, int state_size, int action_size) {   // src/policy.cu:207-227
    GaussianPolicy* policy = (GaussianPolicy*)malloc(sizeof(GaussianPolicy));
    policy->state_size = state_size;
    policy->action_size = action_size;
    policy->log_std = (float*)malloc(action_size * sizeof(float));
    policy->log_std_grad = (float*)calloc(action_size, sizeof(float));
    policy->d_log_std = dmalloc<float>(action_size);
    policy->d_log_std_grad = dmalloc<float>(action_size);
    if (fread(policy->log_std, sizeof(float), action_size, file) != (size_t)action_size) B200_FATAL("checkpoint truncated");
    CUDA_CHECK(cudaMemcpy(policy->d_log_std, policy->log_std, action_size * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(policy->d_log_std_grad, 0, action_size * sizeof(float)));
    policy->mu = load_neural_network(file);
    policy->input_action = nullptr;
    policy->d_input_action = nullptr;
    return policy;
}

// ---- include/ppo.h:36,39 ---------------------------------------------------------------------------------
float policy_loss_and_grad_cuda(float* grad_logprob, float* grad_entropy, float* adv, float* logprobs,
                                float* old_logprobs, float entropy, float ent_coeff, float epsilon, int m) {
    const int blocks = div_up(m, 256);
    float* partial = static_cast<float*>(scratch(kScratchLoss, 256 + (size_t)blocks * sizeof(float))) + 64;
    B200_LAUNCH(policy_loss_kernel, blocks, 256, 0, partial, grad_logprob, adv, logprobs, old_logprobs, epsilon, m);
    const float s = sum_block_partials(partial, blocks);
    *grad_entropy = -ent_coeff;
    // src/ppo.cu:101-103; the CUDA twin subtracts the entropy term once per block (src/ppo.cu:141) — a
    // defect not replicated.
    return -s / m - ent_coeff * entropy;
}

float policy_loss_and_grad(float* grad_logprob, float* grad_entropy, float* adv, float* logprobs,
                           float* old_logprobs, float entropy, float ent_coeff, float epsilon, int m) {
    HostStage st;
    st.add(grad_logprob, (size_t)m * 4, false, true);
    st.add(adv, (size_t)m * 4, true, false);
    st.add(logprobs, (size_t)m * 4, true, false);
    st.add(old_logprobs, (size_t)m * 4, true, false);
    st.upload();
    const float loss = policy_loss_and_grad_cuda(st.dev<float>(0), grad_entropy, st.dev<float>(1), st.dev<float>(2),
                                                 st.dev<float>(3), entropy, ent_coeff, epsilon, m);
    st.download();
    return loss;
}

// ---- include/loss.h ----------------------------------------------------------------------------------------
float mean_squared_error_cuda(float* y, float* y_true, int m, int n) {
    const int count = m * n, blocks = div_up(count, 256);
    float* partial = static_cast<float*>(scratch(kScratchLoss, 256 + (size_t)blocks * sizeof(float))) + 64;
    B200_LAUNCH(mse_kernel, blocks, 256, 0, partial, y, y_true, count);
    return sum_block_partials(partial, blocks) / count;
}

void mean_squared_error_derivative_cuda(float* grad, float* y, float* y_true, int m, int n) {
    const int count = m * n;
    B200_LAUNCH(mse_grad_kernel, div_up(count, 256), 256, 0, grad, y, y_true, count, count);
}

float mean_squared_error(float* y, float* y_true, int m, int n) {
    HostStage st;
    st.add(y, (size_t)m * n * 4, true, false);
    st.add(y_true, (size_t)m * n * 4, true, false);
    st.upload();
    return mean_squared_error_cuda(st.dev<float>(0), st.dev<float>(1), m, n);
}

void mean_squared_error_derivative(float* grad, float* y, float* y_true, int m, int n) {
    HostStage st;
    st.add(grad, (size_t)m * n * 4, false, true);
    st.add(y, (size_t)m * n * 4, true, false);
    st.add(y_true, (size_t)m * n * 4, true, false);
    st.upload();
    mean_squared_error_derivative_cuda(st.dev<float>(0), st.dev<float>(1), st.dev<float>(2), m, n);
    st.download();
}

}  // extern "C"
