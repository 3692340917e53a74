// internal.h — C++ interfaces between the translation units of libppo_b200.so (not exported).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200 {

// Stage host arrays through device scratch, run `fn` on device pointers, copy outputs back.
struct HostStage {
    struct Item { void* host; size_t bytes; bool in, out; void* dev; };
    std::vector<Item> items;
    void* add(const void* host, size_t bytes, bool in, bool out) {
        items.push_back({const_cast<void*>(host), bytes, in, out, nullptr});
        return nullptr;
    }
    void upload() {
        size_t total = 0;
        for (auto& it : items) total += (it.bytes + 255) & ~size_t(255);
        char* base = static_cast<char*>(scratch(kScratchStage, total));
        size_t off = 0;
        for (auto& it : items) {
            it.dev = base + off;
            off += (it.bytes + 255) & ~size_t(255);
            if (it.in) CUDA_CHECK(cudaMemcpyAsync(it.dev, it.host, it.bytes, cudaMemcpyHostToDevice, stream()));
        }
    }
    void download() {
        for (auto& it : items)
            if (it.out) CUDA_CHECK(cudaMemcpyAsync(it.host, it.dev, it.bytes, cudaMemcpyDeviceToHost, stream()));
        CUDA_CHECK(cudaStreamSynchronize(stream()));
    }
    template <typename T> T* dev(int i) { return static_cast<T*>(items[i].dev); }
};


// ---- gae.cu -----------------------------------------------------------------------------------
struct GaeWork {
    double* stats_d;   // device {mean, M2, n}
    float* stats_f;    // device {mean, std}
    float4* wstats;
    int nchunks;
};
GaeWork gae_scan(const float* reward, const float* v, const float* v_next, const bool* terminated,
                 const bool* truncated, int n, float gamma, float lambda, float* advantage,
                 float* adv_target);
void gae_merge_ranks(const double* triples_dev, int world, float* stats_f);
void gae_normalize(float* advantage, int n, const float* stats_f);

// ---- adam.cu ----------------------------------------------------------------------------------
// w,g,m,v flat device vectors.  If partials != nullptr the gradient is first formed as the
// fixed-order sum over `splits` slabs of length `stride` (deterministic split-K reduction) and
// written to g.
void adam_flat(float* w, float* g, float* m, float* v, int n, float lr, float beta1, float beta2,
               int time_step, const float* partials, int splits, size_t stride);
void reduce_partials(float* g, const float* partials, int splits, size_t stride, int n);

// ---- gemm.cu ----------------------------------------------------------------------------------
// y[m][l] = act(x[m][n] . W[l][n]^T + b[l])            (reference mat_mul.cu:132-163 + activation)
void linear_forward(float* y, const float* x, const float* W, const float* b, int m, int n, int l, int act);
// gx[m][n] = (g[m][l] . W[l][n]) * act'(xin)  where xin[m][n] is the POST-activation input of this
// layer produced with activation `act_prev` (kActNone: no mask).      (mat_mul.cu:175-188 + K8)
void linear_backward_input(float* gx, const float* g, const float* W, const float* xin, int m, int n, int l, int act_prev);
// partial dW / db slabs: for split s, rows [s*rows_per_split, ...): gW_part[s][l][n] = g^T . x,
// gb_part[s][l] = column sums of g.  Slab s of layer tensor lives at part + s*stride.  (mat_mul.cu:195-208, K9)
// db_done: the db slabs are already written (honoured by the tensor-core path only; every other path computes db itself)
void linear_backward_params(float* gW_part, float* gb_part, size_t stride, int splits, const float* g,
                            const float* x, int m, int n, int l, bool db_done = false);
int choose_splits(int m, size_t param_count);
void activation_inplace(float* x, long long count, int act);
void activation_grad_inplace(const float* y, float* grad, long long count, int act);

// ---- narrow.cu (streaming kernels for layers with one side <= 32 columns) ---------------------------
bool narrow_first_layer_backward(float* gW_part, float* gb_part, size_t stride, int splits, const float* g, const float* x, int m, int n, int l);
// dW slabs of an l <= 8 wide head and, when gx != null, gx = (g W) act'(h) (+ its lo companion when gx_lo != null) in one pass over h
// (+ the column sums of gx = db slabs of the layer below when gb_below != null)
bool narrow_head_backward(float* gW_part, size_t stride, int splits, float* gx, float* gx_lo, float* gb_below, const float* g, const float* h,
                          const float* W, int m, int n, int l, int act_prev);

bool narrow_head_forward(float* y, const float* h, const float* W, const float* b, int m, int n, int l, int act);
// first layer (n <= 32 inputs); y_lo != null additionally receives the 3xTF32 lo companion of y
bool narrow_first_forward(float* y, float* y_lo, const float* x, const float* W, const float* b, int m, int n, int l, int act);

// ---- tc_gemm.cu (tcgen05 TF32 tensor-core path for wide layers) --------------------------------
bool tc_shape_ok(const void* a, const void* b, int lda_cols, int ldb_cols);
void tc_linear_forward(float* y, const float* x, const float* W, const float* b, int m, int n, int l, int act);
void tc_linear_backward_input(float* gx, const float* g, const float* W, const float* xin, int m, int n, int l, int act_prev);
void tc_linear_backward_weights(float* gW_part, size_t stride, int splits, const float* g, const float* x, int m, int n, int l);
// the NEXT tensor-core dX launch also writes the column sums of its output (= the db slabs of the layer below) to gb_part
void tc_request_colsum(float* gb_part, size_t stride, int splits);
bool tc_colsum_pending();
void tc_round_copy(const float* src, float* dst, size_t n);   // RNA-rounded TF32 shadow of a weight arena
int matmul_precision();   // 0 fp32 FFMA (default), 1 TF32 tcgen05 for layers with n, l >= 64, 2 BF16 tcgen05 (widths % 8 == 0),
                          // 3 "3xTF32" split on tcgen05: fp32-accurate (meets the 1e-5 parity of mode 0), layers with n, l >= 64
// 3xTF32 split mode: every operand comes with lo = x - top19bits(x) (tc_split_lo)
void tc_split_lo(const float* src, float* lo, size_t n);
void tc_linear_forward_x3(float* y, const float* x, const float* xlo, const float* W, const float* Wlo, const float* b, int m, int n, int l, int act);
void tc_linear_backward_input_x3(float* gx, const float* g, const float* glo, const float* W, const float* Wlo, const float* xin, int m, int n,
                                 int l, int act_prev);
void tc_linear_backward_weights_x3(float* gW_part, size_t stride, int splits, const float* g, const float* glo, const float* x, const float* xlo,
                                   int m, int n, int l);
// bf16 operand mode (tc_gemm.cu): shadows are __nv_bfloat16 arrays, passed as void* between translation units
bool tc_bf16_shape_ok(int m, int n, int l);
void tc_linear_forward_bf16_v(float* y, void* y16, const void* x16, const void* W16, const float* b, int m, int n, int l, int act);
void tc_linear_backward_input_bf16_v(float* gx, void* gx16, const void* g16, const void* Wt16, const float* xin, int m, int n, int l, int act_prev);
void tc_linear_backward_weights_bf16_v(float* gW_part, size_t stride, int splits, const void* g16, const void* x16, int m, int n, int l);
void tc_to_bf16_v(const float* src, void* dst, size_t n);
void tc_weights_bf16_v(const float* W, void* W16, void* Wt16, int l, int n);
void launch_colsum(float* gb_part, size_t stride, int splits, const float* g, int m, int l);
void tc_colsum_bf16_v(float* gb_part, size_t stride, int splits, const void* g16, int m, int l);
void tc_pad_bf16_v(const float* src, void* dst, size_t rows, int n, int npad);
void tc_unpad_slabs(const float* src, float* gW_part, size_t stride, int splits, int l, int n, int npad);

// ---- nn.cu ------------------------------------------------------------------------------------
struct NetDev {               // device-side view of one NeuralNetwork (side table keyed by pointer)
    int num_layers = 0;       // reference convention: number of sizes (weight layers = num_layers-1)
    std::vector<int> sizes;
    std::vector<int> acts;
    float* params = nullptr;  // flat arena W0,b0,W1,b1,...
    float* grads = nullptr;
    size_t param_count = 0;
    std::vector<size_t> w_off, b_off;
    // activation cache: layer inputs; act[0] is either owned copy or borrowed pointer
    std::vector<float*> a;    // a[i] : [cap][sizes[i]]
    std::vector<float*> gx;   // gradient wrt a[i]
    int cap_fwd = 0, cap_bwd = 0;
    bool a0_borrowed = false;
    float* a0_owned = nullptr;
    int a0_cap = 0;
    float* partials = nullptr;   // [splits][slab_stride()]
    size_t slab_stride() const { return (param_count + 31) & ~size_t(31); }   // 128-byte multiple: keeps every slab TMA/float4-aligned
    size_t partials_cap = 0;     // floats
    int last_splits = 1;
    int last_m = 0;
    float* params_tf32 = nullptr;   // RNA-rounded shadow of `params` read by the tensor-core layers
    // bf16 operand mode: per-layer weight copies W16 [out][in] | Wt16 [in][out] and shadows of the activations / gradients
    void* params_bf16 = nullptr;    // 2 * param_count bf16: [W16 of every layer at w_off | Wt16 of every layer at param_count + w_off]
    void* w0pad_bf16 = nullptr;     // first layer with fewer than 64 inputs: W16 [out][64] zero padded (K padded to one k-block)
    std::vector<void*> a16, gx16;   // a16[i] : bf16 [cap][sizes[i]] (null until needed); a16[0] is [cap][64] when layer 0 is K-padded
    std::vector<int> a16_cap, gx16_cap;
    float* params_lo = nullptr;     // 3xTF32 mode: lo companion of `params` (same offsets), refreshed per forward
    std::vector<void*> alo, glo;    // lo companions of a[i] / of the gradient wrt a[i] (fp32 [cap][sizes[i]], null until needed)
    std::vector<int> alo_cap, glo_cap;
    float* image = nullptr;      // pre-transposed weight image staged by the fused kernels (fused_mlp.cu)
    int image_floats = 0;
    bool image_dirty = true;     // set by every writer of `params` other than fused_reduce_adam_kernel
};
NetDev* net_dev(NeuralNetwork* nn);
void net_mark_params_written(const float* params_ptr);
// forward without copying the input (training path); output = a.back()
void net_forward(NeuralNetwork* nn, const float* input, int m, bool borrow_input);
// backward from grad wrt output (device, [m][out]); leaves split-K slabs in nd->partials
void net_backward_partials(NeuralNetwork* nn, const float* grad_out, int m);
void net_reduce_grads(NeuralNetwork* nn);   // partials -> nd->grads (fixed order)

// ---- fused_mlp.cu -----------------------------------------------------------------------------
constexpr int kFusedMaxLayers = 5;     // weight layers
struct FusedNet {
    int L;                              // weight layers
    int sizes[kFusedMaxLayers + 1];
    int acts[kFusedMaxLayers];
    int w_off[kFusedMaxLayers], b_off[kFusedMaxLayers];   // offsets in the flat parameter vector
    int wt_off[kFusedMaxLayers];        // offsets of Wt[in][ldw] inside the weight image (floats)
    int ldw[kFusedMaxLayers];           // row stride of Wt (>= pad4(out); the 64-wide tile kernel pads it by 4)
    int bs_off[kFusedMaxLayers];        // offsets of the layer biases inside the weight image (floats)
    int img_floats;                     // image size (multiple of 32 floats = 128 B; TMA bulk needs 16 B)
    int a_off[kFusedMaxLayers + 1];     // offsets of At buffers in shared memory (floats, after the image)
    int P;
    int max_width_pad;
};

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }
bool fused_image64(NeuralNetwork* nn, FusedNet* layout, const float** image);   // false: net outside the 64-wide kernels
bool fused_supported(NeuralNetwork* nn);
void fused_forward(NeuralNetwork* nn, const float* x, int m, float* y_out);
// One persistent cooperative launch for `n_epochs` x `num_batches` minibatches of one net (fused_phase_kernel).
// perms: [n_epochs][limit] device permutations.  Minibatch k of an epoch covers rows k*batch_stride + row0 .. + mb of the permutation.
bool fused_phase_supported(NeuralNetwork* nn);
void fused_phase_update(NeuralNetwork* nn, GaussianPolicy* policy, Adam* adam_net, Adam* adam_ls, float lr, const int* perms,
                        int limit, int mb, int m_total, int row0, int batch_stride, int num_batches, int n_epochs,
                        const TrajectoryBuffer* b, float epsilon, float ent_coeff, float* loss_slot, bool dp_peer);
bool fused_minibatch_update(NeuralNetwork* nn, GaussianPolicy* policy, Adam* adam_net, Adam* adam_ls, float lr,
                            const int* perm, int offset, int limit, int m, int m_total, const TrajectoryBuffer* b,
                            float epsilon, float ent_coeff, float* loss_slot, float* reduced_out, bool chained = false,
                            bool dp_peer = false);

// ---- policy.cu --------------------------------------------------------------------------------
void launch_log_prob(const float* mu, const float* log_std, const float* action, float* out, int m, int A);
// fused value head: grad[i] = 2 (y-t)/m_total ; loss_slot += sum (t-y)^2 / m_total
void launch_value_head(const float* y, const float* target, float* grad, int m, int m_total, float* loss_slot);
// fused policy head (src/policy.cu:91-111 + src/ppo.cu:82-107): log-prob, clipped surrogate,
// grad_mu[m][A], grad_log_std[A] (+ entropy grad), loss
void launch_policy_head(const float* mu, const float* log_std, const float* action, const float* logp_old,
                        const float* adv, int m, int A, int m_total, float epsilon, float ent_coeff,
                        float* logp_out, float* grad_mu, float* grad_log_std, float* loss_slot);

// ---- buffer.cu --------------------------------------------------------------------------------
int packed_row_floats(int S, int A);
void launch_pack_rows(float* packed, long long rows, int S, int A, const float* state, const float* action, const float* logprob,
                      const float* advantage, const float* adv_target);
void launch_gather_packed(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* packed, float* states, float* actions,
                          float* logprobs, float* advantages, float* adv_targets);
void launch_gather(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* state,
                   const float* action, const float* logprob, const float* advantage, const float* adv_target,
                   float* states, float* actions, float* logprobs, float* advantages, float* adv_targets);
void host_shuffle(int* idx, int limit);
void buffer_upload_inputs(TrajectoryBuffer* b);      // 7 input arrays host -> device (async), device set active
void buffer_download_outputs(TrajectoryBuffer* b);   // advantage, adv_target device -> host, host set active   // the reference's rand() swap chain, trajectory_buffer.cu:132-141

// ---- env.cu -----------------------------------------------------------------------------------
struct DeviceEnv;
DeviceEnv* as_device_env(Env* env);
int device_env_count(DeviceEnv* e);
// fused rollout: T steps for every env; writes the env-major buffer (index = env*T + t)
void device_rollout(DeviceEnv* e, GaussianPolicy* policy, TrajectoryBuffer* buffer, int T,
                    const float* obs_mean, const float* obs_inv_std, float* return_stats /* dev: sum, episodes */);
void device_env_reset_obs_norm(DeviceEnv* e);
void device_env_set_obs_norm(bool enabled);
// persistent mailbox sampler for opaque host envs (env.cu)
bool host_sampler_supported(GaussianPolicy* policy);
void host_sampler_step(GaussianPolicy* policy, const float* state, float* action, float* logprob, const int* draws, int n_draws);
void host_sampler_stop();
void launch_sample_action(GaussianPolicy* policy, const float* state_hostmapped, float* action_hostmapped,
                          float* logprob_hostmapped, const int* rand_draws, int n_draws);

// ---- dist.cu ----------------------------------------------------------------------------------
constexpr int kPeerMaxRanks = 8;
constexpr size_t kPeerCap = 1 << 15;                 // elements per (parity, source rank) receive lane: nets up to 32768 parameters
struct PeerArena {
    bool tried = false, ready = false;
    char* local = nullptr;                           // receive buffer: u64 [2 parities][kPeerMaxRanks sources][kPeerCap]
    char* base[kPeerMaxRanks] = {};
    unsigned long long epoch = 0;
};
struct PeerView {                                    // kernel argument: one gradient exchange over NVLink peer memory
    int ready, world, rank;
    unsigned int epoch;                              // tag of this exchange (never 0)
    unsigned long long* my_recv;                     // my receive lanes of this epoch's parity: [source rank][kPeerCap]
    unsigned long long* peer_recv[kPeerMaxRanks];    // lane [my rank] inside rank r's receive buffer (same parity)
    size_t parity_stride;                            // dist_peer_reserve views: pointers are parity 0, exchange e uses (e & 1) * parity_stride
};
PeerView dist_peer_next(size_t vec_floats);
// `count` consecutive exchanges for one persistent kernel: epoch = tag of the first; pointers address parity 0
PeerView dist_peer_reserve(size_t vec_floats, int count);
long long dist_spin_limit();                          // clock64 ticks a device-side wait may last (PPO_B200_SPIN_TIMEOUT_S, default 120 s)
bool dist_peer_ready();
bool dist_active();
int dist_rank();
int dist_world();
int dist_shard_mode();
void dist_allreduce_sum(float* buf, size_t count);
void dist_allgather_doubles(const double* send, double* recv, int count_per_rank);

}  // namespace b200
