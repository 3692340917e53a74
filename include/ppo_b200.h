/* include/ppo_b200.h — the drop-in boundary of the B200-native PPO training path.
 *
 * One C-ABI shared library (ppo.c_b200/libppo_b200.so) exports
 *   (1) every symbol of the reference's header set, with the reference's struct layouts, so that a
 *       caller written against cube1324/ppo.c (its src/main.c) relinks unchanged, and
 *   (2) additive `ppo_b200_*` / `create_pendulum_env*` entry points for what the reference cannot
 *       express (vectorised device envs, TxN buffers, data parallelism, stage-level kernels).
 * The thin headers ppo.h / policy.h / neural_network.h / trajectory_buffer.h / adam.h / loss.h /
 * mat_mul.h / activation_function.h / env.h / gym_env.h in this directory just include this file,
 * so `#include "ppo.h"` keeps working.
 *
 * Every declaration cites the reference interface it replaces (paths relative to the reference
 * repository root).  Semantics differences are stated next to the declaration.
 *
 * All arithmetic behind these entry points runs in hand-written sm_100a CUDA kernels; there is no
 * CPU fallback: without a usable CUDA device every compute entry point aborts with a message.
 * Error convention = the reference's debug build (include/cuda_helper.h:4-19): any CUDA/NCCL error
 * prints file:line to stderr and aborts; functions keep the reference's void/pointer returns.
 */
#ifndef PPO_B200_H
#define PPO_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846 /* reference include/policy.h:7 */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* The reference drags <cublas_v2.h> into every includer (include/neural_network.h:8-9) only for
 * this handle type.  The product never calls cuBLAS; the slot is kept for ABI compatibility. */
#if !defined(CUBLAS_V2_H_) && !defined(CUBLAS_API_H_)
typedef struct cublasContext* cublasHandle_t;
#endif

/* ================================================================================================
 * Types — field order and types are ABI (SURVEY.md §8b "Struct layout is ABI")
 * ============================================================================================== */

/* reference include/activation_function.h:10-13 */
typedef struct {
    void (*activation)(float* x, int m, int n);
    void (*activation_derivative)(float* x, float* grad, int m, int n);
} ActivationFunction;

/* reference include/neural_network.h:18-38.  In the product `d_weights/d_biases` (and the grads) of
 * all layers of one network are slices of ONE contiguous device arena ordered W0,b0,W1,b1,...
 * (the tensor order of src/adam.cu:25-42) so Adam and the gradient all-reduce see a flat vector. */
typedef struct {
    float* weights;
    float* biases;
    float* grad_weights;
    float* grad_biases;
    float* input;

    float* d_weights;
    float* d_biases;
    float* d_grad_weights;
    float* d_grad_biases;
    float* d_input;

    float* d_grad_x;

    ActivationFunction* activation_function;
    ActivationFunction* d_activation_function;
    int input_size;
    int output_size;
} Layer;

/* reference include/neural_network.h:41-53 */
typedef struct {
    Layer* layers;
    int num_layers;
    int output_size;

    int cache_m_forward;
    int cache_m_backward;

    float* output;
    float* d_output;
    char** activation_functions;

    cublasHandle_t cublas_handle; /* always NULL here */
} NeuralNetwork;

/* reference include/neural_network.h:55-58 */
typedef struct {
    float (*loss)(float* y, float* y_true, int m, int n);
    void (*loss_derivative)(float* grad, float* y, float* y_true, int m, int n);
} LossFunction;

/* reference include/policy.h:13-24 */
typedef struct {
    NeuralNetwork* mu;
    float* log_std;
    float* log_std_grad;
    float* d_log_std;
    float* d_log_std_grad;
    int state_size;
    int action_size;

    float* input_action;
    float* d_input_action;
} GaussianPolicy;

/* reference include/trajectory_buffer.h:15-62.  Three pointer sets: active / h_ / d_.  The product
 * allocates the h_ set as PINNED host memory (cudaHostAlloc) so the mirrors move at PCIe speed;
 * release only through free_trajectory_buffer. */
typedef struct TrajectoryBuffer TrajectoryBuffer;
struct TrajectoryBuffer {
    float* state_p;
    float* action_p;
    float* next_state_p;
    float* reward_p;
    float* logprob_p;
    float* advantage_p;
    float* adv_target_p;
    bool* terminated_p;
    bool* truncated_p;

    float* h_state_p;
    float* h_action_p;
    float* h_next_state_p;
    float* h_reward_p;
    float* h_logprob_p;
    float* h_advantage_p;
    float* h_adv_target_p;
    bool* h_terminated_p;
    bool* h_truncated_p;

    float* d_state_p;
    float* d_action_p;
    float* d_next_state_p;
    float* d_reward_p;
    float* d_logprob_p;
    float* d_advantage_p;
    float* d_adv_target_p;
    bool* d_terminated_p;
    bool* d_truncated_p;

    int* random_idx;
    int state_size;
    int action_size;
    int capacity;
    int idx;
    bool full;
    float* (*state)(TrajectoryBuffer* buffer, int idx);
    float* (*action)(TrajectoryBuffer* buffer, int idx);
    float* (*next_state)(TrajectoryBuffer* buffer, int idx);
    float* (*reward)(TrajectoryBuffer* buffer, int idx);
    float* (*logprob)(TrajectoryBuffer* buffer, int idx);
    float* (*advantage)(TrajectoryBuffer* buffer, int idx);
    float* (*adv_target)(TrajectoryBuffer* buffer, int idx);
    bool* (*terminated)(TrajectoryBuffer* buffer, int idx);
    bool* (*truncated)(TrajectoryBuffer* buffer, int idx);
};

/* reference include/adam.h:10-21.  For the *_cuda constructors `weights/grad_weights/lengths` are
 * device arrays exactly as in the reference (lengths = inclusive prefix sums, src/adam.cu:88-97). */
typedef struct {
    float** weights;
    float** grad_weights;
    int* lengths;
    float* m;
    float* v;
    float beta1;
    float beta2;
    int time_step;
    int size;
    int num_layers;
} Adam;

/* reference include/env.h:7-15: three context-free hooks + sizes. */
typedef struct {
    void (*free_env)();
    void (*reset_env)(float* state);
    void (*step_env)(float* action, float* obs, float* reward, bool* terminated, bool* truncated,
                     int action_size);
    int state_size;
    int action_size;
    int horizon;
    float gamma;
} Env;

/* reference include/ppo.h:15-28 */
typedef struct {
    TrajectoryBuffer* buffer;
    GaussianPolicy* policy;
    NeuralNetwork* V;
    Adam* adam_policy;
    Adam* adam_V;
    Adam* adam_entropy;
    float lambda;
    float epsilon;
    float ent_coeff;
    float lr_policy;
    float lr_V;
    bool use_cuda;
} PPO;

/* ================================================================================================
 * (1) The reference API.  `_cuda` twins take DEVICE pointers, like the reference's.
 *     The non-_cuda twins exist for link compatibility and take HOST pointers, but they too run on
 *     the GPU (staged through device memory) — the product contains no CPU arithmetic path.
 * ============================================================================================== */

/* ---- include/ppo.h:30-47 ------------------------------------------------------------------ */
PPO* create_ppo(char** activation_functions, int* layer_sizes, int num_layers, int buffer_size,
                float lr_policy, float lr_v, float lambda, float epsilon, float ent_coeff,
                float init_std, bool use_cuda);
void free_ppo(PPO* ppo);
/* src/ppo.cu:54-79.  Opaque host Env: one hook call per step, the policy forward / Box-Muller /
 * log-prob run in a one-CTA kernel per step; the two glibc rand() draws per step stay on the host
 * so the RNG stream is the reference's.  Device env (create_pendulum_env_cuda): the whole rollout
 * is one persistent kernel. */
void collect_trajectories(TrajectoryBuffer* buffer, Env* env, GaussianPolicy* policy, int steps);
void compute_gae(NeuralNetwork* V, TrajectoryBuffer* buffer, float gamma, float lambda); /* src/ppo.cu:326-369 */
float policy_loss_and_grad(float* grad_logprob, float* grad_entropy, float* adv, float* logprobs,
                           float* old_logprobs, float entropy, float ent_coeff, float epsilon, int m);
/* src/ppo.cu:261-323.  `horizon` is accepted and ignored: the segmented scan is exact for any
 * horizon (the reference's is only correct for horizon < 512). */
void compute_gae_cuda(NeuralNetwork* V, TrajectoryBuffer* buffer, float gamma, float lambda, int horizon);
float policy_loss_and_grad_cuda(float* grad_logprob, float* grad_entropy, float* adv, float* logprobs,
                                float* old_logprobs, float entropy, float ent_coeff, float epsilon, int m);
void train_ppo_epoch(PPO* ppo, Env* env, int steps_per_epoch, int batch_size, int n_epochs_policy,
                     int n_epochs_value);                                   /* src/ppo.cu:552-558 */
void eval_ppo(PPO* ppo, Env* env, int steps);                               /* src/ppo.cu:560-583 */
void save_ppo(PPO* ppo, const char* filename);                              /* src/ppo.cu:585-607 */
PPO* load_ppo(const char* filename, bool use_cuda);                         /* src/ppo.cu:610-648 */

/* ---- include/policy.h:26-41 ----------------------------------------------------------------- */
GaussianPolicy* create_gaussian_policy(int* layer_sizes, char** activation_functions, int num_layers,
                                       float init_std);
void free_gaussian_policy(GaussianPolicy* policy);
void sample_action(GaussianPolicy* policy, float* state, float* action, float* log_prob, int m);
void compute_log_prob(GaussianPolicy* policy, float* out, float* state, float* action, int m);
void log_prob_backwards(GaussianPolicy* policy, float* grad_in, float* grad_mu, float* grad_log_std, int m);
void compute_log_prob_cuda(GaussianPolicy* policy, float* out, float* state, float* action, int m);
/* grad_in is PER SAMPLE (m floats).  Identical to the reference for action_size == 1; for
 * action_size > 1 the reference indexes grad_in[i*A+j] out of bounds (src/policy.cu:106,153). */
void log_prob_backwards_cuda(GaussianPolicy* policy, float* grad_in, float* grad_mu, float* grad_log_std, int m);
float compute_entropy_cuda(GaussianPolicy* policy);
float compute_entropy(GaussianPolicy* policy);
void policy_to_host(GaussianPolicy* policy);
void save_policy(GaussianPolicy* policy, FILE* file);
GaussianPolicy* load_policy(FILE* file, int state_size, int action_size);

/* ---- include/neural_network.h:60-72 ----------------------------------------------------------- */
NeuralNetwork* create_neural_network(int* layer_sizes, char** activation_functions, int num_layers);
void forward_propagation(NeuralNetwork* nn, float* input, int m);
void free_neural_network(NeuralNetwork* nn);
void backward_propagation(NeuralNetwork* nn, float* grad_in, int m);
void forward_propagation_cuda(NeuralNetwork* nn, float* input, int m);
void backward_propagation_cuda(NeuralNetwork* nn, float* grad_in, int m);
void nn_write_weights_to_device(NeuralNetwork* nn);
void nn_write_weights_to_host(NeuralNetwork* nn);
void save_neural_network(NeuralNetwork* nn, FILE* file);
NeuralNetwork* load_neural_network(FILE* file);

/* ---- include/trajectory_buffer.h:66-79 --------------------------------------------------------- */
TrajectoryBuffer* create_trajectory_buffer(int capacity, int state_size, int action_size);
void free_trajectory_buffer(TrajectoryBuffer* buffer, bool use_cuda);
void shuffle_buffer(TrajectoryBuffer* buffer);
void get_batch(TrajectoryBuffer* buffer, int batch_idx, int batch_size, float* states, float* actions,
               float* logprobs, float* advantages, float* adv_targets);
void shuffle_buffer_cuda(TrajectoryBuffer* buffer);
void get_batch_cuda(TrajectoryBuffer* buffer, int batch_idx, int batch_size, float* states,
                    float* actions, float* logprobs, float* advantages, float* adv_targets);
void reset_buffer(TrajectoryBuffer* buffer);
void buffer_to_device(TrajectoryBuffer* buffer);
void buffer_to_host(TrajectoryBuffer* buffer);

/* ---- include/adam.h:24-38 ---------------------------------------------------------------------- */
Adam* create_adam(float** weights, float** grad_weights, int* length, int num_layers, int size,
                  float beta1, float beta2);
Adam* create_adam_from_nn(NeuralNetwork* nn, float beta1, float beta2);
void free_adam(Adam* adam);
void adam_update(Adam* adam, float lr);
Adam* create_adam_cuda(float** weights, float** grad_weights, int* length, int num_layers, int size,
                       float beta1, float beta2);
Adam* create_adam_from_nn_cuda(NeuralNetwork* nn, float beta1, float beta2);
void free_adam_cuda(Adam* adam);
void adam_update_cuda(Adam* adam, float lr);
void save_adam(Adam* adam, FILE* file, bool cuda);
Adam* load_adam(FILE* file, float** weights, float** grad_weights, int* length, bool cuda);
Adam* load_adam_from_nn(FILE* file, NeuralNetwork* nn, bool cuda);

/* ---- include/loss.h:10-14 ------------------------------------------------------------------------ */
float mean_squared_error(float* y, float* y_true, int m, int n);
void mean_squared_error_derivative(float* grad, float* y, float* y_true, int m, int n);
float mean_squared_error_cuda(float* y, float* y_true, int m, int n);
void mean_squared_error_derivative_cuda(float* grad, float* y, float* y_true, int m, int n);

/* ---- include/mat_mul.h:16-20 (handle ignored) ---------------------------------------------------- */
void mat_mul(float* out, float* x, float* weight, float* bias, int m, int n, int l);
void mat_mul_backwards(float* grad_x, float* grad_weight, float* grad_in, float* x, float* weight,
                       int m, int n, int l);
void mat_mul_cuda(cublasHandle_t handle, float* out, float* x, float* weight, float* bias, int m, int n, int l);
void mat_mul_backwards_cuda(cublasHandle_t handle, float* grad_x, float* grad_weight, float* grad_in,
                            float* x, float* weight, int m, int n, int l);

/* ---- include/activation_function.h:15-22.  Names: "relu", "tanh" (new), anything else = identity */
void ReLU(float* x, int m, int n);
void ReLU_derivative(float* x, float* grad, int m, int n);
void ReLU_cuda(float* x, int m, int n);
void ReLU_derivative_cuda(float* x, float* grad, int m, int n);
void Tanh_cuda(float* x, int m, int n);                          /* new */
void Tanh_derivative_cuda(float* x, float* grad, int m, int n);  /* new; x is the POST-activation value */
ActivationFunction* build_activation_function(char* name);
ActivationFunction* build_activation_function_cuda(char* name);

/* ---- include/env.h:18, include/gym_env.h:8 ------------------------------------------------------- */
Env* create_simple_env(int id, int seed);  /* src/env.c:41-51, toy "walk to 5" env */
/* gymnasium is not part of the product: id 0 returns the native Pendulum-v1 below (same sizes,
 * horizon 200, gamma 0.99 as src/gym_env.c:96-104 + scripts/gym_env.py:12); other ids abort. */
Env* create_gym_env(int id, int seed);
void openblas_set_num_threads(int n);      /* called by the reference's main.c:18; no-op */

/* ================================================================================================
 * (2) Additive extensions
 * ============================================================================================== */

/* ---- environments ---------------------------------------------------------------------------- */
/* Native host Pendulum-v1 behind the reference hooks (dynamics: public gymnasium definition,
 * SURVEY.md §A.10).  Own splitmix64 stream; never touches glibc rand(). */
Env* create_pendulum_env(int id, int seed);
/* n_envs vectorised Pendulum-v1 environments living on the device.  The returned Env has the
 * reference's sizes (S=3, A=1, horizon=200, gamma=0.99); its hooks drive env 0 only (for eval through
 * the reference path).  collect_trajectories / train_ppo_epoch recognise it and run the fused
 * rollout kernel; the buffer is then env-major: flat index = env * T + t, T = capacity / n_envs. */
Env* create_pendulum_env_cuda(int n_envs, int seed);
int  ppo_b200_env_is_device(const Env* env);
int  ppo_b200_env_num_envs(const Env* env);

/* ---- runtime plumbing (device memory for callers without a CUDA runtime of their own) --------- */
int   ppo_b200_device_count(void);
void  ppo_b200_set_device(int device);
void  ppo_b200_set_stream(void* cuda_stream); /* all later launches go to this cudaStream_t */
void* ppo_b200_malloc(size_t bytes);
void  ppo_b200_free(void* dptr);
void* ppo_b200_malloc_host(size_t bytes);     /* pinned */
void  ppo_b200_free_host(void* hptr);
void  ppo_b200_h2d(void* dst, const void* src, size_t bytes);
void  ppo_b200_d2h(void* dst, const void* src, size_t bytes);
void  ppo_b200_memset(void* dst, int value, size_t bytes);
void  ppo_b200_sync(void);
unsigned long long ppo_b200_launch_count(void); /* kernels launched by this library so far */
const char* ppo_b200_version(void);
/* measured fp32 FMA throughput of the current device in TFLOP/s: form 0 = FFMA with constant-bank operands (the pipe's peak),
 * 1 = scalar FFMA on a register 8x4 tile, 2 = packed FFMA2 on the same tile (the form the MLP kernels issue) */
double ppo_b200_measure_fp32_peak(int form);
/* Per-kernel timing without a profiler: between begin and end every launch of the library is
 * bracketed by a CUDA-event pair on the launching stream.  end() writes "name count total_ms" lines. */
void ppo_b200_profile_begin(void);
int  ppo_b200_profile_end(char* out, int out_bytes);
/* Debug aid (env PPO_B200_PHASE_DEBUG=1): %globaltimer stamps of the phases of the last fused minibatch kernel,
 * [blocks][16] 64-bit values (scratch/phase_debug.py prints the timeline).  No-op when the env var is not set. */
void ppo_b200_debug_phase_stamps(unsigned long long* out, int blocks);

/* ---- stage-level kernels on plain DEVICE arrays ------------------------------------------------ */
/* GAE + returns (src/ppo.cu:338-353) as one segmented reverse scan; optional normalisation
 * (src/ppo.cu:355-368) with float64-combined Welford statistics.  stats_out (device, 2 floats:
 * mean, std) may be NULL.  Requires nothing about T/N: any flat buffer, any done pattern. */
void ppo_b200_gae(const float* reward, const float* v, const float* v_next, const bool* terminated,
                  const bool* truncated, int n, float gamma, float lambda, float* advantage,
                  float* adv_target, int normalize, float* stats_out);
/* Adam (src/adam.cu:53-74) over one flat vector; bit-exact with the reference arithmetic. */
void ppo_b200_adam_flat(float* w, const float* g, float* m, float* v, int n, float lr, float beta1,
                        float beta2, int time_step /* already incremented */);
/* Gather (src/trajectory_buffer.cu:168-200) with explicit index array. */
void ppo_b200_gather(const int* idx, int offset, int limit, int batch_size, int S, int A,
                     const float* state, const float* action, const float* logprob,
                     const float* advantage, const float* adv_target, float* states, float* actions,
                     float* logprobs, float* advantages, float* adv_targets);
/* Row-packed mirror of the gathered fields (new): packed[row][PW] = state | action | logprob | advantage | adv_target | pad
 * with PW = ppo_b200_packed_row_floats(S, A) (the row rounded up to whole 32-byte sectors).  Built by ONE streaming pass
 * after GAE; ppo_b200_gather_packed then reads one contiguous, sector-aligned row per sample instead of five scattered
 * pieces (same outputs as ppo_b200_gather, bit for bit; src/trajectory_buffer.cu:168-200). */
int ppo_b200_packed_row_floats(int S, int A);
void ppo_b200_pack_rows(float* packed, long long rows, int S, int A, const float* state, const float* action,
                        const float* logprob, const float* advantage, const float* adv_target);
void ppo_b200_gather_packed(const int* idx, int offset, int limit, int batch_size, int S, int A, const float* packed,
                            float* states, float* actions, float* logprobs, float* advantages, float* adv_targets);
/* Device permutation (new): writes a uniformly random permutation of [0,n) from a counter-based
 * generator keyed by (seed, epoch).  NOT the reference's rand() chain — see shuffle_buffer_cuda. */
void ppo_b200_permutation(int* idx, int n, unsigned long long seed, unsigned long long epoch);
/* Pendulum dynamics on arrays (theta, theta_dot double; action float) — one step for n envs. */
void ppo_b200_pendulum_step(double* theta, double* theta_dot, const float* action, float* obs,
                            float* reward, int n);

/* ---- training-path controls -------------------------------------------------------------------- */
/* Update phase only (src/ppo.cu:485-538 minus the rollout): buffer_to_device, GAE, the value and
 * policy epochs, mirrors back to host.  The buffer's HOST arrays must be filled (limit = capacity). */
void ppo_b200_update(PPO* ppo, float gamma, int batch_size, int n_epochs_policy, int n_epochs_value);
/* Same, device-resident: no host<->device copies of the buffer (used after a device rollout or
 * with ppo_b200_buffer_upload); weights stay on the device until ppo_b200_sync_host. */
void ppo_b200_update_device(PPO* ppo, float gamma, int batch_size, int n_epochs_policy, int n_epochs_value);
void ppo_b200_buffer_upload(PPO* ppo);   /* host arrays -> device arrays, marks buffer full */
void ppo_b200_sync_host(PPO* ppo);       /* device buffer + weights -> host mirrors (src/ppo.cu:536-538) */
/* Whole iterations, device-resident, no host mirrors (bench `value`): rollout + update, n_iters times. */
void ppo_b200_train_iterations(PPO* ppo, Env* env, int n_iters, int batch_size, int n_epochs_policy,
                               int n_epochs_value);
/* permutation source: 0 = reference glibc rand() chain on the host (bit-exact indices),
 *                     1 = device counter-based permutation (no host work, no H2D),
 *                    -1 = auto (default): 0 for host envs / host-filled buffers (the reference-
 *                         compatible path), 1 after a device rollout (whose noise is Philox anyway). */
void ppo_b200_set_permutation_mode(PPO* ppo, int mode, unsigned long long seed);
/* kernel path of the update/GAE-forward: -1 default (fused small-net kernels when every layer width
 * is <= 128, else layer-wise; env PPO_B200_FUSED=0 disables), 0 = force layer-wise, 1 = force fused. */
void ppo_b200_set_kernel_path(int path);
/* matmul precision of dense layers whose in/out widths are >= 64 and batch >= 128:
 * 0 = fp32 FFMA (default; 1e-5 tolerance), 1 = TF32 tcgen05 tensor cores with fp32 accumulation
 * (wide-MLP configs; tolerance ~1e-3, stated separately), 2 = BF16 operands (tcgen05 kind::f16, fp32 accumulation in TMEM;
 * layers additionally need widths that are multiples of 8; tolerance ~1e-2, stated separately), 3 = "3xTF32": every fp32
 * operand is split exactly into hi (the 19 bits the tensor core reads) + lo and each contraction is the sum of the three
 * tcgen05 TF32 products hi.hi + hi.lo + lo.hi with fp32 accumulation -- as accurate as the fp32 FFMA path (1e-5 tolerance,
 * same parity tests) at a third of the TF32 rate; replaces cublasSgemm of src/mat_mul.cu:149-208 for the 2x256 nets.
 * Parameters, gradients and the optimiser stay fp32 in every mode.  Env PPO_B200_TF32=1 / 2 / 3 sets the default. */
void ppo_b200_set_matmul_precision(int mode);
/* raw tensor-core layer kernels (tests/bench): mode 0 forward (aux = bias), 1 backward-input
 * (aux = post-activation input), 2 backward-weights (out = `splits` slabs of l*n floats). */
void ppo_b200_tc_linear(int mode, float* out, const float* a, const float* b, const float* aux, int m, int n, int l,
                        int act, int splits);
/* the same three contractions in the 3xTF32 split mode (lo companions are formed in scratch first) */
void ppo_b200_tc_linear_x3(int mode, float* out, const float* a, const float* b, const float* aux, int m, int n, int l,
                           int act, int splits);
/* the same three contractions with BF16 operands: a / b are converted to bf16 scratch copies first (skip_convert: reuse the
 * copies of the previous call; convert_only: stop after the conversion), out16 (may be NULL) receives the bf16 shadow of the
 * fp32 output of modes 0 and 1. */
void ppo_b200_tc_linear_bf16(int mode, float* out, void* out16, const float* a, const float* b, const float* aux, int m, int n, int l,
                             int act, int splits, int convert_only, int skip_convert);
/* Running observation normalisation for the device rollout (new capability; the reference only has the
 * Welford merge, include/welford_var.h:33-40,58-66, applied to advantages).  The rollout kernel keeps a
 * Welford triple of the RAW observations per lane, merges them per CTA, and a one-thread-per-feature kernel
 * folds the CTA triples into a running float64 state in fixed order.  Rollout i feeds the policy / value
 * nets and fills the buffer with (obs - mean) / (std + 1e-8) using the statistics of rollouts < i (identity
 * for the first).  Needs a device env (create it first) and a policy net the 64-wide kernels support. */
void ppo_b200_set_obs_norm(PPO* ppo, int enabled);
void ppo_b200_get_obs_norm(const Env* env, float* mean3, float* std3, double* count);
/* mean undiscounted return per episode of the last device rollout (eval_ppo's "R", src/ppo.cu:581) */
float ppo_b200_last_mean_return(PPO* ppo);
/* the three numbers the last eval_ppo call printed ("J: %f R: %f Episodes: %d", src/ppo.cu:581) */
void ppo_b200_last_eval(PPO* ppo, float* J, float* R, int* episodes);
float ppo_b200_last_value_loss(PPO* ppo);
float ppo_b200_last_policy_loss(PPO* ppo);

/* ---- data parallelism (one process per GPU; NCCL all-reduce of the flat gradient) --------------- */
#define PPO_B200_NCCL_ID_BYTES 128
void ppo_b200_dist_unique_id(char id[PPO_B200_NCCL_ID_BYTES]);        /* rank 0 creates, caller broadcasts */
void ppo_b200_dist_init(const char id[PPO_B200_NCCL_ID_BYTES], int rank, int world_size);
void ppo_b200_dist_finalize(void);
int  ppo_b200_dist_rank(void);
int  ppo_b200_dist_world(void);
/* shard mode: 0 = every rank holds the same buffer + permutation and takes rows
 *                 [rank*mb/G, (rank+1)*mb/G) of each global minibatch (1-GPU-equivalent results);
 *             1 = every rank holds its own envs/buffer; the global minibatch is the union of the
 *                 rank-local minibatches (weak scaling). */
void ppo_b200_dist_set_shard_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* PPO_B200_H */
