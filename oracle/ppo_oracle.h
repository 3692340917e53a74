/* oracle/ppo_oracle.h — CPU restatement of the ppo.c training path on plain arrays.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ppo.c_b200/ (the product) may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / reported baseline.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.  Parity is
 * PINNED: tests/test_oracle_vs_golden.py checks each function bit-for-bit against golden vectors
 * produced by the unmodified reference compiled in oracle/_ref (generator: tests/golden/make_golden.py),
 * and tests/test_oracle_vs_ref.py re-checks live whenever oracle/_ref/libppo_ref.so is present.
 *
 * Extensions with NO counterpart in the reference are marked [EXT] (tanh, A>1 corrected gradient
 * index, Pendulum dynamics from the public gymnasium definition, float64 arbiters): for those the
 * parity is "unpinned" by reference outputs and pinned only by hand-derived known answers.
 */
#ifndef PPO_ORACLE_H
#define PPO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ACT_NONE = 0, ORC_ACT_RELU = 1, ORC_ACT_TANH = 2 /* [EXT] */ };

/* ---- MLP: flat parameter vector W0,b0,W1,b1,... (tensor order of adam.cu:25-42) ------------ */
int  orc_param_count(const int* sizes, int num_layers);
void orc_init_params(float* params, const int* sizes, int num_layers);   /* neural_network.cu:40-51 */
int  orc_cache_floats(const int* sizes, int num_layers, int m);
/* cache = layer inputs back to back: x (m*sizes[0]), post-act h1 (m*sizes[1]), ..., output. */
void orc_mlp_forward(const float* params, const int* sizes, const int* acts, int num_layers,
                     const float* x, int m, float* cache);               /* neural_network.cu:163-189 */
void orc_mlp_backward(const float* params, const int* sizes, const int* acts, int num_layers,
                      const float* cache, const float* grad_out, int m, float* grads,
                      float* grad_x0 /* may be NULL */);                 /* neural_network.cu:192-231 */
const float* orc_mlp_output(const float* cache, const int* sizes, int num_layers, int m);

/* ---- GAE / returns / normalisation ---------------------------------------------------------- */
/* ppo.cu:338-368 in the reference's float arithmetic.  The buffer MUST end with a done flag
 * (reference reads advantage[limit] times 0, ppo.cu:346); adv[limit] is taken as 0. */
void orc_gae(const float* reward, const float* v, const float* v_next, const uint8_t* terminated,
             const uint8_t* truncated, int n, float gamma, float lambda, float* adv_raw,
             float* adv_target, float* adv_norm, float* mean_out, float* std_out);
/* [EXT] float64 arbiter of the same recurrence. */
void orc_gae_f64(const float* reward, const float* v, const float* v_next, const uint8_t* terminated,
                 const uint8_t* truncated, int n, float gamma, float lambda, double* adv_raw,
                 double* adv_target, double* adv_norm, double* mean_out, double* std_out);
/* welford_var.h:53-69 host combine of (mean, m2, n) triples, in order. */
void orc_welford_combine(const float* means, const float* m2s, const int* ns, int k,
                         float* mean, float* m2, int* n);

/* ---- permutation / gather ------------------------------------------------------------------- */
void orc_shuffle(int* idx, int limit);                                   /* trajectory_buffer.cu:132-141 */
void orc_get_batch(const int* random_idx, int limit, int batch_idx, int batch_size, int S, int A,
                   const float* state, const float* action, const float* logprob,
                   const float* advantage, const float* adv_target, float* states, float* actions,
                   float* logprobs, float* advantages, float* adv_targets); /* trajectory_buffer.cu:202-220 */

/* ---- Gaussian policy / losses --------------------------------------------------------------- */
void  orc_gaussian_noise(float* out, int n);                             /* policy.cu:46-65 (rand()) */
float orc_log_prob_one(const float* mu, const float* log_std, const float* action, int A); /* policy.cu:67-74 */
void  orc_log_prob(const float* mu, const float* log_std, const float* action, int m, int A, float* out);
/* policy.cu:101-111.  ref_index!=0 reproduces the reference's grad_in[i*A+j] (only valid A==1);
 * ref_index==0 is the corrected per-sample grad_in[i] ([EXT] for A>1). */
void  orc_log_prob_backwards(const float* mu, const float* log_std, const float* action,
                             const float* grad_in, int m, int A, int ref_index, float* grad_mu,
                             float* grad_log_std);
float orc_entropy(const float* log_std, int A);                          /* policy.cu:171-178 */
float orc_policy_loss_and_grad(float* grad_logprob, float* grad_entropy, const float* adv,
                               const float* logprobs, const float* old_logprobs, float entropy,
                               float ent_coeff, float epsilon, int m);   /* ppo.cu:82-107 */
float orc_mse(const float* y, const float* y_true, int m, int n);        /* loss.cu:5-13 */
void  orc_mse_derivative(float* grad, const float* y, const float* y_true, int m, int n); /* loss.cu:16-23 */

/* ---- Adam ----------------------------------------------------------------------------------- */
void orc_adam(float* w, const float* g, float* m, float* v, int n, float lr, float beta1,
              float beta2, int* time_step);                              /* adam.cu:53-74 */

/* ---- environments --------------------------------------------------------------------------- */
/* [EXT] Pendulum-v1 per the public gymnasium definition, float64 state. */
void orc_pendulum_step(double* theta, double* theta_dot, float action, float* obs, float* reward);
void orc_pendulum_obs(double theta, double theta_dot, float* obs);

/* ---- whole update phase on a filled buffer (ppo.cu:395-444 minus the rollout) -------------- */
typedef struct {
    int S, A, num_layers;
    const int* sizes_mu;   /* {S, H.., A} */
    const int* sizes_v;    /* {S, H.., 1} */
    const int* acts;       /* num_layers-1 entries */
    float lr_policy, lr_v, lambda, epsilon, ent_coeff, gamma;
    int batch_size, n_epochs_policy, n_epochs_value;
    int ref_index;         /* see orc_log_prob_backwards */
} OrcConfig;

typedef struct {
    float *mu, *v, *log_std;                 /* parameters (flat) */
    float *m_mu, *v_mu, *m_v, *v_v, *m_ls, *v_ls; /* Adam moments */
    int t_mu, t_v, t_ls;
} OrcModel;

typedef struct {
    int n;                                    /* limit */
    float *state, *next_state, *action, *reward, *logprob, *advantage, *adv_target;
    uint8_t *terminated, *truncated;
} OrcBuffer;

/* GAE (incl. the two V forwards) + n_epochs_value x V-minibatches + n_epochs_policy x policy
 * minibatches, consuming glibc rand() exactly like the reference's shuffles.  If perm_log != NULL
 * every permutation (n ints each) is appended to it.  Returns last policy loss. */
float orc_update(const OrcConfig* cfg, OrcModel* model, OrcBuffer* buf, int* perm_log,
                 float* loss_log /* may be NULL: [v losses..., policy losses...] */);

/* Rollout with the reference bookkeeping (ppo.cu:54-79) on a built-in env:
 * env_id 0 = toy env (env.c:9-33), env_id 1 = [EXT] Pendulum (resets drawn from rand()).
 * Returns the new ring index; buffer arrays have capacity `capacity`. */
int orc_collect(const OrcConfig* cfg, const OrcModel* model, OrcBuffer* buf, int capacity,
                int start_idx, int steps, int env_id);

/* eval_ppo (ppo.cu:560-583): rollout of `steps` transitions from index 0 + the J / R / Episodes statistics it prints. */
void orc_eval(const OrcConfig* cfg, const OrcModel* model, OrcBuffer* buf, int capacity, int steps, int env_id,
              float gamma, float* J_out, float* R_out, int* episodes_out);

/* [EXT] Pendulum behind the reference's Env hooks (include/env.h:7-15): lets the unmodified reference train on it. */
void orc_pendulum_hook_reset(float* obs);
void orc_pendulum_hook_step(float* action, float* obs, float* reward, _Bool* terminated, _Bool* truncated, int action_size);
void orc_pendulum_hook_free(void);

#ifdef __cplusplus
}
#endif
#endif
